"""Pins the CPU oracle against every known answer the reference's own tests hold (src/render/test.rs)."""
import math

import numpy as np
import pytest

import oracle_lib as O
from conftest import SCENES, scene_path, sphere_obj

f32 = np.float32


def test_vector_operations():  # test.rs:3-27
    v1, v2 = O.f32(1, 2, 3), O.f32(2, 3, 4)
    L = O.lib()
    assert L.pto_vec_dot(O._fp(v1), O._fp(v2)) == 20.0
    out = np.zeros(3, f32)
    L.pto_vec_cross(O._fp(v1), O._fp(v2), O._fp(out))
    assert out.tolist() == [-1.0, 2.0, -1.0]
    L.pto_vec_normalize(O._fp(O.f32(1, 0, 0)), O._fp(out))
    assert out.tolist() == [1.0, 0.0, 0.0]
    L.pto_vec_normalize(O._fp(O.f32(1, 1, 0)), O._fp(out))
    # the Rust literal 0.7071067811865475 is narrowed to f32 by the compiler
    assert out[0] == f32(0.7071067811865475) and out[1] == f32(0.7071067811865475) and out[2] == 0.0
    assert f32(L.pto_vec_length(O._fp(v1))) == f32(3.7416573867739413)
    L.pto_vec_divs(O._fp(v2), 2.0, O._fp(out))
    assert out.tolist() == [1.0, 1.5, 2.0]


def test_helpers_gamma():  # test.rs:29-35
    assert O.gamma_u8(0.0) == 0
    assert O.gamma_u8(0.5) == 186
    assert O.gamma_u8(0.75) == 224
    assert O.gamma_u8(1.0) == 255
    assert O.gamma_u8(-3.0) == 0 and O.gamma_u8(7.0) == 255  # clamp, mod.rs:58


def _one(scene, ray):
    obj, tri, t, pt, n = scene.intersect(np.array([ray], f32))
    return int(obj[0]), int(tri[0]), float(t[0]), pt[0].tolist(), n[0].tolist()


def test_intersect_scene(kat_scene):  # test.rs:43-69
    sc = O.OracleScene(kat_scene([sphere_obj((0, 0, -3))]))
    assert _one(sc, [0, 0, 0, 0, 0, -1]) == (0, -1, 2.0, [0.0, 0.0, -2.0], [0.0, 0.0, 1.0])


def test_ray_misses_sphere(kat_scene):  # test.rs:72-87
    sc = O.OracleScene(kat_scene([sphere_obj((0, 0, -3))]))
    d = O.f32(1, 0, -1)
    out = np.zeros(3, f32)
    O.lib().pto_vec_normalize(O._fp(d), O._fp(out))
    assert _one(sc, [2, 0, 0, *out.tolist()])[0] == -1


def test_ray_inside_sphere(kat_scene):  # test.rs:90-116
    sc = O.OracleScene(kat_scene([sphere_obj((0, 0, 0))]))
    assert _one(sc, [0, 0, 0, 0, 0, -1]) == (0, -1, 1.0, [0.0, 0.0, -1.0], [0.0, 0.0, -1.0])


def test_ray_tangent_to_sphere(kat_scene):  # test.rs:119-144
    sc = O.OracleScene(kat_scene([sphere_obj((0, 0, -3))]))
    assert _one(sc, [0, 1, 0, 0, 0, -1]) == (0, -1, 3.0, [0.0, 1.0, -3.0], [0.0, 1.0, 0.0])


@pytest.mark.parametrize("cfg", [
    dict(rng=O.RNG_SEQ, sincos=O.SINCOS_LIBM, accum=O.ACCUM_RECURSIVE),   # the reference's behaviour
    dict(rng=O.RNG_PHILOX, sincos=O.SINCOS_DET, accum=O.ACCUM_RECURSIVE),
    dict(rng=O.RNG_PHILOX, sincos=O.SINCOS_DET, accum=O.ACCUM_FORWARD),   # the GPU twin
])
def test_radiance(kat_scene, cfg):  # test.rs:146-183 (asserts mean.x > 0.3); analytic value 50/144
    sc = O.OracleScene(kat_scene([
        sphere_obj((0, 0, -3), color=(1, 0, 0)),
        sphere_obj((0, 0, 10), color=(0, 0, 0), emission=(50, 50, 50))]))
    n = 400_000
    m = sc.radiance_mean([0, 0, 0, 0, 0, -1], n, seed=7, **cfg)
    assert m[0] > 0.3
    sigma = math.sqrt((50 / 144) * (50 - 50 / 144) / n)
    assert abs(m[0] - 50 / 144) < 5 * sigma
    assert m[1] == 0.0 and m[2] == 0.0


def test_philox_known_answers():  # Random123 kat_vectors, philox4x32-10
    assert O.philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox([0xffffffff] * 4, [0xffffffff] * 2).tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sincos_det_accuracy():
    xs = np.linspace(0, 2 * math.pi, 20001).astype(f32)
    worst = 0.0
    for x in xs:
        s, c = O.sincos_det(float(x))
        worst = max(worst, abs(float(s) - math.sin(float(x))), abs(float(c) - math.cos(float(x))))
    assert worst < 2.5e-7  # ~2 ulp at 1.0: the direction distribution is unchanged vs libm


def test_scene_loading_counts_and_mesh_new():
    c = O.OracleScene(scene_path("cornell")).counts()
    assert c == dict(objects=11, spheres=4, meshes=7, triangles=14)
    m = O.OracleScene(scene_path("mesh"))
    assert m.counts() == dict(objects=8, spheres=0, meshes=8, triangles=810 + 14)
    # inline meshes keep the JSON bounding sphere (mod.rs:446); cornell.json object 8
    p, r = O.OracleScene(scene_path("cornell")).mesh_bounds(8)
    assert p.tolist() == [f32(-1.3), f32(-1.0), 0.0] and r == f32(4.920366)
    # MeshFile goes through Mesh::new (mod.rs:478-482: centre = min + max*0.5): mctri.off spans
    # x in [-3,5]*0.16 ... recompute from the file in numpy fp32
    verts = []
    with open(scene_path("mesh").replace("scenes/mesh.json", "meshes/mctri.off")) as f:
        lines = [l.strip() for l in f if l.strip() and not l.strip().startswith("#")]
    nv = int(lines[1].split()[0])
    verts = np.array([[float(t) for t in l.split()] for l in lines[2:2 + nv]], f32) * f32(0.16)
    faces = np.array([[int(t) for t in l.split()[1:4]] for l in lines[2 + nv:]])
    used = verts[np.unique(faces)]
    mn, mx = used.min(0), used.max(0)
    centre = mn + mx * f32(0.5)
    rad = max(f32(np.sqrt(np.sum(((mn - centre) ** 2).astype(f32), dtype=f32))),
              f32(np.sqrt(np.sum(((mx - centre) ** 2).astype(f32), dtype=f32))))
    p, r = m.mesh_bounds(0)
    assert np.array_equal(p, centre)
    assert abs(float(r) - float(rad)) <= 2e-7 * float(rad)


def test_hdodec_rejected(tmp_path):  # load_off.rs:73-76: pentagon faces are an error (reference would panic)
    import json, os
    from conftest import ROOT, CAMERA
    p = tmp_path / "hd.json"
    obj = {"type_": {"MeshFile": {"path": "meshes/hdodec.off", "scale": 1.0}}, "position": [0, 0, 0],
           "material": {"color": [1, 1, 1], "emmission": [0, 0, 0], "reflect_type": "Diffuse"}}
    p.write_text(json.dumps({"id": "hd", "objects": [obj], "camera": CAMERA}))
    with pytest.raises(ValueError, match="Invalid face"):
        O.OracleScene(str(p), ROOT)


def test_exact_image_answers():  # answers derivable from the scenes
    W, H, spp = 96, 64, 8
    fb, st = O.OracleScene(scene_path("cartesian")).render_sum(W, H, spp, seed=3)
    assert not fb.any()                      # no emitter -> exactly black
    ss = O.OracleScene(scene_path("single-sphere"))
    fb, st = ss.render_sum(W, H, spp, seed=3, rng=O.RNG_SEQ, sincos=O.SINCOS_LIBM, accum=O.ACCUM_RECURSIVE)
    img = O.resolve(fb, spp).reshape(H, W, 3)
    obj, _, _ = ss.primary_hits(W, H)
    inside = (obj.reshape(H, W) == 0)
    core = inside.copy()                     # pixels whose 8-neighbourhood is inside the disc
    core[1:-1, 1:-1] &= inside[:-2, 1:-1] & inside[2:, 1:-1] & inside[1:-1, :-2] & inside[1:-1, 2:]
    core[1:-1, 1:-1] &= inside[:-2, :-2] & inside[2:, 2:] & inside[:-2, 2:] & inside[2:, :-2]
    core[0, :] = core[-1, :] = False; core[:, 0] = core[:, -1] = False
    assert core.sum() > 50
    assert (img[core] == 1.0).all()
    far = ~inside
    far[1:-1, 1:-1] &= ~inside[:-2, 1:-1] & ~inside[2:, 1:-1] & ~inside[1:-1, :-2] & ~inside[1:-1, 2:]
    far[1:-1, 1:-1] &= ~inside[:-2, :-2] & ~inside[2:, 2:] & ~inside[:-2, 2:] & ~inside[2:, :-2]
    far[0, :] = far[-1, :] = False; far[:, 0] = far[:, -1] = False
    assert (img[far] == 0.0).all()


def test_mock_random_is_deterministic():  # mod.rs:31-45 / 1017-1018
    sc = O.OracleScene(scene_path("cornell"))
    O.lib().pto_mock_reset()
    a, _ = sc.render_sum(24, 16, 2, rng=O.RNG_MOCK, sincos=O.SINCOS_LIBM, accum=O.ACCUM_RECURSIVE)
    O.lib().pto_mock_reset()
    b, _ = sc.render_sum(24, 16, 2, rng=O.RNG_MOCK, sincos=O.SINCOS_LIBM, accum=O.ACCUM_RECURSIVE)
    assert np.array_equal(a, b) and a.any()


def test_forward_twin_matches_recursive_form():
    """Same draws, same paths: the two accumulation forms differ by fp32 rounding only."""
    for sid in ["cornell", "mesh", "three-spheres"]:
        sc = O.OracleScene(scene_path(sid))
        W, H, spp = 48, 32, 4
        a, sa = sc.render_sum(W, H, spp, seed=11, accum=O.ACCUM_FORWARD)
        b, sb = sc.render_sum(W, H, spp, seed=11, accum=O.ACCUM_RECURSIVE)
        assert sa.tolist() == sb.tolist()          # identical path geometry (integer counters)
        np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-6)


def test_thread_and_batch_invariance():
    sc = O.OracleScene(scene_path("cornell"))
    W, H = 32, 24
    a, sa = sc.render_sum(W, H, 6, seed=5, threads=1, shuffle=0)
    b, sb = sc.render_sum(W, H, 6, seed=5, threads=4, shuffle=1)
    assert np.array_equal(a, b) and sa.tolist() == sb.tolist()
    c, _ = sc.render_sum(W, H, 2, seed=5)
    c, _ = sc.render_sum(W, H, 4, spp_begin=2, seed=5, sum_in=c)
    assert np.array_equal(a, c)                    # sequential sum is batch-invariant


@pytest.mark.parametrize("sid", SCENES)
def test_all_scenes_load(sid):
    sc = O.OracleScene(scene_path(sid))
    assert sc.id == sid
    obj, tri, t = sc.primary_hits(30, 20)
    assert obj.shape == (600,)
