"""Randomised scenes: the CUDA path (both integrators, BVH and brute force) against the CPU oracle, bit for bit."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
f32 = np.float32


def bits(a):
    return np.ascontiguousarray(a, f32).view(np.uint32)


def random_scene(rng, out_dir, idx, n_spheres, n_inline, n_file_tris):
    """Spheres, inline meshes (with deliberately odd bounding spheres so the gate matters) and one OFF mesh."""
    mats = ["Diffuse", "Specular", "Refract"]

    def material(emissive=False):
        kind = mats[int(rng.integers(0, 3))]
        col = rng.uniform(0.1, 0.99, 3) if kind == "Diffuse" else np.full(3, 0.999)
        emi = rng.uniform(0.5, 4.0, 3) if emissive else np.zeros(3)
        return {"color": [float(x) for x in col], "emmission": [float(x) for x in emi], "reflect_type": kind}

    objs = []
    for i in range(n_spheres):
        objs.append({"type_": {"Sphere": {"radius": float(rng.uniform(0.2, 1.2))}}, "position": [float(x) for x in rng.uniform(-3, 3, 3)],
                     "material": material(emissive=(i % 4 == 0))})
    for i in range(n_inline):
        nt = int(rng.integers(1, 6))
        tris = []
        verts = rng.uniform(-2.5, 2.5, (nt, 3, 3))
        for t in verts:
            tris.append({"a": [float(x) for x in t[0]], "b": [float(x) for x in t[1]], "c": [float(x) for x in t[2]]})
        # bounding sphere as the JSON gives it: sometimes generous, sometimes too small (then the gate culls real hits)
        centre = verts.reshape(-1, 3).mean(0)
        rad = float(np.abs(verts.reshape(-1, 3) - centre).max() * rng.choice([0.4, 1.0, 2.0, 5.0]))
        bb = [{"a": [0.0, 0.0, 0.0], "b": [0.0, 0.0, 0.0], "c": [0.0, 0.0, 0.0]}] * 12
        objs.append({"type_": {"Mesh": {"triangles": tris, "bounding_sphere": {"position": [float(x) for x in centre], "radius": rad},
                                        "bounding_box": bb}},
                     "position": [float(x) for x in rng.uniform(-1, 1, 3)], "material": material(emissive=(i % 3 == 0))})
    if idx % 2 == 0:  # a mesh without triangles (never hit; must not disturb the scene stream)
        bb = [{"a": [0.0, 0.0, 0.0], "b": [0.0, 0.0, 0.0], "c": [0.0, 0.0, 0.0]}] * 12
        objs.append({"type_": {"Mesh": {"triangles": [], "bounding_sphere": {"position": [0.0, 0.0, 0.0], "radius": 30.0}, "bounding_box": bb}},
                     "position": [0.0, 0.0, 0.0], "material": material()})
    if n_file_tris:
        off = os.path.join(out_dir, "meshes", f"fuzz{idx}.off")
        os.makedirs(os.path.dirname(off), exist_ok=True)
        nv = n_file_tris + 2
        v = rng.uniform(-1, 1, (nv, 3))
        with open(off, "w") as fh:
            fh.write("OFF\n# fuzz\n\n%d %d 0\n" % (nv, n_file_tris))
            for p in v:
                fh.write("%.6f %.6f %.6f\n" % tuple(p))
            for k in range(n_file_tris):
                a, b, c = rng.choice(nv, 3, replace=False)
                fh.write("3 %d %d %d 0.5 0.5 0.5\n" % (a, b, c))
        objs.append({"type_": {"MeshFile": {"path": f"meshes/fuzz{idx}.off", "scale": float(rng.uniform(0.5, 2.0))}},
                     "position": [float(x) for x in rng.uniform(-1, 1, 3)], "material": material()})
    order = rng.permutation(len(objs))
    objs = [objs[i] for i in order]
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    cam = {"position": [float(x) for x in (-d * 9.0)], "direction": [float(np.float32(x)) for x in d], "focal_length": 0.035,
           "sensor_width": 0.036, "aspect_ratio": 1.5}
    path = os.path.join(out_dir, "scenes", f"fuzz{idx}.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        json.dump({"id": f"fuzz{idx}", "objects": objs, "camera": cam}, fh)
    return path


@pytest.mark.parametrize("idx,n_spheres,n_inline,n_file_tris", [(0, 3, 4, 0), (1, 6, 10, 40), (2, 60, 3, 300), (3, 1, 0, 26), (4, 0, 8, 0)])
def test_random_scene_matches_oracle(tmp_path, idx, n_spheres, n_inline, n_file_tris):
    import path_tracer_rust_b200 as P
    import path_tracer_rust_b200.api as A
    rng = np.random.default_rng(1000 + idx)
    out = str(tmp_path)
    path = random_scene(rng, out, idx, n_spheres, n_inline, n_file_tris)
    sc = P.Scene.load(path, base_dir=out)
    osc = O.OracleScene(path, out)
    n = 100_000
    o = rng.uniform(-4, 4, (n, 3)).astype(f32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(f32)
    rays = np.concatenate([o, d], 1)
    ref = osc.intersect(rays)
    W, H, spp = 64, 48, 6
    ref_fb, ref_st = osc.render_sum(W, H, spp, seed=idx)
    for opts in ({"integrator": 1}, {"integrator": 2, "wavefront_paths": 5000},
                 {"integrator": 1, "bvh_min_tris": 1e18, "bvh_min_spheres": 1e18}, {"integrator": 2, "bvh_min_tris": 2, "bvh_min_spheres": 2},
                 {"integrator": 2, "bvh_leaf_max": 4, "bvh_min_tris": 2, "bvh_min_spheres": 2},
                 {"integrator": 2, "bvh_sah_max_prims": 0, "bvh_min_tris": 2, "bvh_min_spheres": 2},   # device LBVH instead of the host SAH topology
                 {"integrator": 2, "bvh_wide": 1, "bvh_min_tris": 2, "bvh_min_spheres": 2},            # compressed eight-wide BVH in the trace kernel
                 {"integrator": 2, "bvh_wide": 1, "bvh_sah_max_prims": 0, "wavefront_paths": 6000, "wf_trace_threads": 256, "bvh_min_tris": 2,
                  "bvh_min_spheres": 2},
                 {"integrator": 1, "bvh_sah_max_prims": 0, "bvh_leaf_max": 3, "bvh_min_tris": 2, "bvh_min_spheres": 2},
                 {"integrator": 2, "wf_refill": 1, "wf_descend_min": 30, "bvh_top_levels": 2, "bvh_min_tris": 2, "bvh_min_spheres": 2,
                  "wavefront_paths": 7000},
                 {"integrator": 1, "quad_min_ratio": 0.0},      # every one-pair mesh takes the vote-free path, failing gates included
                 {"integrator": 2, "quad_min_ratio": 1e9},      # none does
                 {"integrator": 2, "bvh_top_levels": 3, "wf_trace_threads": 512, "bvh_min_tris": 2, "bvh_min_spheres": 2},
                 {"integrator": 2, "bvh_top_levels": 5, "wf_trace_threads": 1024, "bvh_leaf_max": 1, "wavefront_paths": 5000,
                  "bvh_min_tris": 2, "bvh_min_spheres": 2}):
        be = P.Backend(0)
        try:
            for k, v in opts.items():
                be.set_option(k, v)
            be.upload_scene(sc)
            got = be.intersect(rays)
            for a, b in zip(got, ref):
                assert np.array_equal(bits(a) if a.dtype == f32 else a, bits(b) if b.dtype == f32 else b), opts
            hit = ref[0] >= 0
            if hit.sum() > 10:  # second generation from the surfaces
                d2 = rng.normal(size=(int(hit.sum()), 3))
                d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(f32)
                rays2 = np.concatenate([ref[3][hit], d2], 1)
                for a, b in zip(be.intersect(rays2), osc.intersect(rays2)):
                    assert np.array_equal(bits(a) if a.dtype == f32 else a, bits(b) if b.dtype == f32 else b), opts
            fb = be.render(W, H, spp, seed=idx, out_kind=A.PTB_OUT_SUM)
            assert be.stats()["segments"] == int(ref_st[0]), opts
            assert np.array_equal(bits(fb), bits(ref_fb)), opts
        finally:
            be.close()
