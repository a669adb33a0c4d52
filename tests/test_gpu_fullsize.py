"""GPU parity at BASELINE.json's full sizes:

* the level-8 synthetic scene -- 1 310 720 mesh triangles + 10 000 spheres inside the Cornell walls, world scale S = 10
  (BASELINE.md C5): primary-hit crops of the 3840x2160 frame, random + second-generation surface rays and a lock-step
  framebuffer region against the brute-force oracle (mod.rs:545-616, 631-659), bit for bit;
* deterministic primary hits of every scenes/*.json at 1920x1080 and 3840x2160: 16:9 images through the 1.5-aspect camera
  (mod.rs:833-834 divide x by W and y by H separately), bit for bit;
* the N-rank peer-memory reduce against the rank-ordered oracle partial sums (skipped below 2 devices).
The oracle side is brute force: ~20 ms per ray that passes the big mesh's gate, so the ray counts are sized for ~1-2 minutes on
the GPU box's host cores.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O
from conftest import ROOT, SCENES, scene_path

pytestmark = pytest.mark.gpu
f32 = np.float32


def bits(a):
    return np.ascontiguousarray(a, f32).view(np.uint32)


def same(a, b):
    return np.array_equal(bits(a) if a.dtype == f32 else a, bits(b) if b.dtype == f32 else b)


@pytest.fixture(scope="module")
def be():
    import path_tracer_rust_b200 as P
    b = P.Backend(0)
    yield b
    b.close()


# ---- (b) BASELINE resolutions, all six scenes ---------------------------------------------------------------------------
@pytest.mark.parametrize("sid", SCENES)
@pytest.mark.parametrize("res", [(1920, 1080), (3840, 2160)])
def test_primary_hits_at_baseline_resolutions(be, sid, res):
    import path_tracer_rust_b200 as P
    W, H = res
    be.upload_scene(P.Scene.load(scene_path(sid)))
    osc = O.OracleScene(scene_path(sid))
    g_obj, g_tri, g_t = be.primary_hits(W, H)
    o_obj, o_tri, o_t, _, _ = osc.intersect_mt(osc.primary_rays(W, H))
    assert np.array_equal(g_obj, o_obj), (sid, res, int((g_obj != o_obj).sum()))
    assert np.array_equal(g_tri, o_tri)
    assert np.array_equal(bits(g_t), bits(o_t))
    if sid != "cartesian":
        assert (g_obj >= 0).mean() > 0.01


# ---- (a) the full-size synthetic scene -----------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def synthetic_full(tmp_path_factory):
    """BASELINE config 5 at full size: level 8 (1 310 720 triangles) + 10 000 spheres, S = 10; ~10 s to generate."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tools", "make_synthetic_scene.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    out = str(tmp_path_factory.mktemp("syn_full"))
    path = mk.make_synthetic(out, level=8, n_spheres=10000, scale=10.0)
    import path_tracer_rust_b200 as P
    sc = P.Scene.load(path, base_dir=out)
    osc = O.OracleScene(path, out)
    assert sc.n_triangles == 1310720 + 14 and sc.n_objects == 10008
    return sc, osc


def test_fullsize_synthetic_primary_hit_crops(be, synthetic_full):
    """Windows of the 3840x2160 frame (centre of the mesh, its silhouette, a sphere-filled corner): object, triangle, t bits."""
    sc, osc = synthetic_full
    be.upload_scene(sc)
    st = be.stats()
    assert st["n_bvh_triangles"] == 1310720 and st["n_bvh_spheres"] == 10000 and st["n_bvh_nodes"] > 100_000
    W, H = 3840, 2160
    g_obj, g_tri, g_t = be.primary_hits(W, H)
    rays = osc.primary_rays(W, H)
    mesh = (g_obj == 0).reshape(H, W)                              # object 0 is the million-triangle mesh
    assert mesh.sum() > 20_000, "the mesh must be in view"
    ys, xs = np.nonzero(mesh)
    cx, cy = int(xs.mean()), int(ys.mean())
    edge_x = int(xs.min())                                         # its silhouette (leftmost column that shows it)
    edge_y = int(ys[xs == edge_x].mean())
    clampx = lambda x, w: max(0, min(W - w, x))
    clampy = lambda y, h: max(0, min(H - h, y))
    windows = ((clampx(cx - 32, 64), clampy(cy - 20, 40), 64, 40), (clampx(edge_x - 24, 48), clampy(edge_y - 12, 24), 48, 24),
               (40, 60, 48, 24), (2700, 1500, 48, 24))
    n_checked = n_mesh = 0
    for (x0, y0, w, h) in windows:
        idx = (np.arange(y0, y0 + h)[:, None] * W + np.arange(x0, x0 + w)[None, :]).reshape(-1)
        o_obj, o_tri, o_t, _, _ = osc.intersect_mt(rays[idx])
        assert np.array_equal(g_obj[idx], o_obj), (x0, y0, int((g_obj[idx] != o_obj).sum()))
        assert np.array_equal(g_tri[idx], o_tri), (x0, y0)
        assert np.array_equal(bits(g_t[idx]), bits(o_t)), (x0, y0)
        n_checked += idx.size
        n_mesh += int((o_obj == 0).sum())
    assert n_checked >= 64 * 40 + 3 * 48 * 24 and n_mesh > 300       # (thousands of spheres float in front of the mesh)
    assert np.unique(g_obj).size > 200                                # mesh, many spheres, walls all in view


def test_fullsize_synthetic_rays_and_surface_rays(be, synthetic_full):
    """>= 20 000 random rays (a third aimed at the mesh) + the second generation starting exactly on the hit points."""
    sc, osc = synthetic_full
    be.upload_scene(sc)
    rng = np.random.default_rng(808)
    n = 21_000
    ext = np.array([26.0, 20.0, 88.0], f32)
    o = (rng.uniform(-1, 1, (n, 3)) * ext).astype(f32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(f32)
    centre = np.array([0.0, -20.0 + 16.0 * 1.02, -10.0], f32)
    tgt = centre + rng.normal(size=(n // 3, 3)).astype(f32) * f32(9.0)
    dd = tgt - o[: n // 3]
    d[: n // 3] = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(f32)
    rays = np.concatenate([o, d], 1)
    g = be.intersect(rays)
    ref = osc.intersect_mt(rays)
    for name, a, b in zip(("obj", "tri", "t", "point", "normal"), g, ref):
        assert same(a, b), (name, int((bits(a) != bits(b)).sum()) if a.dtype == f32 else int((a != b).sum()))
    assert (g[1] >= 0).sum() > 2000                      # plenty of hits on the million-triangle mesh
    hit = np.flatnonzero(g[0] >= 0)[:9000]
    d2 = rng.normal(size=(hit.size, 3))
    d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(f32)
    rays2 = np.concatenate([g[3][hit], d2], 1)
    for name, a, b in zip(("obj", "tri", "t", "point", "normal"), be.intersect(rays2), osc.intersect_mt(rays2)):
        assert same(a, b), ("second generation", name)


def test_fullsize_synthetic_lockstep_region(be, synthetic_full):
    """A lock-step framebuffer crop: pixel ranges of a 240x135 frame of the full-size scene, sum bits and segment counts equal
    to the oracle's (pto_render_region), with both integrators."""
    import path_tracer_rust_b200.api as A
    sc, osc = synthetic_full
    W, H, spp = 240, 135, 2
    regions = ((67 * W + 90, 160), (30 * W + 10, 96), (110 * W + 120, 96))
    want = {}
    for (p0, cnt) in regions:
        fb, st = osc.render_sum(W, H, spp, seed=77, region=(p0, cnt))
        want[(p0, cnt)] = fb[p0:p0 + cnt]
    for integ in (2, 1):
        be.set_option("integrator", integ)
        try:
            be.upload_scene(sc)
            g = be.render(W, H, spp, seed=77, out_kind=A.PTB_OUT_SUM)
        finally:
            be.set_option("integrator", 0)
        for (p0, cnt), w in want.items():
            assert np.array_equal(bits(g[p0:p0 + cnt]), bits(w)), (integ, p0)
        assert g.sum() > 0


def test_intersect_accepts_non_unit_directions(be):
    """ptb_intersect takes caller rays as they are (intersect_scene, mod.rs:631-659 does not normalise); the gate's
    origin-inside shortcut, derived for |d| = 1, must not change the answer for scaled directions."""
    import path_tracer_rust_b200 as P
    for sid in ("cornell", "mesh"):
        be.upload_scene(P.Scene.load(scene_path(sid)))
        osc = O.OracleScene(scene_path(sid))
        rng = np.random.default_rng(3)
        n = 60_000
        o = rng.uniform(-2.5, 2.5, (n, 3)).astype(f32)
        d = rng.normal(size=(n, 3))
        d = d / np.linalg.norm(d, axis=1, keepdims=True) * rng.choice([1e-3, 0.03, 0.5, 1.0, 7.0, 300.0], (n, 1))
        rays = np.concatenate([o, d.astype(f32)], 1)
        for name, a, b in zip(("obj", "tri", "t", "point", "normal"), be.intersect(rays), osc.intersect_mt(rays)):
            assert same(a, b), (sid, name)


# ---- (c) N-rank peer-memory reduce (was tools/check_peer_reduce.py) --------------------------------------------------------
def test_peer_reduce_n_ranks_bit_exact():
    """torchrun, one process per GPU: the fused reduce + resolve kernel over CUDA-IPC peer memory must give exactly the
    rank-ordered sum ((fb0 + fb1) + ...) of the oracle's per-shard framebuffers, divided by spp and clamped."""
    import path_tracer_rust_b200 as P
    n = P.load_library().ptb_device_count()
    if n < 2:
        pytest.skip("needs at least 2 CUDA devices")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                        "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tools", "check_peer_reduce.py")],
                       cwd=ROOT, capture_output=True, text=True, timeout=600, env=dict(os.environ, PTB_PEER_CHECK_QUICK="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bit-exact" in r.stdout
