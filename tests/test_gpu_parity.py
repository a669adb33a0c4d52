"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.  Run on the B200 box: pytest -m gpu."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_lib as O
from conftest import SCENES, scene_path, sphere_obj

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def be():
    import path_tracer_rust_b200 as P
    b = P.Backend(0)
    yield b
    b.close()


def load_both(be, sid_or_path):
    import path_tracer_rust_b200 as P
    path = sid_or_path if sid_or_path.endswith(".json") else scene_path(sid_or_path)
    sc = P.Scene.load(path)
    be.upload_scene(sc)
    return sc, O.OracleScene(path)


def bits(a):
    return np.ascontiguousarray(a, f32).view(np.uint32)


def test_device_reciprocal_is_correctly_rounded(be):
    """rcp_rn_normal (used in Moeller-Trumbore) == __frcp_rn for every float with exponent field in [1, 252]"""
    assert be.selftest() == 0


# ---- known answers of the reference's own tests, through the CUDA kernels (test.rs:43-144) ------------------
def test_reference_sphere_kats_on_gpu(be, kat_scene):
    load_both(be, kat_scene([sphere_obj((0, 0, -3))], "k1"))
    obj, tri, t, pt, n = be.intersect(np.array([[0, 0, 0, 0, 0, -1], [0, 1, 0, 0, 0, -1],
                                                [2, 0, 0, math.sqrt(0.5), 0, -math.sqrt(0.5)]], f32))
    assert obj.tolist() == [0, 0, -1] and tri.tolist() == [-1, -1, -1]
    assert t[:2].tolist() == [2.0, 3.0]
    assert pt[0].tolist() == [0, 0, -2] and n[0].tolist() == [0, 0, 1]
    assert pt[1].tolist() == [0, 1, -3] and n[1].tolist() == [0, 1, 0]
    load_both(be, kat_scene([sphere_obj((0, 0, 0))], "k2"))
    obj, tri, t, pt, n = be.intersect(np.array([[0, 0, 0, 0, 0, -1]], f32))
    assert (obj[0], t[0], pt[0].tolist(), n[0].tolist()) == (0, 1.0, [0, 0, -1], [0, 0, -1])


# ---- (1) deterministic primary rays: indices identical, t bit-identical -------------------------------------
@pytest.mark.parametrize("sid", SCENES)
@pytest.mark.parametrize("res", [(192, 128), (450, 300)])
def test_primary_hits_bit_exact(be, sid, res):
    _, osc = load_both(be, sid)
    W, H = res
    g_obj, g_tri, g_t = be.primary_hits(W, H)
    o_obj, o_tri, o_t = osc.primary_hits(W, H)
    assert np.array_equal(g_obj, o_obj)
    assert np.array_equal(g_tri, o_tri)
    assert np.array_equal(bits(g_t), bits(o_t))          # criterion allows 1e-5 rel; we get bit-identical
    assert (g_obj >= 0).any() or sid == "none"


# ---- (2) arbitrary rays, including rays that start on surfaces (self-hit behaviour) ----------
def random_rays(rng, n, extent=3.0):
    o = rng.uniform(-extent, extent, (n, 3)).astype(f32)
    d = rng.normal(size=(n, 3)).astype(f32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(f32)
    return np.concatenate([o, d.astype(f32)], 1)


@pytest.mark.parametrize("sid", SCENES)
def test_intersect_arbitrary_rays_bit_exact(be, sid):
    _, osc = load_both(be, sid)
    rng = np.random.default_rng(1234)
    rays = random_rays(rng, 200_000)
    g = be.intersect(rays)
    o = osc.intersect(rays)
    for a, b in zip(g, o):
        assert np.array_equal(bits(a) if a.dtype == f32 else a, bits(b) if b.dtype == f32 else b)
    # second generation: start exactly on the hit points, cosine-ish random directions
    hit = o[0] >= 0
    if hit.sum() > 100:
        d2 = rng.normal(size=(int(hit.sum()), 3)).astype(f32)
        d2 /= np.linalg.norm(d2, axis=1, keepdims=True).astype(f32)
        rays2 = np.concatenate([o[3][hit], d2.astype(f32)], 1)
        g2 = be.intersect(rays2)
        o2 = osc.intersect(rays2)
        for a, b in zip(g2, o2):
            assert np.array_equal(bits(a) if a.dtype == f32 else a, bits(b) if b.dtype == f32 else b)


# ---- (3) lock-step images: same counter RNG, same sin/cos -> bit-identical sum framebuffer ------------------
@pytest.mark.parametrize("sid,W,H,spp", [("cornell", 96, 64, 16), ("mesh", 60, 40, 4), ("single-sphere", 96, 64, 8),
                                         ("two-spheres", 96, 64, 8), ("three-spheres", 96, 64, 8), ("cartesian", 48, 32, 4),
                                         ("cornell", 37, 23, 5)])
@pytest.mark.parametrize("integrator", [1, 2, 3])       # 1 = megakernel, 2 = wavefront, 3 = sample-parallel megakernel (auto picks by scene / size)
def test_lockstep_framebuffer_bit_exact(be, sid, W, H, spp, integrator):
    import path_tracer_rust_b200.api as A
    be.set_option("integrator", integrator)
    try:
        _, osc = load_both(be, sid)
        g = be.render(W, H, spp, seed=42, out_kind=A.PTB_OUT_SUM)
        o, ost = osc.render_sum(W, H, spp, seed=42, rng=O.RNG_PHILOX, sincos=O.SINCOS_DET, accum=O.ACCUM_FORWARD)
        st = be.stats()
        assert st["segments"] == int(ost[0])                 # identical path geometry
        assert st["samples"] == W * H * spp
        assert np.array_equal(bits(g), bits(o))
        gm = be.render(W, H, spp, seed=42, out_kind=A.PTB_OUT_MEAN)
        assert np.array_equal(bits(gm), bits(O.resolve(o, spp)))
    finally:
        be.set_option("integrator", 0)


def test_batch_and_offset_invariance(be):
    import path_tracer_rust_b200.api as A
    _, osc = load_both(be, "cornell")
    W, H = 64, 40
    full = be.render(W, H, 12, seed=9, out_kind=A.PTB_OUT_SUM)
    a = be.render(W, H, 5, spp_begin=0, seed=9, out_kind=A.PTB_OUT_SUM)
    b = be.render(W, H, 7, spp_begin=5, seed=9, out_kind=A.PTB_OUT_SUM)
    o_b, _ = osc.render_sum(W, H, 7, spp_begin=5, seed=9)
    assert np.array_equal(bits(b), bits(o_b))            # global sample indices: shard-invariant streams
    np.testing.assert_allclose(a + b, full, rtol=1e-5, atol=1e-6)
    assert not np.array_equal(full, be.render(W, H, 12, seed=10, out_kind=A.PTB_OUT_SUM))


# ---- (4) images against the reference's own behaviour (sequential RNG, libm sin/cos, recursive radiance) ----
@pytest.mark.parametrize("sid,W,H,spp", [("cornell", 120, 80, 1024), ("three-spheres", 120, 80, 256), ("mesh", 96, 64, 512)])
def test_image_statistics_match_reference_mode(be, sid, W, H, spp):
    """Stated tolerance: per channel RMSE(gpu, ref) <= 1.25 * RMSE(ref_seedA, ref_seedB) and
    |mean_gpu - mean_ref| / mean_ref <= 0.5 % at equal spp (unclamped means).  The sample counts are sized so that 0.5 % is
    about four standard deviations of the difference of two independent image means (3-10 M samples per image)."""
    import path_tracer_rust_b200.api as A
    _, osc = load_both(be, sid)
    g = be.render(W, H, spp, seed=1, out_kind=A.PTB_OUT_SUM) / f32(spp)
    ref = dict(rng=O.RNG_SEQ, sincos=O.SINCOS_LIBM, accum=O.ACCUM_RECURSIVE)
    ra = osc.render_sum(W, H, spp, seed=100, **ref)[0] / f32(spp)
    rb = osc.render_sum(W, H, spp, seed=200, **ref)[0] / f32(spp)
    for c in range(3):
        noise = math.sqrt(float(np.mean((ra[:, c] - rb[:, c]) ** 2)))
        err = math.sqrt(float(np.mean((g[:, c] - ra[:, c]) ** 2)))
        assert err <= 1.25 * noise + 1e-7, (sid, c, err, noise)
        m_ref = 0.5 * (float(ra[:, c].mean()) + float(rb[:, c].mean()))
        if m_ref > 1e-4:
            assert abs(float(g[:, c].mean()) - m_ref) / m_ref <= 0.005, (sid, c, float(g[:, c].mean()), m_ref)


def test_exact_images(be):
    import path_tracer_rust_b200.api as A
    load_both(be, "cartesian")
    assert not be.render(96, 64, 8, seed=3).any()        # no emitter: exactly black
    _, osc = load_both(be, "single-sphere")
    W, H, spp = 96, 64, 8
    img = be.render(W, H, spp, seed=3).reshape(H, W, 3)
    inside = (osc.primary_hits(W, H)[0].reshape(H, W) == 0)
    core = np.zeros_like(inside)
    core[1:-1, 1:-1] = (inside[1:-1, 1:-1] & inside[:-2, 1:-1] & inside[2:, 1:-1] & inside[1:-1, :-2] & inside[1:-1, 2:]
                        & inside[:-2, :-2] & inside[2:, 2:] & inside[:-2, 2:] & inside[2:, :-2])
    assert core.sum() > 50 and (img[core] == 1.0).all()


# ---- behaviour of the boundary --------------------------------------------------------------------------------
def test_api_errors_and_cancel(kat_scene):
    import path_tracer_rust_b200 as P
    import path_tracer_rust_b200.api as A
    b = P.Backend(0)
    try:
        with pytest.raises(P.BackendError) as e:
            b.render(8, 8, 1)
        assert e.value.code == -5                          # PTB_ERR_STATE: no scene yet
        b.upload_scene(P.Scene.load(scene_path("cornell")))
        with pytest.raises(P.BackendError) as e:
            b.render(0, 8, 1)
        assert e.value.code == -1
        cancel = C.c_int32(1)
        done = C.c_uint64(123)
        out = b.render(16, 16, 4, cancel=cancel, samples_done=done, out_kind=A.PTB_OUT_SUM)
        assert b.last_rc == A.PTB_CANCELLED and not out.any()
    finally:
        b.close()
    with pytest.raises(P.BackendError):
        P.Backend(9999)


def test_render_mirror_of_reference_api(be):
    """render(RenderConfig, progress sink, cancel flag) -> RenderDone{image, duration}  (mod.rs:928-934)"""
    import path_tracer_rust_b200 as P
    updates = []
    cfg = P.RenderConfig(samples_per_pixel=8, resolution=P.Resolution(height=60, width=90), scene=P.Scene.load("cornell"), seed=5)
    done = P.render(cfg, send_update_progress=updates.append, backend=be)
    assert done.image.pixels.shape == (90 * 60, 3) and done.image.pixels.max() <= 1.0 and done.image.pixels.min() >= 0.0
    assert done.duration > 0 and done.image.hash != 0
    rgb = done.image.display_rgb8()
    assert rgb.shape == (60, 90, 3)
    # ceiling light is at the top of the displayed image, floor at the bottom (SURVEY.md 8a orientation)
    assert rgb[:5].mean() > rgb[-5:].mean()


def test_progressive_previews_converge_and_final_is_unchanged(be):
    """RenderUpdate{progress, image} (mod.rs:882-886, 965-982): previews arrive during the render with growing progress, each is
    the clamped mean of the samples finished so far (bit-identical to rendering exactly that many samples), and the final image
    is bit-identical to the unpreviewed render."""
    import path_tracer_rust_b200 as P
    be.upload_scene(P.Scene.load("cornell"))
    W, H, spp = 640, 480, 4096          # 2^28 samples per launch when watched: 873 spp per launch, so four previews
    plain = be.render(W, H, spp, seed=11)
    seen = []
    got = be.render_progressive(W, H, spp, lambda img, d, t: seen.append((img, d, t)), preview_interval_ms=0.0, seed=11)
    assert np.array_equal(bits(got), bits(plain))
    assert len(seen) >= 3 and all(t == spp for _, _, t in seen)
    dones = [d for _, d, _ in seen]
    assert dones == sorted(dones) and len(set(dones)) == len(dones) and 0 < dones[0] and dones[-1] < spp
    img, d, _ = seen[0]
    assert img.shape == (W * H, 3) and img.min() >= 0.0 and img.max() <= 1.0
    assert np.array_equal(bits(img), bits(be.render(W, H, d, seed=11)))      # the partial mean of samples [0, d)
    err = [float(np.abs(i - plain).mean()) for i, _, _ in seen]
    assert err[-1] < err[0]                                                  # converging towards the final image
    # the render() mirror forwards them as RenderUpdate.image
    ups = []
    cfg = P.RenderConfig(samples_per_pixel=spp, resolution=P.Resolution(height=H, width=W), scene=P.Scene.load("cornell"), seed=11)
    done = P.render(cfg, send_update_progress=ups.append, backend=be, progress_interval=0.0)
    assert np.array_equal(bits(done.image.pixels), bits(plain))
    assert len(ups) >= 3 and all(u.image is not None and 0.0 < u.progress < 1.0 for u in ups)
    assert [u.progress for u in ups] == sorted(u.progress for u in ups)


def test_unwatched_render_does_not_sync_per_launch(be):
    """ptb_render without cancel / samples_done must not turn interactive: one launch covers 2^31 samples."""
    import path_tracer_rust_b200 as P
    be.upload_scene(P.Scene.load("cornell"))
    be.set_option("integrator", 1)
    try:
        be.render(640, 480, 2000, seed=1)
        assert be.stats()["kernel_launches"] == 2                           # one k_render + the resolve
        done = C.c_uint64(0)
        be.render(640, 480, 2000, seed=1, samples_done=done)
        assert be.stats()["kernel_launches"] == 4 and done.value == 640 * 480 * 2000   # watched: 873 spp per launch
    finally:
        be.set_option("integrator", 0)


def test_multi_gpu_context_bit_exact():
    """ptb_create_multi (the seam stays render(), mod.rs:928-934): one context, several GPUs, no torch.  The image must
    equal the oracle's per-device partial sums added in device order, resolved -- and the single-GPU image up to that order."""
    import path_tracer_rust_b200 as P
    import path_tracer_rust_b200.api as A
    from path_tracer_rust_b200.distributed import shard_samples
    n = P.load_library().ptb_device_count()
    if n < 2:
        pytest.skip("needs at least 2 CUDA devices")
    G = 2 if n < 4 else 4
    W, H, spp = 96, 64, 13
    for sid in ("cornell", "mesh"):
        sc = P.Scene.load(sid)
        osc = O.OracleScene(scene_path(sid))
        mb = P.Backend(list(range(G)))
        try:
            mb.upload_scene(sc)
            acc = None
            for g in range(G):
                b, c = shard_samples(spp, G, g)
                part = osc.render_sum(W, H, c, spp_begin=b, seed=5)[0]
                acc = part if acc is None else (acc + part).astype(f32)
            got_sum = mb.render(W, H, spp, seed=5, out_kind=A.PTB_OUT_SUM)
            assert np.array_equal(bits(got_sum), bits(acc)), sid
            st = mb.stats()
            assert st["samples"] == W * H * spp and st["segments"] == int(osc.render_sum(W, H, spp, seed=5)[1][0])
            got = mb.render(W, H, spp, seed=5)
            assert np.array_equal(bits(got), bits(O.resolve(acc, spp))), sid
            # progress and previews through the multi context
            done = C.c_uint64(0)
            seen = []
            mb.render_progressive(W, H, spp, lambda img, d, t: seen.append(d), preview_interval_ms=0.0, seed=5, samples_done=done)
            assert done.value == W * H * spp
            # a cancel flag that is already set: nothing is rendered, the frame is black, every device reports zero samples
            cancel = C.c_int32(1)
            img = mb.render(W, H, spp, seed=5, cancel=cancel)
            assert mb.last_rc == A.PTB_CANCELLED and not img.any() and mb.stats()["samples"] == 0
            # a multi-GPU context renders through ptb_render; the device-pointer entry point refuses it
            with pytest.raises(A.BackendError):
                mb.render_device(W, H, 1, 0)
        finally:
            mb.close()


# ---- BVH: same closest hit as the brute-force scan and as the oracle, bit for bit -------------------------------
@pytest.fixture(scope="module")
def synthetic_small(tmp_path_factory):
    import importlib.util, os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tools", "make_synthetic_scene.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    out = str(tmp_path_factory.mktemp("syn"))
    return mk.make_synthetic(out, level=3, n_spheres=150, scale=4.0, seed=5), out


def _cmp(a, b):
    for x, y in zip(a, b):
        assert np.array_equal(bits(x) if x.dtype == f32 else x, bits(y) if y.dtype == f32 else y)


@pytest.mark.parametrize("which", ["mesh", "synthetic"])
def test_bvh_equals_bruteforce_equals_oracle(which, synthetic_small):
    import path_tracer_rust_b200 as P
    import path_tracer_rust_b200.api as A
    if which == "mesh":
        path, base, ext = scene_path("mesh"), None, np.array([2.6, 2.0, 8.8], f32)
    else:
        path, base, ext = synthetic_small[0], synthetic_small[1], np.array([10.4, 8.0, 35.2], f32)
    sc = P.Scene.load(path, base_dir=base)
    osc = O.OracleScene(path, base)
    bvh, bf = P.Backend(0), P.Backend(0)
    try:
        bf.set_option("bvh_min_tris", 1e18)
        bf.set_option("bvh_min_spheres", 1e18)
        bvh.upload_scene(sc)
        bf.upload_scene(sc)
        sb, sf = bvh.stats(), bf.stats()
        assert sb["n_bvh_triangles"] > 0 and sb["n_bvh_nodes"] > 0
        assert sf["n_bvh_triangles"] == 0 and sf["n_bvh_spheres"] == 0
        if which == "synthetic":
            assert sb["n_bvh_spheres"] == 150
        rng = np.random.default_rng(77)
        n = 300_000
        o = (rng.uniform(-1, 1, (n, 3)) * ext).astype(f32)
        d = rng.normal(size=(n, 3))
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(f32)
        # a third of the rays aim at the mesh so the BVH actually gets traversed deeply
        centre = np.array([-0.8, -1.5, 0.0], f32) if which == "mesh" else np.array([0.0, -8.0 + 6.4 * 1.02, -4.0], f32)
        tgt = centre + rng.normal(size=(n // 3, 3)).astype(f32) * f32(0.5 if which == "mesh" else 4.0)
        dd = tgt - o[: n // 3]
        d[: n // 3] = (dd / np.linalg.norm(dd, axis=1, keepdims=True)).astype(f32)
        rays = np.concatenate([o, d], 1)
        g_bvh, g_bf = bvh.intersect(rays), bf.intersect(rays)
        _cmp(g_bvh, g_bf)
        ref = osc.intersect(rays[:60_000])
        _cmp([x[:60_000] for x in g_bvh], ref)
        assert (g_bvh[1] >= 0).sum() > 1000              # plenty of triangle hits
        # second generation from the hit points (self-intersection behaviour)
        hit = g_bvh[0] >= 0
        d2 = rng.normal(size=(int(hit.sum()), 3))
        d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(f32)
        rays2 = np.concatenate([g_bvh[3][hit], d2], 1)
        _cmp(bvh.intersect(rays2), bf.intersect(rays2))
        _cmp([x[:40_000] for x in bvh.intersect(rays2)], osc.intersect(rays2[:40_000]))
        # primary rays and a lock-step image
        W, H, spp = 120, 80, 4
        _cmp(bvh.primary_hits(W, H), bf.primary_hits(W, H))
        _cmp(bvh.primary_hits(W, H), osc.primary_hits(W, H))
        a = bvh.render(W, H, spp, seed=3, out_kind=A.PTB_OUT_SUM)
        seg_bvh = bvh.stats()["segments"]
        b = bf.render(W, H, spp, seed=3, out_kind=A.PTB_OUT_SUM)
        assert np.array_equal(bits(a), bits(b)) and seg_bvh == bf.stats()["segments"]
        o_fb, o_st = osc.render_sum(W, H, spp, seed=3)
        assert np.array_equal(bits(a), bits(o_fb)) and seg_bvh == int(o_st[0])
    finally:
        bvh.close()
        bf.close()


# ---- wavefront integrator == megakernel == oracle (same events, same branch sums, same per-pixel order) ---------
@pytest.mark.parametrize("sid,W,H,spp,paths", [("cornell", 96, 64, 16, 96 * 64 * 5), ("cornell", 37, 23, 7, 1024),
                                               ("mesh", 60, 40, 6, 60 * 40 * 4), ("three-spheres", 64, 48, 8, 1 << 23),
                                               ("single-sphere", 50, 30, 3, 4000)])
def test_wavefront_equals_megakernel_equals_oracle(sid, W, H, spp, paths):
    import path_tracer_rust_b200 as P
    import path_tracer_rust_b200.api as A
    sc = P.Scene.load(scene_path(sid))
    osc = O.OracleScene(scene_path(sid))
    wf, mk = P.Backend(0), P.Backend(0)
    try:
        wf.set_option("integrator", 2)
        wf.set_option("wavefront_paths", paths)          # forces several sample batches per frame
        mk.set_option("integrator", 1)
        wf.upload_scene(sc)
        mk.upload_scene(sc)
        a = wf.render(W, H, spp, seed=21, out_kind=A.PTB_OUT_SUM)
        b = mk.render(W, H, spp, seed=21, out_kind=A.PTB_OUT_SUM)
        o, ost = osc.render_sum(W, H, spp, seed=21)
        assert wf.stats()["segments"] == mk.stats()["segments"] == int(ost[0])
        assert wf.stats()["kernel_launches"] > mk.stats()["kernel_launches"]
        assert np.array_equal(bits(a), bits(b))
        assert np.array_equal(bits(a), bits(o))
        # sample offsets: global sample indices make the streams shard-invariant in both integrators
        a2 = wf.render(W, H, 3, spp_begin=4, seed=21, out_kind=A.PTB_OUT_SUM)
        o2, _ = osc.render_sum(W, H, 3, spp_begin=4, seed=21)
        assert np.array_equal(bits(a2), bits(o2))
    finally:
        wf.close()
        mk.close()


def test_sample_parallel_megakernel_equals_oracle(be):
    """Small frames: a lane takes a chunk of a pixel's samples and k_accumulate_samples adds the per-sample radiance in sample order
    (mod.rs:846) -- same bits as one lane per pixel and as the oracle, for any chunking, sample offset and accumulation into a
    non-empty framebuffer; and it is what `auto` picks for the reference's default 450x300 frame."""
    import path_tracer_rust_b200 as P
    import path_tracer_rust_b200.api as A
    for sid, W, H, spp, begin in (("cornell", 45, 30, 37, 0), ("three-spheres", 64, 48, 20, 5), ("cornell", 450, 300, 12, (1 << 32) + 3)):
        sc, osc = load_both(be, sid)
        o, ost = osc.render_sum(W, H, spp, spp_begin=begin, seed=8)
        for integ in (3, 0):
            be.set_option("integrator", integ)
            g = be.render(W, H, spp, spp_begin=begin, seed=8, out_kind=A.PTB_OUT_SUM)
            st = be.stats()
            assert st["segments"] == int(ost[0]) and st["kernel_launches"] == 2, (sid, integ, st["kernel_launches"])
            assert np.array_equal(bits(g), bits(o)), (sid, integ)
    # accumulate into a framebuffer that already holds samples (checkpoint / resume path)
    be.set_option("integrator", 3)
    W, H = 45, 30
    sc, osc = load_both(be, "cornell")
    fb = be.device_alloc(W * H * 12)
    try:
        be.device_memset(fb, 0, W * H * 12)
        be.render_device(W, H, 9, fb, spp_begin=0, seed=8, sync=True)
        be.render_device(W, H, 28, fb, spp_begin=9, seed=8, sync=True)
        got = np.empty((W * H, 3), f32)
        be.device_to_host(got, fb)
        assert np.array_equal(bits(got), bits(osc.render_sum(W, H, 37, seed=8)[0]))
    finally:
        be.device_free(fb)
        be.set_option("integrator", 0)


def test_render_cli_writes_reference_ppm(tmp_path):
    """The resurrected `render <spp> <res_y> <scene>` command (cmd_render.rs:17-44): its PPM must equal, byte for byte, the PPM the
    oracle writes for the same seed (P3, two comment lines, reversed pixel order, gamma 2.2 -> u8; mod.rs:1042-1076)."""
    import os, subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "path_tracer_rust_b200", "render")
    out = tmp_path / "cli.ppm"
    r = subprocess.run([exe, "6", "40", "cornell", "--seed", "13", "--out", str(out)], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Rendering scene cornell (11 objects), 6 samples per pixel, 60x40 resolution" in r.stdout
    osc = O.OracleScene(scene_path("cornell"))
    fb, _ = osc.render_sum(60, 40, 6, seed=13)
    mean = O.resolve(fb, 6)
    ref = tmp_path / "ref.ppm"
    O.lib().pto_write_ppm(str(ref).encode(), O._fp(np.ascontiguousarray(mean)), 60, 40, 6, b"cornell", 0)
    strip = lambda t: [l for l in t.split("\n") if not l.startswith("# rendering time")]
    assert strip(out.read_text()) == strip(ref.read_text())
    bad = subprocess.run([exe, "6", "40", "no-such-scene"], cwd=ROOT, capture_output=True, text=True, timeout=60)
    assert bad.returncode == 1 and "cannot open" in bad.stderr


def test_hdodec_fan_triangulated_scene(be):
    """BASELINE config 2 names hdodec.off; the reference cannot load it (pentagons).  With the opt-in fan triangulation the glass
    dodecahedron + mctri scene renders, and the CUDA path equals the oracle (refraction on mesh triangles, BVH, wavefront)."""
    import path_tracer_rust_b200 as P
    import path_tracer_rust_b200.api as A
    sc = P.Scene.load("mesh-hdodec", triangulate_polygons=True)
    osc = O.OracleScene(scene_path("mesh-hdodec"), fan_polygons=True)
    be.upload_scene(sc)
    W, H, spp = 96, 64, 6
    for a, b in zip(be.primary_hits(W, H), osc.primary_hits(W, H)):
        assert np.array_equal(bits(a) if a.dtype == f32 else a, bits(b) if b.dtype == f32 else b)
    assert (be.primary_hits(W, H)[0] == 1).sum() > 20          # the dodecahedron is in view
    fb = be.render(W, H, spp, seed=4, out_kind=A.PTB_OUT_SUM)
    ofb, ost = osc.render_sum(W, H, spp, seed=4)
    assert be.stats()["segments"] == int(ost[0]) and np.array_equal(bits(fb), bits(ofb))


def test_peer_memory_frame_single_rank(be):
    """The fused reduce + resolve kernel and the IPC export path with one rank (the N-rank case runs in tools/check_peer_reduce.py
    under torchrun: bit-exact against rank-ordered oracle partial sums at 2 and 8 GPUs)."""
    import path_tracer_rust_b200 as P
    import path_tracer_rust_b200.api as A
    be.upload_scene(P.Scene.load("cornell"))
    W, H, spp = 70, 50, 9
    frame = P.PeerMemoryFrame(be, W, H, seed=8, rank=0, world_size=1)
    try:
        assert len(be.ipc_export(frame.fb)) == 64
        img = frame.render(spp)
        want = be.render(W, H, spp, seed=8, out_kind=A.PTB_OUT_MEAN)
        assert np.array_equal(bits(img), bits(want))
        osc = O.OracleScene(scene_path("cornell"))
        assert np.array_equal(bits(img), bits(O.resolve(osc.render_sum(W, H, spp, seed=8)[0], spp)))
    finally:
        frame.close()


def test_checkpoint_resume_is_bit_identical(be):
    """The fp32 sum framebuffer + the number of samples done is a checkpoint: accumulating the remaining
    samples into it later gives the same bits as one uninterrupted render, because the per-pixel sum stays in sample order."""
    import path_tracer_rust_b200 as P
    import path_tracer_rust_b200.api as A
    W, H = 80, 52
    for sid, integ in (("cornell", 1), ("cornell", 2), ("cornell", 3), ("mesh", 0), ("mesh", 3)):
        be.set_option("integrator", integ)
        be.upload_scene(P.Scene.load(sid))
        full = be.render(W, H, 11, seed=6, out_kind=A.PTB_OUT_SUM)
        fb = be.device_alloc(W * H * 12)
        try:
            be.device_memset(fb, 0, W * H * 12)
            be.render_device(W, H, 4, fb, spp_begin=0, seed=6, sync=True)
            ckpt = np.empty((W * H, 3), f32)
            be.device_to_host(ckpt, fb)                      # "save"
            assert np.array_equal(bits(ckpt), bits(be.render(W, H, 4, seed=6, out_kind=A.PTB_OUT_SUM)))
            be.render_device(W, H, 7, fb, spp_begin=4, seed=6, sync=True)   # "resume"
            got = np.empty((W * H, 3), f32)
            be.device_to_host(got, fb)
            assert np.array_equal(bits(got), bits(full)), (sid, integ)
        finally:
            be.device_free(fb)
    be.set_option("integrator", 0)


# ---- edge cases -------------------------------------------------------------------------------------------------------
def test_edge_cases(be, kat_scene):
    import path_tracer_rust_b200 as P
    import path_tracer_rust_b200.api as A
    # empty scene: every ray misses, the image is black, one segment per sample
    path = kat_scene([], "empty")
    sc, osc = load_both(be, path)
    for integ in (1, 2, 3):
        be.set_option("integrator", integ)
        be.upload_scene(sc)
        img = be.render(17, 9, 3, seed=1, out_kind=A.PTB_OUT_SUM)
        assert not img.any() and be.stats()["segments"] == 17 * 9 * 3
        obj, tri, t = be.primary_hits(5, 3)
        assert (obj == -1).all() and (tri == -1).all() and not t.any()
    be.set_option("integrator", 0)
    ofb, ost = osc.render_sum(17, 9, 3, seed=1)
    assert not ofb.any() and int(ost[0]) == 17 * 9 * 3
    # one pixel, one sample; zero rays; zero samples in SUM mode; widths that are not multiples of the 8x4 tile
    sc, osc = load_both(be, "cornell")
    for (W, H, spp) in ((1, 1, 1), (1, 7, 2), (9, 1, 3), (33, 5, 1), (33, 5, 21)):
        for integ in (1, 2, 3):
            be.set_option("integrator", integ)
            g = be.render(W, H, spp, seed=2, out_kind=A.PTB_OUT_SUM)
            o, _ = osc.render_sum(W, H, spp, seed=2)
            assert np.array_equal(bits(g), bits(o)), (W, H, spp, integ)
    be.set_option("integrator", 0)
    assert not be.render(8, 8, 0, seed=2, out_kind=A.PTB_OUT_SUM).any()
    r = be.intersect(np.zeros((0, 6), f32))
    assert all(len(a) == 0 for a in r)
    # large sample indices (beyond 2^32) keep working: the Philox counter carries the high word
    big = (1 << 32) + 5
    g = be.render(12, 8, 2, spp_begin=big, seed=3, out_kind=A.PTB_OUT_SUM)
    o, _ = osc.render_sum(12, 8, 2, spp_begin=big, seed=3)
    assert np.array_equal(bits(g), bits(o))


def test_bvh_padding_holds_for_grazing_rays(synthetic_small):
    """Worst case for the conservative box padding (pt_bvh_build.cu: triangle_pad): rays almost parallel to a triangle (|det| just
    above the reference's 1e-4 cut-off, where the fp32 Moeller-Trumbore error is amplified by 1/det), aimed at its edges, started
    far away.  The BVH must still return exactly what the brute-force scan returns."""
    import ctypes as C
    import path_tracer_rust_b200 as P
    rng = np.random.default_rng(2024)
    for path, base, far in ((scene_path("mesh"), None, 12.0), (synthetic_small[0], synthetic_small[1], 60.0)):
        sc = P.Scene.load(path, base_dir=base)
        d = sc._desc.contents
        mesh = max(range(sc.n_objects), key=lambda i: d.objects[i].tri_count)
        o = d.objects[mesh]
        tris = np.frombuffer(C.string_at(d.triangles, 36 * sc.n_triangles), f32).reshape(-1, 3, 3)[o.tri_begin:o.tri_begin + o.tri_count]
        tris = tris + np.array(list(o.position), f32)
        n = 400_000
        pick = rng.integers(0, len(tris), n)
        A, E1, E2 = tris[pick, 0], tris[pick, 1] - tris[pick, 0], tris[pick, 2] - tris[pick, 0]
        N = np.cross(E1, E2)
        area2 = np.linalg.norm(N, axis=1, keepdims=True)
        nrm = N / area2
        # a point on or just outside an edge of the triangle
        u = rng.choice([0.0, 1.0, 0.5], n) + rng.normal(scale=0.02, size=n)
        v = rng.uniform(-0.02, 1.02, n) * (1 - np.clip(u, 0, 1))
        Pnt = A + E1 * u[:, None] + E2 * v[:, None]
        # in-plane direction plus just enough normal component for |det| = k * 1e-4, k in [0.5, 20]
        phi = rng.uniform(0, 2 * np.pi, n)
        e1n = E1 / np.linalg.norm(E1, axis=1, keepdims=True)
        inpl = e1n * np.cos(phi)[:, None] + np.cross(nrm, e1n) * np.sin(phi)[:, None]
        k = rng.uniform(0.5, 20.0, (n, 1)) * rng.choice([-1.0, 1.0], (n, 1))
        dirs = inpl + nrm * (k * 1e-4 / area2)
        dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
        dist = rng.uniform(0.05, far, (n, 1))
        rays = np.concatenate([Pnt - dirs * dist, dirs], 1).astype(f32)
        bvh, bf = P.Backend(0), P.Backend(0)
        try:
            bf.set_option("bvh_min_tris", 1e18)
            bf.set_option("bvh_min_spheres", 1e18)
            bvh.upload_scene(sc)
            bf.upload_scene(sc)
            assert bvh.stats()["n_bvh_triangles"] > 0 and bf.stats()["n_bvh_triangles"] == 0
            a, b = bvh.intersect(rays), bf.intersect(rays)
            _cmp(a, b)
            assert (a[1] >= 0).sum() > n // 50          # the grazing hits really happen
        finally:
            bvh.close()
            bf.close()
