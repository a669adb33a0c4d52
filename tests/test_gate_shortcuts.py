"""The mesh-gate shortcuts of sphere_gate (path_tracer_rust_b200/csrc/pt_device.cuh) against the exact fp32 evaluation.

The reference's gate is `intersect_sphere(..).is_some()` (mod.rs:267-272, :412-438): with s = fl(sqrt(det)),
pass <=> det >= 0 and (fl(b - s) >= eps or fl(b + s) >= eps).  The device skips the square root when one of three
inequalities already decides the outcome; this test replays those inequalities in IEEE fp32 (numpy) on samples
concentrated at the decision boundaries, where a wrong constant would show first.  CPU only.
"""
import numpy as np

EPS = np.float32(1e-4)
K_PASS = np.float32(1.000002)
K_FAIL = np.float32(0.999997)


def exact_gate(b, det):
    with np.errstate(invalid="ignore"):
        s = np.sqrt(det)                      # correctly rounded fp32, like sqrtf / __fsqrt_rn
        return (det >= 0) & (((b - s) >= EPS) | ((b + s) >= EPS))


def shortcut(b, det):
    """-> (decided, value) of the device's early-outs; `decided` False means it evaluates exactly."""
    q = EPS - b
    q2 = q * q
    ok = det >= 0
    p = ok & ((b >= EPS) | (det > q2 * K_PASS))
    f = ok & ~p & (b <= 0) & (det < q2 * K_FAIL)
    return (~ok) | p | f, p


def _check(b, det):
    b = b.astype(np.float32)
    det = det.astype(np.float32)
    decided, value = shortcut(b, det)
    want = exact_gate(b, det)
    bad = decided & (value != want)
    assert not bad.any(), (b[bad][:5], det[bad][:5], value[bad][:5], want[bad][:5])
    return decided.mean()


def test_shortcuts_agree_at_the_boundaries():
    rng = np.random.default_rng(2024)
    n = 4_000_000
    for scale in (1e-6, 1e-4, 1e-2, 1.0, 30.0, 1e3):
        b = (-rng.random(n) * scale).astype(np.float32)             # b <= 0: the sphere centre is behind the ray
        q2 = ((EPS - b) * (EPS - b)).astype(np.float32)
        for width in (1e-7, 2e-6, 1e-5, 1e-3):                       # det within (1 +- width) of (eps - b)^2
            det = q2 * (1.0 + (rng.random(n) * 2 - 1) * width).astype(np.float32)
            _check(b, det)
        b2 = ((rng.random(n) * 2 - 1) * scale).astype(np.float32)    # either sign, det anywhere (also negative)
        det2 = ((rng.random(n) * 2 - 0.5) * scale * scale).astype(np.float32)
        _check(b2, det2)
        b3 = (EPS * (1 + (rng.random(n) * 2 - 1) * 1e-5)).astype(np.float32)   # b at eps
        _check(b3, (rng.random(n) * scale).astype(np.float32))


def test_shortcuts_every_float_near_the_fail_threshold():
    """All consecutive floats det in a window around q^2 * 0.999997, for a sweep of b <= 0."""
    for b in np.float32([-0.0, -1e-7, -1e-4, -3.3e-3, -0.5, -1.0, -7.25, -123.0, -999.0]):
        q = np.float32(EPS - b)
        thr = np.float32(np.float32(q * q) * K_FAIL)
        base = thr.view(np.uint32).astype(np.int64)
        bits = (base + np.arange(-300_000, 300_000)).astype(np.uint32)
        det = bits.view(np.float32)
        _check(np.full(det.shape, b, np.float32), det)


def test_special_values_fall_through_or_fail_like_the_reference():
    b = np.float32([0.0, -1.0, 1.0, np.nan, -np.inf, np.inf, -1e20, 1e20, -1.0])
    det = np.float32([np.nan, np.nan, np.inf, 1.0, np.inf, np.inf, np.inf, np.inf, -0.0])
    with np.errstate(all="ignore"):
        _check(b, det)
