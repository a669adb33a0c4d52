"""ctypes binding of the CPU oracle (oracle/pt_oracle.c).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libpt_oracle.so")

RNG_PHILOX, RNG_SEQ, RNG_MOCK = 0, 1, 2
SINCOS_DET, SINCOS_LIBM = 0, 1
ACCUM_FORWARD, ACCUM_RECURSIVE = 0, 1


class RenderCfg(C.Structure):
    _fields_ = [("rng_mode", C.c_int32), ("sincos_mode", C.c_int32), ("accum_mode", C.c_int32),
                ("seed", C.c_uint64), ("threads", C.c_int32), ("shuffle", C.c_int32)]


def build(force: bool = False) -> str:
    src = os.path.join(ORACLE_DIR, "pt_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int32)
    up = C.POINTER(C.c_uint32)
    u64p = C.POINTER(C.c_uint64)
    L.pto_scene_load.restype = C.c_void_p
    L.pto_scene_load.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    L.pto_scene_load_ex.restype = C.c_void_p
    L.pto_scene_load_ex.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    L.pto_scene_free.argtypes = [C.c_void_p]
    L.pto_scene_counts.argtypes = [C.c_void_p, ip, ip, ip, ip]
    L.pto_scene_id.restype = C.c_char_p
    L.pto_scene_id.argtypes = [C.c_void_p]
    L.pto_scene_mesh_bounds.argtypes = [C.c_void_p, C.c_int, fp, fp]
    L.pto_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(RenderCfg), fp, u64p]
    L.pto_render_region.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64,
                                    C.POINTER(RenderCfg), fp, u64p]
    L.pto_resolve.argtypes = [fp, C.c_size_t, C.c_uint64, fp]
    L.pto_to_int_with_gamma_correction.restype = C.c_uint32
    L.pto_to_int_with_gamma_correction.argtypes = [C.c_float]
    L.pto_write_ppm.argtypes = [C.c_char_p, fp, C.c_int, C.c_int, C.c_uint64, C.c_char_p, C.c_uint64]
    L.pto_intersect.argtypes = [C.c_void_p, fp, C.c_int, ip, ip, fp, fp, fp]
    L.pto_primary_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, fp]
    L.pto_primary_hits.argtypes = [C.c_void_p, C.c_int, C.c_int, ip, ip, fp]
    L.pto_camera_frame.argtypes = [C.c_void_p, fp]
    L.pto_radiance_mean.argtypes = [C.c_void_p, fp, C.c_uint64, C.POINTER(RenderCfg), fp]
    L.pto_philox4x32_10.argtypes = [up, up, up]
    L.pto_sincos_det.argtypes = [C.c_float, fp, fp]
    L.pto_vec_dot.restype = C.c_float
    L.pto_vec_dot.argtypes = [fp, fp]
    L.pto_vec_length.restype = C.c_float
    L.pto_vec_length.argtypes = [fp]
    L.pto_vec_cross.argtypes = [fp, fp, fp]
    L.pto_vec_normalize.argtypes = [fp, fp]
    L.pto_vec_divs.argtypes = [fp, C.c_float, fp]
    _lib = L
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def f32(*v):
    return np.array(v, dtype=np.float32)


class OracleScene:
    """A scene loaded by the oracle's own JSON/OFF reader (mod.rs:92-110, load_off.rs)."""

    def __init__(self, json_path: str, base_dir: str | None = None, fan_polygons: bool = False):
        L = lib()
        err = C.create_string_buffer(512)
        base = base_dir if base_dir is not None else ROOT
        self.h = L.pto_scene_load_ex(json_path.encode(), base.encode(), 1 if fan_polygons else 0, err, 512)
        if not self.h:
            raise ValueError(err.value.decode())
        self.path = json_path

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:   # `lib` may already be torn down at interpreter exit
            try:
                lib().pto_scene_free(self.h)
            except Exception:
                pass
            self.h = None

    @property
    def id(self) -> str:
        return lib().pto_scene_id(self.h).decode()

    def counts(self):
        v = [C.c_int32() for _ in range(4)]
        lib().pto_scene_counts(self.h, *[C.byref(x) for x in v])
        return dict(objects=v[0].value, spheres=v[1].value, meshes=v[2].value, triangles=v[3].value)

    def mesh_bounds(self, obj: int):
        p = np.zeros(3, np.float32)
        r = C.c_float()
        if lib().pto_scene_mesh_bounds(self.h, obj, _fp(p), C.byref(r)):
            raise ValueError("not a mesh")
        return p, np.float32(r.value)

    def camera_frame(self):
        o = np.zeros(12, np.float32)
        lib().pto_camera_frame(self.h, _fp(o))
        return o.reshape(4, 3)

    def intersect(self, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        n = rays.shape[0]
        obj = np.empty(n, np.int32)
        tri = np.empty(n, np.int32)
        t = np.empty(n, np.float32)
        pt = np.empty((n, 3), np.float32)
        nr = np.empty((n, 3), np.float32)
        lib().pto_intersect(self.h, _fp(rays), n, _ip(obj), _ip(tri), _fp(t), _fp(pt), _fp(nr))
        return obj, tri, t, pt, nr

    def intersect_mt(self, rays: np.ndarray, threads: int | None = None):
        """intersect() over host threads (ctypes releases the GIL; pto_intersect is re-entrant): the brute-force scan of a
        million-triangle mesh costs ~20 ms per ray."""
        from concurrent.futures import ThreadPoolExecutor
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        n = rays.shape[0]
        threads = max(1, min(threads or os.cpu_count() or 1, max(n // 16, 1)))
        if threads == 1:
            return self.intersect(rays)
        # small interleaved chunks: the cost per ray varies by orders of magnitude (gate miss vs full scan)
        step = max(16, min(512, n // (threads * 8)))
        chunks = [(i, min(i + step, n)) for i in range(0, n, step)]
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(lambda c: self.intersect(rays[c[0]:c[1]]), chunks))
        return tuple(np.concatenate([p[k] for p in parts], 0) for k in range(5))

    def primary_rays(self, W: int, H: int):
        r = np.empty((W * H, 6), np.float32)
        lib().pto_primary_rays(self.h, W, H, _fp(r))
        return r

    def primary_hits(self, W: int, H: int):
        obj = np.empty(W * H, np.int32)
        tri = np.empty(W * H, np.int32)
        t = np.empty(W * H, np.float32)
        lib().pto_primary_hits(self.h, W, H, _ip(obj), _ip(tri), _fp(t))
        return obj, tri, t

    def render_sum(self, W, H, spp_count, spp_begin=0, seed=0, rng=RNG_PHILOX, sincos=SINCOS_DET,
                   accum=ACCUM_FORWARD, threads=None, shuffle=1, sum_in=None, region=None):
        """Returns (sum framebuffer [W*H,3] fp32, stats[4] = segments, sphere, gate, triangle tests)."""
        cfg = RenderCfg(rng, sincos, accum, seed, threads or os.cpu_count() or 1, shuffle)
        fb = np.zeros((W * H, 3), np.float32) if sum_in is None else np.array(sum_in, dtype=np.float32).reshape(W * H, 3)
        stats = (C.c_uint64 * 4)()
        if region is None:
            rc = lib().pto_render(self.h, W, H, spp_begin, spp_count, C.byref(cfg), _fp(fb), stats)
        else:
            rc = lib().pto_render_region(self.h, W, H, region[0], region[1], spp_begin, spp_count, C.byref(cfg), _fp(fb), stats)
        if rc:
            raise RuntimeError("pto_render failed")
        return fb, np.array(list(stats), dtype=np.uint64)

    def radiance_mean(self, ray6, n, seed=0, rng=RNG_SEQ, sincos=SINCOS_LIBM, accum=ACCUM_RECURSIVE):
        cfg = RenderCfg(rng, sincos, accum, seed, 1, 0)
        ray = np.asarray(ray6, np.float32)
        out = np.zeros(3, np.float32)
        lib().pto_radiance_mean(self.h, _fp(ray), n, C.byref(cfg), _fp(out))
        return out


def resolve(sum_fb: np.ndarray, spp: int) -> np.ndarray:
    s = np.ascontiguousarray(sum_fb, np.float32)
    out = np.empty_like(s)
    lib().pto_resolve(_fp(s), s.size, spp, _fp(out))
    return out


def gamma_u8(x: float) -> int:
    return int(lib().pto_to_int_with_gamma_correction(C.c_float(x)))


def philox(ctr, key):
    c = np.array(ctr, np.uint32)
    k = np.array(key, np.uint32)
    o = np.zeros(4, np.uint32)
    up = C.POINTER(C.c_uint32)
    lib().pto_philox4x32_10(c.ctypes.data_as(up), k.ctypes.data_as(up), o.ctypes.data_as(up))
    return o


def sincos_det(x):
    s, c = C.c_float(), C.c_float()
    lib().pto_sincos_det(C.c_float(x), C.byref(s), C.byref(c))
    return np.float32(s.value), np.float32(c.value)
