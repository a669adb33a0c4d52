"""ctypes binding of libptb.so + the reference-shaped host API.

Names follow the reference (src/render/mod.rs): RenderConfig :860, Resolution :867, RenderUpdate :882, RenderDone :888,
Image :894, render() :928, intersect_scene :631, gamma_correction :57.  The binding calls the C ABI only; if the
shared library has not been built this module raises -- it never computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
import threading
import time
from typing import Callable, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)

PTB_OK, PTB_CANCELLED = 0, 1
PTB_OUT_MEAN, PTB_OUT_SUM = 0, 1


class BackendError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ptb error {code}: {msg}")
        self.code = code


def library_path() -> str:
    return os.environ.get("PTB_LIBRARY", os.path.join(_HERE, "libptb.so"))


class _Object(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reflect_type", C.c_int32), ("position", C.c_float * 3), ("color", C.c_float * 3),
                ("emission", C.c_float * 3), ("radius", C.c_float), ("bs_position", C.c_float * 3), ("bs_radius", C.c_float),
                ("tri_begin", C.c_uint64), ("tri_count", C.c_uint64)]


class _Triangle(C.Structure):
    _fields_ = [("a", C.c_float * 3), ("b", C.c_float * 3), ("c", C.c_float * 3)]


class _Camera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("direction", C.c_float * 3), ("focal_length", C.c_float),
                ("sensor_width", C.c_float), ("aspect_ratio", C.c_float)]


class _SceneDesc(C.Structure):
    _fields_ = [("objects", C.POINTER(_Object)), ("n_objects", C.c_uint64), ("triangles", C.POINTER(_Triangle)),
                ("n_triangles", C.c_uint64), ("camera", _Camera)]


class Stats(C.Structure):
    _fields_ = [("segments", C.c_uint64), ("samples", C.c_uint64), ("render_ms", C.c_double), ("upload_ms", C.c_double),
                ("bvh_build_ms", C.c_double), ("kernel_launches", C.c_uint32), ("n_loose_objects", C.c_uint32),
                ("n_loose_triangles", C.c_uint32), ("n_bvh_triangles", C.c_uint32), ("n_bvh_spheres", C.c_uint32),
                ("n_bvh_nodes", C.c_uint32), ("bvh_nodes_visited", C.c_uint64), ("bvh_prims_tested", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None
_lib_lock = threading.Lock()

# every symbol include/ptb.h declares
ABI_SYMBOLS = ["ptb_scene_load_json", "ptb_scene_load_json_ex", "ptb_scene_save_json", "ptb_scene_set_camera", "ptb_scene_get_desc", "ptb_scene_id", "ptb_scene_free", "ptb_abi_version",
               "ptb_device_count", "ptb_create", "ptb_destroy", "ptb_last_error", "ptb_upload_scene", "ptb_get_stats", "ptb_set_option", "ptb_selftest",
               "ptb_render", "ptb_render_device", "ptb_resolve_device", "ptb_primary_hits", "ptb_intersect",
               "ptb_to_int_with_gamma_correction", "ptb_write_ppm", "ptb_hash_pixels", "ptb_device_alloc", "ptb_device_free",
               "ptb_device_memset", "ptb_device_to_host", "ptb_device_sync", "ptb_ipc_export", "ptb_ipc_open", "ptb_ipc_close",
               "ptb_peer_reduce_resolve", "ptb_flatten_loose", "ptb_create_multi", "ptb_device_ids", "ptb_render_progressive"]

PREVIEW_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_uint64, C.c_uint64)  # ptb_preview_fn


def load_library():
    """Loads libptb.so; raises (loudly) when the CUDA extension has not been built."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            raise BackendError(-5, f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(path)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.ptb_scene_load_json.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
        L.ptb_scene_load_json_ex.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
        L.ptb_scene_save_json.argtypes = [C.c_void_p, C.c_char_p]
        L.ptb_scene_set_camera.argtypes = [C.c_void_p, C.POINTER(_Camera)]
        L.ptb_scene_get_desc.restype = C.POINTER(_SceneDesc)
        L.ptb_scene_get_desc.argtypes = [C.c_void_p]
        L.ptb_scene_id.restype = C.c_char_p
        L.ptb_scene_id.argtypes = [C.c_void_p]
        L.ptb_scene_free.argtypes = [C.c_void_p]
        L.ptb_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.ptb_create_multi.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        L.ptb_device_ids.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
        L.ptb_render_progressive.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, fp,
                                             C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.c_double, PREVIEW_FN, C.c_void_p]
        L.ptb_destroy.argtypes = [C.c_void_p]
        L.ptb_last_error.restype = C.c_char_p
        L.ptb_last_error.argtypes = [C.c_void_p]
        L.ptb_upload_scene.argtypes = [C.c_void_p, C.POINTER(_SceneDesc)]
        L.ptb_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.ptb_flatten_loose.argtypes = [C.POINTER(_SceneDesc), C.c_double, fp, C.c_uint64, C.POINTER(C.c_uint64), fp, C.c_uint64,
                                        C.POINTER(C.c_uint64)]
        L.ptb_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.ptb_selftest.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.ptb_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, fp,
                                 C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        L.ptb_render_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p,
                                        C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]
        L.ptb_resolve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.ptb_primary_hits.argtypes = [C.c_void_p, C.c_int, C.c_int, ip, ip, fp]
        L.ptb_device_alloc.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        L.ptb_device_free.argtypes = [C.c_void_p, C.c_void_p]
        L.ptb_device_memset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]
        L.ptb_device_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ptb_device_sync.argtypes = [C.c_void_p, C.c_void_p]
        L.ptb_ipc_export.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
        L.ptb_ipc_open.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.ptb_ipc_close.argtypes = [C.c_void_p, C.c_void_p]
        L.ptb_peer_reduce_resolve.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.ptb_intersect.argtypes = [C.c_void_p, fp, C.c_uint64, ip, ip, fp, fp, fp]
        L.ptb_to_int_with_gamma_correction.restype = C.c_uint32
        L.ptb_to_int_with_gamma_correction.argtypes = [C.c_float]
        L.ptb_write_ppm.argtypes = [C.c_char_p, fp, C.c_int, C.c_int, C.c_uint64, C.c_char_p, C.c_uint64]
        L.ptb_hash_pixels.restype = C.c_uint64
        L.ptb_hash_pixels.argtypes = [fp, C.c_uint64]
        if L.ptb_abi_version() != 2:
            raise BackendError(-5, "libptb.so ABI version mismatch")
        _lib = L
        return L


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def to_int_with_gamma_correction(x: float) -> int:  # mod.rs:61-63
    return int(load_library().ptb_to_int_with_gamma_correction(C.c_float(x)))


def gamma_correction(x: float) -> float:  # mod.rs:57-59
    return float(np.float32(np.clip(np.float32(x), 0, 1)) ** np.float32(1.0 / 2.2))


class Scene:
    """SceneData (mod.rs:121-125): loaded from scenes/<id>.json by the library's own host loader, or built from arrays."""

    def __init__(self, handle=None, desc=None, keepalive=None, scene_id="scene"):
        self._h = handle
        self._desc = desc
        self._keep = keepalive
        self._id = scene_id

    @classmethod
    def load(cls, scene: str, base_dir: Optional[str] = None, triangulate_polygons: bool = False) -> "Scene":
        """SceneDescriptor::load(id) (mod.rs:93-98) + to_data(); `scene` is an id under scenes/ or a .json path."""
        L = load_library()
        base = base_dir or REPO_ROOT
        path = scene if scene.endswith(".json") else os.path.join(base, "scenes", f"{scene}.json")
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = L.ptb_scene_load_json_ex(path.encode(), base.encode(), 1 if triangulate_polygons else 0, C.byref(h), err, 512)
        if rc != PTB_OK:
            raise BackendError(rc, err.value.decode())
        return cls(handle=h, desc=L.ptb_scene_get_desc(h), scene_id=L.ptb_scene_id(h).decode())

    @classmethod
    def from_arrays(cls, objects: list, triangles: np.ndarray, camera: dict, scene_id: str = "scene") -> "Scene":
        """What a Rust caller would pass to ptb_upload_scene: objects = list of dicts in SceneObjectData's terms."""
        tris = np.ascontiguousarray(triangles, np.float32).reshape(-1, 9)
        objs = (_Object * max(len(objects), 1))()
        for i, o in enumerate(objects):
            objs[i].kind = 0 if o["kind"] == "sphere" else 1
            objs[i].reflect_type = {"Diffuse": 0, "Specular": 1, "Refract": 2}[o.get("reflect_type", "Diffuse")]
            for k in range(3):
                objs[i].position[k] = o["position"][k]
                objs[i].color[k] = o["color"][k]
                objs[i].emission[k] = o["emission"][k]
                objs[i].bs_position[k] = o.get("bs_position", (0, 0, 0))[k]
            objs[i].radius = o.get("radius", 0.0)
            objs[i].bs_radius = o.get("bs_radius", 0.0)
            objs[i].tri_begin = o.get("tri_begin", 0)
            objs[i].tri_count = o.get("tri_count", 0)
        desc = _SceneDesc()
        desc.objects = C.cast(objs, C.POINTER(_Object))
        desc.n_objects = len(objects)
        desc.triangles = C.cast(tris.ctypes.data, C.POINTER(_Triangle))
        desc.n_triangles = tris.shape[0]
        for k in range(3):
            desc.camera.position[k] = camera["position"][k]
            desc.camera.direction[k] = camera["direction"][k]
        desc.camera.focal_length = camera.get("focal_length", 0.035)
        desc.camera.sensor_width = camera.get("sensor_width", 0.036)
        desc.camera.aspect_ratio = camera.get("aspect_ratio", 1.5)
        return cls(desc=C.pointer(desc), keepalive=(objs, tris, desc), scene_id=scene_id)

    def save(self, path: str):
        """SceneDescriptor::save (mod.rs:112-117) for a scene that came from Scene.load."""
        if not self._h:
            raise BackendError(-5, "only scenes loaded from JSON can be saved")
        rc = load_library().ptb_scene_save_json(self._h, path.encode())
        if rc != PTB_OK:
            raise BackendError(rc, "ptb_scene_save_json failed")

    def set_camera(self, position, direction, focal_length=0.035, sensor_width=0.036, aspect_ratio=1.5):
        if not self._h:
            raise BackendError(-5, "only scenes loaded from JSON can be edited")
        cam = _Camera()
        for k in range(3):
            cam.position[k] = position[k]
            cam.direction[k] = direction[k]
        cam.focal_length, cam.sensor_width, cam.aspect_ratio = focal_length, sensor_width, aspect_ratio
        load_library().ptb_scene_set_camera(self._h, C.byref(cam))

    @property
    def id(self) -> str:
        return self._id

    @property
    def n_objects(self) -> int:
        return int(self._desc.contents.n_objects)

    @property
    def n_triangles(self) -> int:
        return int(self._desc.contents.n_triangles)

    def objects(self):
        d = self._desc.contents
        return [d.objects[i] for i in range(d.n_objects)]

    def flatten_loose(self, quad_min_ratio: float = 0.125):
        """Diagnostic, host only: (object stream [n, 4], triangle records [m, 4]) as fp32 arrays, exactly what ptb_upload_scene
        stages into shared memory when every object stays in the lock-step list (layout: DESIGN.md section 3).  Needs no GPU."""
        L = load_library()
        ns, nt = C.c_uint64(), C.c_uint64()
        rc = L.ptb_flatten_loose(self._desc, quad_min_ratio, None, 0, C.byref(ns), None, 0, C.byref(nt))
        if rc != PTB_OK:
            raise BackendError(rc, L.ptb_last_error(None).decode())
        stream = np.zeros(max(ns.value, 1), np.float32)
        tris = np.zeros(max(nt.value, 1), np.float32)
        rc = L.ptb_flatten_loose(self._desc, quad_min_ratio, _fp(stream), ns.value, C.byref(ns), _fp(tris), nt.value, C.byref(nt))
        if rc != PTB_OK:
            raise BackendError(rc, L.ptb_last_error(None).decode())
        return stream[:ns.value].reshape(-1, 4), tris[:nt.value].reshape(-1, 4)

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                load_library().ptb_scene_free(self._h)
            except Exception:
                pass
            self._h = None


@dataclasses.dataclass
class Resolution:  # mod.rs:867-879
    height: int = 300
    width: int = 300 * 3 // 2


@dataclasses.dataclass
class RenderConfig:  # mod.rs:860-864
    samples_per_pixel: int
    resolution: Resolution
    scene: Scene
    seed: int = 0


@dataclasses.dataclass
class Image:  # mod.rs:894-898; pixels in the reference's buffer order (index i <-> x = i % W, y = H-1 - i // W)
    pixels: np.ndarray
    resolution: Resolution
    hash: int

    def display_rgb8(self) -> np.ndarray:
        """Top-down, left-to-right u8 image exactly as the PPM writer emits it (mod.rs:1065-1076)."""
        L = load_library()
        flat = self.pixels.reshape(-1)
        lut = np.fromiter((L.ptb_to_int_with_gamma_correction(C.c_float(v)) for v in flat), np.uint8, flat.size)
        return lut.reshape(-1, 3)[::-1].reshape(self.resolution.height, self.resolution.width, 3)


@dataclasses.dataclass
class RenderUpdate:  # mod.rs:882-885
    progress: float
    image: Optional[Image]


@dataclasses.dataclass
class RenderDone:  # mod.rs:888-891
    image: Image
    duration: float
    stats: dict


class Backend:
    """One ptb_ctx: one B200 (`Backend(0)`) or several GPUs of the box driven from this process (`Backend([0, 1, 2, 3])`,
    ptb_create_multi: samples per pixel split across the devices, framebuffers summed over peer memory inside render()).
    The context owns the device scene, BVH and framebuffer."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int * len(device))(*device)
            rc = self.L.ptb_create_multi(ids, len(device), C.byref(h))
            self.devices = list(device)
        else:
            rc = self.L.ptb_create(device, C.byref(h))
            self.devices = [device]
        if rc != PTB_OK:
            raise BackendError(rc, self.L.ptb_last_error(None).decode())
        self._h = h
        self.device = self.devices[0]
        self._scene = None

    def close(self):
        if getattr(self, "_h", None):
            self.L.ptb_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc: int):
        if rc < 0:
            raise BackendError(rc, self.L.ptb_last_error(self._h).decode())
        return rc

    def upload_scene(self, scene: Scene):
        self._check(self.L.ptb_upload_scene(self._h, scene._desc))
        self._scene = scene

    def set_option(self, key: str, value: float):
        self._check(self.L.ptb_set_option(self._h, key.encode(), float(value)))

    def selftest(self) -> int:
        bad = C.c_uint64(1)
        self._check(self.L.ptb_selftest(self._h, C.byref(bad)))
        return int(bad.value)

    def stats(self) -> dict:
        s = Stats()
        self.L.ptb_get_stats(self._h, C.byref(s))
        return s.as_dict()

    def render(self, width: int, height: int, spp_count: int, spp_begin: int = 0, seed: int = 0, out_kind: int = PTB_OUT_MEAN,
               cancel=None, samples_done=None, out: Optional[np.ndarray] = None) -> np.ndarray:
        buf = out if out is not None else np.empty((width * height, 3), np.float32)
        cp = C.cast(C.byref(cancel), C.POINTER(C.c_int32)) if cancel is not None else None
        sp = C.cast(C.byref(samples_done), C.POINTER(C.c_uint64)) if samples_done is not None else None
        self.last_rc = self._check(self.L.ptb_render(self._h, width, height, spp_begin, spp_count, seed, out_kind, _fp(buf), cp, sp))
        return buf

    def render_progressive(self, width: int, height: int, spp_count: int, on_preview, preview_interval_ms: float = 500.0,
                           spp_begin: int = 0, seed: int = 0, out_kind: int = PTB_OUT_MEAN, cancel=None, samples_done=None,
                           out: Optional[np.ndarray] = None) -> np.ndarray:
        """ptb_render_progressive: `on_preview(mean_rgb [W*H,3] copy, spp_done, spp_total)` is called between launches, at most
        every `preview_interval_ms`, with the mean of the samples finished so far (RenderUpdate.image, mod.rs:965-982)."""
        buf = out if out is not None else np.empty((width * height, 3), np.float32)
        cp = C.cast(C.byref(cancel), C.POINTER(C.c_int32)) if cancel is not None else None
        sp = C.cast(C.byref(samples_done), C.POINTER(C.c_uint64)) if samples_done is not None else None

        def trampoline(_user, ptr, w, h, done, total):
            on_preview(np.ctypeslib.as_array(ptr, shape=(w * h, 3)).copy(), int(done), int(total))

        cb = PREVIEW_FN(trampoline)
        self.last_rc = self._check(self.L.ptb_render_progressive(self._h, width, height, spp_begin, spp_count, seed, out_kind, _fp(buf),
                                                                 cp, sp, float(preview_interval_ms), cb, None))
        return buf

    def render_device(self, width: int, height: int, spp_count: int, d_sum_ptr: int, spp_begin: int = 0, seed: int = 0,
                      stream: int = 0, sync: bool = False) -> int:
        """Accumulate into a caller-owned device framebuffer (e.g. a torch tensor's data_ptr)."""
        prog = C.c_uint64(0)
        sp = C.byref(prog) if sync else None
        return self._check(self.L.ptb_render_device(self._h, width, height, spp_begin, spp_count, seed, C.c_void_p(d_sum_ptr),
                                                    C.c_void_p(stream), None, C.cast(sp, C.POINTER(C.c_uint64)) if sync else None))

    def resolve_device(self, d_sum_ptr: int, n_floats: int, spp_total: int, d_mean_ptr: int, stream: int = 0):
        self._check(self.L.ptb_resolve_device(self._h, C.c_void_p(d_sum_ptr), n_floats, spp_total, C.c_void_p(d_mean_ptr),
                                              C.c_void_p(stream)))

    # ---- device buffers / CUDA IPC / peer reduce (multi-GPU driver) ----
    def device_alloc(self, n_bytes: int) -> int:
        p = C.c_void_p()
        self._check(self.L.ptb_device_alloc(self._h, n_bytes, C.byref(p)))
        return int(p.value)

    def device_free(self, ptr: int):
        self._check(self.L.ptb_device_free(self._h, C.c_void_p(ptr)))

    def device_memset(self, ptr: int, value: int, n_bytes: int, stream: int = 0):
        self._check(self.L.ptb_device_memset(self._h, C.c_void_p(ptr), value, n_bytes, C.c_void_p(stream)))

    def device_to_host(self, host: np.ndarray, ptr: int, stream: int = 0):
        self._check(self.L.ptb_device_to_host(self._h, host.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), host.nbytes, C.c_void_p(stream)))

    def device_sync(self, stream: int = 0):
        self._check(self.L.ptb_device_sync(self._h, C.c_void_p(stream)))

    def ipc_export(self, ptr: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self.L.ptb_ipc_export(self._h, C.c_void_p(ptr), buf))
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        self._check(self.L.ptb_ipc_open(self._h, handle, C.byref(p)))
        return int(p.value)

    def ipc_close(self, ptr: int):
        self._check(self.L.ptb_ipc_close(self._h, C.c_void_p(ptr)))

    def peer_reduce_resolve(self, peer_ptrs, first_float: int, n_floats: int, spp_total: int, dst_ptr: int, stream: int = 0):
        arr = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
        self._check(self.L.ptb_peer_reduce_resolve(self._h, arr, len(peer_ptrs), first_float, n_floats, spp_total, C.c_void_p(dst_ptr),
                                                   C.c_void_p(stream)))

    def primary_hits(self, width: int, height: int):
        n = width * height
        obj, tri, t = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float32)
        self._check(self.L.ptb_primary_hits(self._h, width, height, _ip(obj), _ip(tri), _fp(t)))
        return obj, tri, t

    def intersect(self, rays: np.ndarray):
        """intersect_scene (mod.rs:631-659) for n rays [n,6] -> obj, tri, t, point, normal."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = rays.shape[0]
        obj, tri, t = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float32)
        pt, nr = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        self._check(self.L.ptb_intersect(self._h, _fp(rays), n, _ip(obj), _ip(tri), _fp(t), _fp(pt), _fp(nr)))
        return obj, tri, t, pt, nr


def hash_pixels(pixels: np.ndarray) -> int:
    p = np.ascontiguousarray(pixels, np.float32)
    return int(load_library().ptb_hash_pixels(_fp(p), p.size // 3))


def write_ppm(path: str, image: Image, spp: int, scene_id: str, seconds: int = 0):
    p = np.ascontiguousarray(image.pixels, np.float32)
    rc = load_library().ptb_write_ppm(path.encode(), _fp(p), image.resolution.width, image.resolution.height, spp,
                                      scene_id.encode(), seconds)
    if rc != PTB_OK:
        raise BackendError(rc, "ptb_write_ppm failed")


def render(render_config: RenderConfig, send_update_progress: Optional[Callable[[RenderUpdate], None]] = None,
           cancel_render: Optional[threading.Event] = None, backend: Optional[Backend] = None,
           progress_interval: float = 0.5) -> RenderDone:
    """render() of the reference (mod.rs:928-1099) on the B200 backend.

    Same contract: blocks until done, reports RenderUpdate{progress in [0,1]} about every 500 ms (mod.rs:965-982),
    stops early when `cancel_render` is set (mod.rs:947-958), returns RenderDone{image, duration}.  Unlike the
    reference the PPM side effect is left to the caller (write_ppm).
    """
    t0 = time.perf_counter()
    own = backend is None
    be = backend or Backend(0)
    try:
        be.upload_scene(render_config.scene)
        res = render_config.resolution
        total = res.width * res.height * render_config.samples_per_pixel
        cancel = C.c_int32(0)
        done = C.c_uint64(0)
        stop = threading.Event()

        def watcher():  # the reference's cancel watcher thread (mod.rs:947-958); progress without an image in between previews
            while not stop.wait(min(0.1, max(progress_interval, 0.01))):
                if cancel_render is not None and cancel_render.is_set():
                    cancel.value = 1

        def on_preview(mean_rgb, spp_done, spp_total):  # RenderUpdate{progress, image} about every 500 ms (mod.rs:965-982)
            if send_update_progress is not None:
                img = Image(pixels=mean_rgb, resolution=res, hash=hash_pixels(mean_rgb))
                send_update_progress(RenderUpdate(progress=spp_done / max(spp_total, 1), image=img))

        th = threading.Thread(target=watcher, daemon=True)
        th.start()
        try:
            pixels = be.render_progressive(res.width, res.height, render_config.samples_per_pixel, on_preview,
                                           preview_interval_ms=progress_interval * 1e3, seed=render_config.seed, cancel=cancel,
                                           samples_done=done)
        finally:
            stop.set()
            th.join()
        img = Image(pixels=pixels, resolution=res, hash=hash_pixels(pixels))
        return RenderDone(image=img, duration=time.perf_counter() - t0, stats=be.stats())
    finally:
        if own:
            be.close()
