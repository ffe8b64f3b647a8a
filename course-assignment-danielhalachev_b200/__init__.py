"""crt-b200: B200-native renderer core for the Chaos course ray tracer's per-pixel hot path.

This package is a thin ctypes binding over two in-tree native libraries:

  csrc/libcrtb200.so   the CUDA (sm_100a) wavefront renderer behind the C ABI of include/crtb200.h
  csrc/libcrtfront.so  the C++ host front end (scene loader, KD build, camera, PPM, RayTracer mirror), include/crtfront.h

There is no Python or CPU rendering path: every render call goes to the CUDA library and raises when it is
missing or no sm_100 GPU is usable.  (The CPU oracle lives under /oracle and is test infrastructure only.)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# CRT_CORE_LIB: load a differently-tuned build of the same CUDA core (tools/ tuning runs only)
CORE_PATH = os.environ.get("CRT_CORE_LIB") or os.path.join(CSRC, "libcrtb200.so")
FRONT_PATH = os.path.join(CSRC, "libcrtfront.so")

INVALID = 0xFFFFFFFF
MAT_DIFFUSE, MAT_REFLECTIVE, MAT_CONSTANT, MAT_REFRACTIVE = 0, 1, 2, 3
RAY_PRIMARY, RAY_SHADOW, RAY_REFLECTION, RAY_REFRACTION = 0, 1, 2, 3
# RenderOptimization (RayTracer.h:12-23) + the new enumerator
MODE_BVH_BUCKETS_THREADPOOL = 8
MODE_B200_WAVEFRONT = 10


class CrtError(RuntimeError):
    pass


# ---- C structs (must match include/crtb200.h) -----------------------------------------------------------------
class KdNode(C.Structure):
    _fields_ = [("box_min", C.c_float * 3), ("box_max", C.c_float * 3), ("child", C.c_uint32 * 2),
                ("leaf_start", C.c_uint32), ("leaf_count", C.c_uint32)]


class Mesh(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("material", "first_triangle", "n_triangles", "first_vertex", "n_vertices",
                                          "first_node", "n_nodes", "first_leaf_ref", "n_leaf_refs")]


class Material(C.Structure):
    _fields_ = [("type", C.c_uint32), ("smooth_shading", C.c_uint32), ("texture", C.c_uint32),
                ("albedo", C.c_float * 3), ("ior", C.c_float)]


class Texture(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("color_a", C.c_float * 3), ("color_b", C.c_float * 3), ("scalar", C.c_float),
                ("width", C.c_uint32), ("height", C.c_uint32), ("texel_offset", C.c_uint64)]


class Light(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("intensity", C.c_uint32)]


class Scene(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("background", C.c_float * 3),
        ("n_vertices", C.c_uint32), ("vertex_position", C.POINTER(C.c_float)), ("vertex_normal", C.POINTER(C.c_float)),
        ("vertex_uv", C.POINTER(C.c_float)),
        ("n_triangles", C.c_uint32), ("triangle_vertex", C.POINTER(C.c_uint32)), ("triangle_normal", C.POINTER(C.c_float)),
        ("n_meshes", C.c_uint32), ("meshes", C.POINTER(Mesh)),
        ("n_materials", C.c_uint32), ("materials", C.POINTER(Material)),
        ("n_textures", C.c_uint32), ("textures", C.POINTER(Texture)),
        ("n_texels", C.c_uint64), ("texels", C.POINTER(C.c_float)),
        ("n_lights", C.c_uint32), ("lights", C.POINTER(Light)),
        ("n_mesh_nodes", C.c_uint32), ("mesh_nodes", C.POINTER(KdNode)),
        ("n_mesh_leaf_refs", C.c_uint32), ("mesh_leaf_refs", C.POINTER(C.c_uint32)),
        ("n_top_nodes", C.c_uint32), ("top_nodes", C.POINTER(KdNode)),
        ("n_top_leaf_refs", C.c_uint32), ("top_leaf_refs", C.POINTER(C.c_uint32)),
    ]


class Camera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("rotation", C.c_float * 9)]

    @staticmethod
    def make(position=(0.0, 0.0, 0.0), rotation=(1, 0, 0, 0, 1, 0, 0, 0, 1)) -> "Camera":
        cam = Camera()
        cam.position[:] = [float(v) for v in position]
        cam.rotation[:] = [float(v) for v in rotation]
        return cam


class Rect(C.Structure):
    _fields_ = [("row", C.c_uint32), ("col", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32)]


class Options(C.Structure):
    _fields_ = [("max_depth", C.c_uint32), ("shadow_bias", C.c_float), ("reflection_bias", C.c_float),
                ("refraction_bias", C.c_float), ("n_rects", C.c_uint32), ("rects", C.POINTER(Rect)),
                ("traversal", C.c_uint32), ("count_work", C.c_uint32), ("shard_index", C.c_uint32),
                ("shard_count", C.c_uint32), ("shard_full_frame", C.c_uint32)]


class Hit(C.Structure):
    _fields_ = [("mesh", C.c_int32), ("triangle", C.c_int32), ("t", C.c_float)]


HIT_DTYPE = np.dtype([("mesh", np.int32), ("triangle", np.int32), ("t", np.float32)])


class Stats(C.Structure):
    _fields_ = [("rays_primary", C.c_uint64), ("rays_shadow", C.c_uint64), ("rays_reflection", C.c_uint64),
                ("rays_refraction", C.c_uint64), ("node_tests_closest", C.c_uint64), ("triangle_tests_closest", C.c_uint64),
                ("node_tests_shadow", C.c_uint64), ("triangle_tests_shadow", C.c_uint64),
                ("device_ms", C.c_double), ("closest_ms", C.c_double), ("shadow_ms", C.c_double), ("total_ms", C.c_double),
                ("kernel_launches", C.c_uint32), ("levels", C.c_uint32),
                ("handoff_closest", C.c_uint64), ("handoff_shadow", C.c_uint64),
                ("coop_closest_ms", C.c_double), ("coop_shadow_ms", C.c_double), ("shadow_rays_zero_term", C.c_uint64)]

    def as_dict(self) -> dict:
        d = {n: getattr(self, n) for n, _ in self._fields_}
        d["rays_total"] = d["rays_primary"] + d["rays_shadow"] + d["rays_reflection"] + d["rays_refraction"]
        d["node_tests"] = d["node_tests_closest"] + d["node_tests_shadow"]
        d["triangle_tests"] = d["triangle_tests_closest"] + d["triangle_tests_shadow"]
        return d


class SceneInfo(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("bucket_size", C.c_uint32), ("n_meshes", C.c_uint32),
                ("n_materials", C.c_uint32), ("n_textures", C.c_uint32), ("n_lights", C.c_uint32),
                ("n_triangles", C.c_uint64), ("n_vertices", C.c_uint64), ("background", C.c_float * 3),
                ("camera", Camera)]


# ---- library loading ---------------------------------------------------------------------------------------------
_core = None
_front = None


def build(force: bool = False, verbose: bool = False) -> None:
    from . import nativebuild as _b
    _b.build_all(force=force, verbose=verbose)


def core() -> C.CDLL:
    """libcrtb200.so; raises (no fallback) when it has not been built."""
    global _core
    if _core is None:
        if not os.path.exists(CORE_PATH):
            raise CrtError(f"{CORE_PATH} is missing: build it with __graft_entry__.build(); there is no CPU fallback")
        lib = C.CDLL(CORE_PATH, mode=C.RTLD_GLOBAL)
        lib.crtb200_last_error.restype = C.c_char_p
        lib.crtb200_abi_version.restype = C.c_uint32
        lib.crtb200_render_device.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Options), C.c_void_p, C.c_void_p, C.c_void_p]
        lib.crtb200_assemble_shards.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.crtb200_render.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(Options), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        lib.crtb200_render_frames.argtypes = [C.c_void_p, C.POINTER(Camera), C.c_uint32, C.POINTER(Options), C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        lib.crtb200_upload_scene.argtypes = [C.c_void_p, C.POINTER(Scene)]
        lib.crtb200_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        lib.crtb200_create_multi.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        lib.crtb200_device_list.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
        lib.crtb200_last_error_ctx.argtypes = [C.c_void_p]
        lib.crtb200_last_error_ctx.restype = C.c_char_p
        lib.crtb200_destroy.argtypes = [C.c_void_p]
        lib.crtb200_set_queue_budget.argtypes = [C.c_void_p, C.c_uint64]
        lib.crtb200_set_concurrency.argtypes = [C.c_void_p, C.c_uint32]
        lib.crtb200_last_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        lib.crtb200_shard_items.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        lib.crtb200_ipc_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]
        lib.crtb200_ipc_free.argtypes = [C.c_int, C.c_void_p]
        lib.crtb200_ipc_export.argtypes = [C.c_void_p, C.c_void_p]
        lib.crtb200_ipc_open.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        lib.crtb200_ipc_close.argtypes = [C.c_int, C.c_void_p]
        lib.crtb200_generate_rays.argtypes = [C.c_void_p, C.POINTER(Camera), C.c_void_p]
        lib.crtb200_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.crtb200_device_count.argtypes = [C.POINTER(C.c_int)]
        lib.crtb200_debug_powf5.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        _core = lib
    return _core


def front() -> C.CDLL:
    """libcrtfront.so (depends on libcrtb200.so)."""
    global _front
    if _front is None:
        core()
        if not os.path.exists(FRONT_PATH):
            raise CrtError(f"{FRONT_PATH} is missing: build it with __graft_entry__.build()")
        lib = C.CDLL(FRONT_PATH)
        lib.crtfe_last_error.restype = C.c_char_p
        lib.crtfe_scene_load.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
        lib.crtfe_scene_free.argtypes = [C.c_void_p]
        lib.crtfe_scene_get_info.argtypes = [C.c_void_p, C.POINTER(SceneInfo)]
        lib.crtfe_scene_flatten.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.POINTER(Scene)), C.POINTER(C.c_double)]
        lib.crtfe_rectangles.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(Rect), C.c_uint32, C.POINTER(C.c_uint32)]
        lib.crtfe_camera_pan.argtypes = [C.POINTER(Camera), C.c_float]
        lib.crtfe_camera_tilt.argtypes = [C.POINTER(Camera), C.c_float]
        lib.crtfe_camera_roll.argtypes = [C.POINTER(Camera), C.c_float]
        lib.crtfe_camera_truck.argtypes = [C.POINTER(Camera), C.POINTER(C.c_float)]
        lib.crtfe_write_ppm.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
        lib.crtfe_append_f32.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64]
        lib.crtfe_append_u32.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64]
        lib.crtfe_tracer_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        lib.crtfe_tracer_free.argtypes = [C.c_void_p]
        lib.crtfe_tracer_set_camera.argtypes = [C.c_void_p, C.POINTER(Camera)]
        lib.crtfe_tracer_get_camera.argtypes = [C.c_void_p, C.POINTER(Camera)]
        lib.crtfe_tracer_render.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(Stats)]
        _front = lib
    return _front


def _check_core(rc: int) -> None:
    if rc != 0:
        raise CrtError(f"crtb200 error {rc}: {core().crtb200_last_error().decode(errors='replace')}")


def _check_front(rc: int) -> None:
    if rc != 0:
        raise CrtError(f"crtfront error {rc}: {front().crtfe_last_error().decode(errors='replace')}")


# ---- front end ---------------------------------------------------------------------------------------------------
def rectangles(width: int, height: int, bucket_size: int, mode: int = MODE_B200_WAVEFRONT, hw_threads: int = 8):
    """The rectangle grid the reference's schedulers hand to renderRectangle (RayTracer.cpp:114-158)."""
    n = C.c_uint32(0)
    _check_front(front().crtfe_rectangles(width, height, mode, bucket_size, hw_threads, None, 0, C.byref(n)))
    arr = (Rect * max(1, n.value))()
    _check_front(front().crtfe_rectangles(width, height, mode, bucket_size, hw_threads, arr, n.value, C.byref(n)))
    return arr, n.value


class SceneFile:
    """SceneParser::parseScene + AccelerationStructure build + flattening (host front end)."""

    def __init__(self, path_to_scene: str, scene_folder: str = ""):
        self._h = C.c_void_p()
        _check_front(front().crtfe_scene_load(path_to_scene.encode(), scene_folder.encode(), C.byref(self._h)))
        self.info = SceneInfo()
        _check_front(front().crtfe_scene_get_info(self._h, C.byref(self.info)))
        self._flat: Optional[C.POINTER(Scene)] = None
        self.build_seconds = 0.0

    @property
    def handle(self):
        return self._h

    def flatten(self, threads: int = 0) -> "C.POINTER(Scene)":
        if self._flat is None:
            p = C.POINTER(Scene)()
            sec = C.c_double(0)
            _check_front(front().crtfe_scene_flatten(self._h, threads, C.byref(p), C.byref(sec)))
            self._flat = p
            self.build_seconds = sec.value
        return self._flat

    def camera(self) -> Camera:
        cam = Camera()
        C.memmove(C.byref(cam), C.byref(self.info.camera), C.sizeof(Camera))
        return cam

    def rects(self, mode: int = MODE_B200_WAVEFRONT, hw_threads: int = 8):
        return rectangles(self.info.width, self.info.height, self.info.bucket_size, mode, hw_threads)

    def close(self):
        if self._h:
            front().crtfe_scene_free(self._h)
            self._h = C.c_void_p()
            self._flat = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def camera_pan(cam: Camera, degrees: float) -> Camera:
    _check_front(front().crtfe_camera_pan(C.byref(cam), degrees))
    return cam


def write_ppm(path: str, rgb: np.ndarray) -> None:
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    h, w, _ = rgb.shape
    _check_front(front().crtfe_write_ppm(path.encode(), rgb.ctypes.data, w, h))


# ---- CUDA core ---------------------------------------------------------------------------------------------------
def make_options(max_depth: int = 5, rects=None, n_rects: int = 0, traversal: int = 0, count_work: int = 0,
                 shard_index: int = 0, shard_count: int = 1, bias: float = 1e-4, shard_full_frame: bool = False) -> Options:
    o = Options()
    o.max_depth = max_depth
    o.shadow_bias = o.reflection_bias = o.refraction_bias = bias
    o.n_rects = n_rects if rects is not None else 0
    o.rects = C.cast(rects, C.POINTER(Rect)) if rects is not None else None
    o.traversal = traversal
    o.count_work = int(count_work)
    o.shard_index = shard_index
    o.shard_count = shard_count
    o.shard_full_frame = 1 if shard_full_frame else 0
    return o


def ipc_alloc(device: int, nbytes: int) -> int:
    """A whole cudaMalloc allocation on `device` (zeroed), exportable with ipc_export."""
    p = C.c_void_p()
    _check_core(core().crtb200_ipc_alloc(device, nbytes, C.byref(p)))
    return p.value


def ipc_free(device: int, d_ptr: int) -> None:
    _check_core(core().crtb200_ipc_free(device, d_ptr))


def ipc_export(d_ptr: int) -> bytes:
    """64-byte CUDA IPC handle of a cudaMalloc allocation of this process (crtb200_ipc_export)."""
    buf = (C.c_uint8 * 64)()
    _check_core(core().crtb200_ipc_export(d_ptr, buf))
    return bytes(buf)


def ipc_open(device: int, handle: bytes) -> int:
    """Maps another process's allocation on `device`; returns the device pointer (crtb200_ipc_open)."""
    buf = (C.c_uint8 * 64)(*handle)
    p = C.c_void_p()
    _check_core(core().crtb200_ipc_open(device, buf, C.byref(p)))
    return p.value


def ipc_close(device: int, d_ptr: int) -> None:
    _check_core(core().crtb200_ipc_close(device, d_ptr))


class Context:
    """One GPU context of libcrtb200 (RayTracer's device-side state)."""

    def __init__(self, device=0):
        """device: one CUDA device index, or a sequence of them (crtb200_create_multi: every frame is split by tiles
        over the GPUs, the scene is replicated, results are identical to a single GPU's)."""
        self._h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int * len(device))(*[int(d) for d in device])
            _check_core(core().crtb200_create_multi(ids, len(device), C.byref(self._h)))
        else:
            _check_core(core().crtb200_create(int(device), C.byref(self._h)))
        self.width = self.height = 0
        self._scene_keepalive = None

    def devices(self):
        n = C.c_int(0)
        _check_core(core().crtb200_device_list(self._h, None, 0, C.byref(n)))
        ids = (C.c_int * n.value)()
        _check_core(core().crtb200_device_list(self._h, ids, n.value, C.byref(n)))
        return list(ids)

    def last_error(self) -> str:
        return core().crtb200_last_error_ctx(self._h).decode(errors="replace")

    def upload(self, scene_ptr, keepalive=None) -> None:
        _check_core(core().crtb200_upload_scene(self._h, scene_ptr))
        self.width, self.height = scene_ptr.contents.width, scene_ptr.contents.height
        self._scene_keepalive = keepalive

    def set_queue_budget(self, nbytes: int) -> None:
        _check_core(core().crtb200_set_queue_budget(self._h, nbytes))

    def set_concurrency(self, chunks_in_flight: int) -> None:
        _check_core(core().crtb200_set_concurrency(self._h, chunks_in_flight))

    def render(self, camera: Camera, options: Options, want_rgb: bool = True, want_rgb8: bool = False,
               want_hits: bool = False, rgb_out: Optional[np.ndarray] = None, rgb8_out: Optional[np.ndarray] = None):
        """crtb200_render with host buffers.  Returns (rgb f32 HxWx3 | None, rgb8 | None, hits | None, stats dict)."""
        h, w = self.height, self.width
        rgb = rgb_out if rgb_out is not None else (np.zeros((h, w, 3), np.float32) if want_rgb else None)
        rgb8 = rgb8_out if rgb8_out is not None else (np.zeros((h, w, 3), np.uint8) if want_rgb8 else None)
        hits = np.zeros((h, w), HIT_DTYPE) if want_hits else None
        st = Stats()
        _check_core(core().crtb200_render(self._h, C.byref(camera), C.byref(options),
                                          rgb.ctypes.data if rgb is not None else None,
                                          rgb8.ctypes.data if rgb8 is not None else None,
                                          hits.ctypes.data if hits is not None else None, C.byref(st)))
        return rgb, rgb8, hits, st.as_dict()

    def render_frames(self, cameras: Sequence[Camera], options: Options, want_rgb: bool = True, want_rgb8: bool = False):
        n = len(cameras)
        arr = (Camera * n)(*cameras)
        h, w = self.height, self.width
        rgb = np.zeros((n, h, w, 3), np.float32) if want_rgb else None
        rgb8 = np.zeros((n, h, w, 3), np.uint8) if want_rgb8 else None
        st = Stats()
        _check_core(core().crtb200_render_frames(self._h, arr, n, C.byref(options),
                                                 rgb.ctypes.data if rgb is not None else None,
                                                 rgb8.ctypes.data if rgb8 is not None else None, C.byref(st)))
        return rgb, rgb8, st.as_dict()

    def render_frames_into(self, cameras: Sequence[Camera], options: Options, rgb_out: Optional[np.ndarray] = None,
                           rgb8_out: Optional[np.ndarray] = None) -> dict:
        """crtb200_render_frames into caller-owned (ideally pinned) host arrays of shape (n, H, W, 3)."""
        n = len(cameras)
        arr = (Camera * n)(*cameras)
        st = Stats()
        _check_core(core().crtb200_render_frames(self._h, arr, n, C.byref(options),
                                                 rgb_out.ctypes.data if rgb_out is not None else None,
                                                 rgb8_out.ctypes.data if rgb8_out is not None else None, C.byref(st)))
        return st.as_dict()

    def render_device(self, camera: Camera, options: Options, d_rgb: int = 0, d_rgb8: int = 0, stream: int = 0) -> None:
        """Asynchronous; d_rgb / d_rgb8 are raw device pointers (e.g. torch.Tensor.data_ptr())."""
        _check_core(core().crtb200_render_device(self._h, C.byref(camera), C.byref(options), d_rgb or None, d_rgb8 or None,
                                                 stream or None))

    def assemble_shards(self, d_slabs: int, shard_count: int, d_rgb: int = 0, d_rgb8: int = 0, stream: int = 0) -> None:
        _check_core(core().crtb200_assemble_shards(self._h, d_slabs, shard_count, d_rgb or None, d_rgb8 or None, stream or None))

    def shard_items(self, shard_count: int) -> int:
        n = C.c_uint32(0)
        _check_core(core().crtb200_shard_items(self._h, shard_count, C.byref(n)))
        return n.value

    def last_stats(self) -> dict:
        st = Stats()
        _check_core(core().crtb200_last_stats(self._h, C.byref(st)))
        return st.as_dict()

    def generate_rays(self, camera: Camera) -> np.ndarray:
        rays = np.zeros((self.height, self.width, 6), np.float32)
        _check_core(core().crtb200_generate_rays(self._h, C.byref(camera), rays.ctypes.data))
        return rays

    def trace_rays(self, rays: np.ndarray, ray_type: int, max_distance: Optional[np.ndarray] = None, traversal: int = 0):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        n = rays.shape[0]
        if ray_type == RAY_SHADOW:
            md = np.ascontiguousarray(max_distance, dtype=np.float32)
            occ = np.zeros(n, np.uint8)
            _check_core(core().crtb200_trace_rays(self._h, rays.ctypes.data, n, ray_type, traversal, md.ctypes.data, None, occ.ctypes.data))
            return occ
        hits = np.zeros(n, HIT_DTYPE)
        _check_core(core().crtb200_trace_rays(self._h, rays.ctypes.data, n, ray_type, traversal, None, hits.ctypes.data, None))
        return hits

    def debug_powf5(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros_like(x)
        _check_core(core().crtb200_debug_powf5(self._h, x.ctypes.data, x.size, out.ctypes.data))
        return out

    def close(self):
        if self._h:
            core().crtb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RayTracer:
    """Python view of the C++ `crt::RayTracer` mirror class (csrc/frontend/crt_raytracer.hpp):
    RayTracer(scene) -> setCamera -> render(pathToImage, mode, maxDepth)."""

    def __init__(self, scene: SceneFile, device: int = -1):
        """device >= 0: that GPU; -1 (default, like the C++ class): every visible GPU, each frame split by tiles."""
        self.scene = scene
        self._h = C.c_void_p()
        _check_front(front().crtfe_tracer_create(scene.handle, device, C.byref(self._h)))

    def set_camera(self, cam: Camera) -> None:
        _check_front(front().crtfe_tracer_set_camera(self._h, C.byref(cam)))

    def get_camera(self) -> Camera:
        cam = Camera()
        _check_front(front().crtfe_tracer_get_camera(self._h, C.byref(cam)))
        return cam

    def render(self, path_to_image: str = "", mode: int = MODE_B200_WAVEFRONT, max_depth: int = 5, literal: bool = False):
        h, w = self.scene.info.height, self.scene.info.width
        rgb = np.zeros((h, w, 3), np.float32)
        st = Stats()
        _check_front(front().crtfe_tracer_render(self._h, path_to_image.encode(), mode, max_depth, 1 if literal else 0,
                                                 rgb.ctypes.data, C.byref(st)))
        return rgb, st.as_dict()

    def close(self):
        if self._h:
            front().crtfe_tracer_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
