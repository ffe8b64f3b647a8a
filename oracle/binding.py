"""TEST INFRASTRUCTURE ONLY.  ctypes binding of the CPU oracle (oracle/libcrt_oracle.so, the plain-C restatement)
and helpers to run the compiled UNMODIFIED reference (oracle/_ref/crt_ref, crt_ref_tex).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only --
never by the product package.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcrt_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "crt_ref")
REF_BIN_TEX = os.path.join(HERE, "_ref", "crt_ref_tex")

HIT_DTYPE = np.dtype([("mesh", np.int32), ("triangle", np.int32), ("t", np.float32)])


class OracleStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays_primary", "rays_shadow", "rays_reflection", "rays_refraction",
                                          "node_tests_closest", "triangle_tests_closest", "node_tests_shadow",
                                          "triangle_tests_shadow", "max_query_tests")]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_}
        d["rays_total"] = d["rays_primary"] + d["rays_shadow"] + d["rays_reflection"] + d["rays_refraction"]
        d["node_tests"] = d["node_tests_closest"] + d["node_tests_shadow"]
        d["triangle_tests"] = d["triangle_tests_closest"] + d["triangle_tests_shadow"]
        return d


_lib = None


def build() -> None:
    subprocess.run(["make", "-s", "-C", HERE, "libcrt_oracle.so"], check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        l.crt_oracle_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(OracleStats), C.c_int]
        l.crt_oracle_quantize.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        l.crt_oracle_generate_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.crt_oracle_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        l.crt_oracle_skip_margins.argtypes = [C.c_void_p, C.c_void_p]
        l.crt_oracle_render_skip_model.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.POINTER(OracleStats), C.c_int]
        _lib = l
    return _lib


def render(scene_ptr, camera, options, threads: int = 0, want_hits: bool = True, rgb: np.ndarray | None = None):
    """crt_oracle_render: returns (rgb HxWx3 f32, hits HxW | None, stats dict)."""
    s = scene_ptr.contents
    h, w = s.height, s.width
    if rgb is None:
        rgb = np.zeros((h, w, 3), np.float32)
    hits = np.zeros((h, w), HIT_DTYPE) if want_hits else None
    st = OracleStats()
    if threads <= 0:
        threads = os.cpu_count() or 1
    rc = lib().crt_oracle_render(C.cast(scene_ptr, C.c_void_p), C.byref(camera), C.byref(options), rgb.ctypes.data,
                                 hits.ctypes.data if hits is not None else None, C.byref(st), threads)
    if rc != 0:
        raise RuntimeError("crt_oracle_render failed")
    return rgb, hits, st.as_dict()


def skip_margins(scene_ptr) -> np.ndarray:
    """Per-mesh margins of the skip-rule model (crt_oracle_skip_margins); +inf = the mesh is never culled."""
    mu = np.zeros(max(1, scene_ptr.contents.n_meshes), np.float32)
    if lib().crt_oracle_skip_margins(C.cast(scene_ptr, C.c_void_p), mu.ctypes.data) != 0:
        raise RuntimeError("crt_oracle_skip_margins failed")
    return mu[:scene_ptr.contents.n_meshes]


def render_skip_model(scene_ptr, camera, options, mu: np.ndarray, threads: int = 0, want_hits: bool = True):
    """The restated reference walk with the product's conservative culling applied (CPU model, see crt_oracle.c)."""
    s = scene_ptr.contents
    h, w = s.height, s.width
    rgb = np.zeros((h, w, 3), np.float32)
    hits = np.zeros((h, w), HIT_DTYPE) if want_hits else None
    st = OracleStats()
    mu = np.ascontiguousarray(mu, dtype=np.float32)
    if threads <= 0:
        threads = os.cpu_count() or 1
    rc = lib().crt_oracle_render_skip_model(C.cast(scene_ptr, C.c_void_p), C.byref(camera), C.byref(options), mu.ctypes.data,
                                            rgb.ctypes.data, hits.ctypes.data if hits is not None else None, C.byref(st), threads)
    if rc != 0:
        raise RuntimeError("crt_oracle_render_skip_model failed")
    return rgb, hits, st.as_dict()


def quantize(rgb: np.ndarray) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    out = np.zeros(rgb.shape, np.uint8)
    lib().crt_oracle_quantize(rgb.ctypes.data, rgb.size, out.ctypes.data)
    return out


def generate_rays(scene_ptr, camera) -> np.ndarray:
    s = scene_ptr.contents
    rays = np.zeros((s.height, s.width, 6), np.float32)
    lib().crt_oracle_generate_rays(C.cast(scene_ptr, C.c_void_p), C.byref(camera), rays.ctypes.data)
    return rays


def trace_rays(scene_ptr, rays: np.ndarray, ray_type: int, max_distance: np.ndarray | None = None):
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
    n = rays.shape[0]
    if ray_type == 1:
        md = np.ascontiguousarray(max_distance, dtype=np.float32)
        occ = np.zeros(n, np.uint8)
        lib().crt_oracle_trace_rays(C.cast(scene_ptr, C.c_void_p), rays.ctypes.data, n, ray_type, md.ctypes.data, None, occ.ctypes.data)
        return occ
    hits = np.zeros(n, HIT_DTYPE)
    lib().crt_oracle_trace_rays(C.cast(scene_ptr, C.c_void_p), rays.ctypes.data, n, ray_type, None, hits.ctypes.data, None)
    return hits


# ---- the compiled, unmodified reference ---------------------------------------------------------------------------
def have_reference(textured: bool = False) -> bool:
    return os.path.exists(REF_BIN_TEX if textured else REF_BIN)


def run_reference(scene_file: str, folder: str, out_prefix: str, textured: bool = False, depth: int = 5,
                  hits: bool = True, ppm: bool = True, repeat: int = 1, camera=None, timeout: float = 3600.0,
                  dump_tree: str | None = None) -> dict:
    """Runs oracle/_ref/crt_ref[_tex]; returns its JSON stats plus loaded arrays ('rgb', 'hits', 'ppm_path')."""
    exe = REF_BIN_TEX if textured else REF_BIN
    cmd = [exe, scene_file, folder, out_prefix, "--depth", str(depth), "--repeat", str(repeat)]
    if not hits:
        cmd.append("--no-hits")
    if not ppm:
        cmd.append("--no-ppm")
    if dump_tree:
        cmd += ["--dump-tree", dump_tree]
    if camera is not None:
        cmd += ["--cam"] + ["%.9g" % float(v) for v in list(camera.position) + list(camera.rotation)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"reference failed ({r.returncode}): {r.stderr[-2000:]}")
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    if out_prefix != "-":
        w, h = out["width"], out["height"]
        out["rgb"] = np.fromfile(out_prefix + ".rgbf32", dtype=np.float32).reshape(h, w, 3)
        if hits:
            out["hits"] = np.fromfile(out_prefix + ".hits", dtype=HIT_DTYPE).reshape(h, w)
        out["ppm_path"] = out_prefix + ".ppm"
    return out


def read_reference_trees(path: str):
    """Parses `crt_ref --dump-tree`: list of per-mesh trees + the top-level tree; each a list of
    (box[6], child0, child1, indexes) in the reference's node numbering."""
    data = np.fromfile(path, dtype=np.uint8).tobytes()
    off = 0

    def u32():
        nonlocal off
        v = int(np.frombuffer(data, np.uint32, 1, off)[0])
        off += 4
        return v

    def tree():
        nonlocal off
        nodes = []
        for _ in range(u32()):
            box = np.frombuffer(data, np.float32, 6, off).copy()
            off += 24
            c0, c1, n = u32(), u32(), u32()
            idx = np.frombuffer(data, np.uint32, n, off).copy()
            off += 4 * n
            nodes.append((box, c0, c1, idx))
        return nodes

    n_mesh = u32()
    meshes = [tree() for _ in range(n_mesh)]
    top = tree()
    return meshes, top


def read_ppm_p3(path: str) -> np.ndarray:
    with open(path) as f:
        toks = f.read().split()
    w, h = int(toks[1]), int(toks[2])
    return np.array(toks[4:4 + 3 * w * h], dtype=np.uint16).reshape(h, w, 3).astype(np.uint8)
