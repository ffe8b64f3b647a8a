"""Deterministic `.crtscene` generators for the five BASELINE.json configs.

The reference's own scenes are not in its repository (the homework PDFs only link to them), so the
parity scenes are authored here from closed-form geometry: no RNG, no seeds.  Every number is written
with at most 9 significant digits, so rapidjson's default (non full-precision) number path used by the
reference (SceneParser.cpp:43-45, 79) and strtod in our front end yield the same binary32 value.

Schema as the reference parses it (SceneParser.cpp:88-322).
"""
from __future__ import annotations

import io
import math
import os
import struct
import zlib

import numpy as np

IDENTITY = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]


# ----------------------------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------------------------
def cube_sphere(n: int, center=(0.0, 0.0, 0.0), radius: float = 1.0, displaced: bool = False):
    """Pole-free sphere: the 6 faces of a cube, n x n quads each, projected on the sphere.

    12*n*n triangles, 6*n*n+2 welded vertices, CCW winding seen from outside (the reference culls
    primary rays with dot(d, n) >= 0, Ray.cpp:13).  `displaced` applies SURVEY 8(d) config 4's
    closed-form bumps.  Vertices are rounded to 6 decimals before the centre is added.
    """
    faces = [
        # (origin corner, u axis, v axis) with u x v pointing outward
        ((-1, -1, 1), (2, 0, 0), (0, 2, 0)),   # +z
        ((1, -1, -1), (-2, 0, 0), (0, 2, 0)),  # -z
        ((1, -1, 1), (0, 0, -2), (0, 2, 0)),   # +x
        ((-1, -1, -1), (0, 0, 2), (0, 2, 0)),  # -x
        ((-1, 1, 1), (2, 0, 0), (0, 0, -2)),   # +y
        ((-1, -1, -1), (2, 0, 0), (0, 0, 2)),  # -y
    ]
    ii, jj = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    all_keys = []
    for (o, u, v) in faces:
        # integer lattice coordinates on the cube surface in [0, n]^3 -> exact welding key
        o_i = (np.array(o) + 1) // 2 * n
        u_i = np.sign(u).astype(np.int64)
        v_i = np.sign(v).astype(np.int64)
        k = o_i[None, None, :] + ii[..., None] * u_i[None, None, :] + jj[..., None] * v_i[None, None, :]
        all_keys.append(k.reshape(-1, 3))
    keys = np.concatenate(all_keys, axis=0).astype(np.int64)
    flat = (keys[:, 0] * (n + 1) + keys[:, 1]) * (n + 1) + keys[:, 2]
    uniq, first, inverse = np.unique(flat, return_index=True, return_inverse=True)
    # keep first-appearance order so the vertex numbering is independent of np.unique's sort
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    vid = rank[inverse]
    lattice = keys[first[order]].astype(np.float64)
    cube = lattice / n * 2.0 - 1.0
    p = cube / np.linalg.norm(cube, axis=1, keepdims=True)
    if displaced:
        x, y, z = p[:, 0], p[:, 1], p[:, 2]
        r = 1.0 + 0.08 * np.sin(9 * x) * np.sin(7 * y + 1) * np.sin(11 * z + 2) + 0.03 * np.sin(31 * x + 41 * y + 23 * z)
        p = p * r[:, None]
    p = np.round(p * radius, 6) + np.asarray(center, dtype=np.float64)[None, :]
    verts = p.astype(np.float32)

    tris = []
    stride = (n + 1) * (n + 1)
    qi, qj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    qi = qi.reshape(-1)
    qj = qj.reshape(-1)
    for f in range(6):
        base = f * stride
        a = vid[base + qi * (n + 1) + qj]
        b = vid[base + (qi + 1) * (n + 1) + qj]
        c = vid[base + (qi + 1) * (n + 1) + qj + 1]
        d = vid[base + qi * (n + 1) + qj + 1]
        t = np.empty((qi.size * 2, 3), dtype=np.uint32)
        t[0::2] = np.stack([a, b, c], axis=1)
        t[1::2] = np.stack([a, c, d], axis=1)
        tris.append(t)
    tris = np.concatenate(tris, axis=0)
    return verts, tris


def sphere_uvs(verts: np.ndarray, center) -> np.ndarray:
    """Closed-form spherical UVs in [0,1]^2 (third component 0), rounded to 6 decimals."""
    d = verts.astype(np.float64) - np.asarray(center, dtype=np.float64)[None, :]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    u = 0.5 + np.arctan2(d[:, 0], d[:, 2]) / (2 * math.pi)
    v = 0.5 + np.arcsin(np.clip(d[:, 1], -1, 1)) / math.pi
    uv = np.stack([u, v, np.zeros_like(u)], axis=1)
    return np.round(np.clip(uv, 0.0, 1.0), 6).astype(np.float32)


def uv_sphere(stacks: int, slices: int, center, radius: float):
    """Classic latitude/longitude sphere WITH degenerate pole triangles (SURVEY App. B-3/B-6 regression)."""
    verts = []
    for i in range(stacks + 1):
        th = math.pi * i / stacks
        for j in range(slices + 1):
            ph = 2 * math.pi * j / slices
            verts.append((round(radius * math.sin(th) * math.sin(ph), 6) + center[0],
                          round(radius * math.cos(th), 6) + center[1],
                          round(radius * math.sin(th) * math.cos(ph), 6) + center[2]))
    tris = []
    for i in range(stacks):
        for j in range(slices):
            a = i * (slices + 1) + j
            b = a + slices + 1
            tris.append((a, b, b + 1))
            tris.append((a, b + 1, a + 1))
    return np.asarray(verts, dtype=np.float32), np.asarray(tris, dtype=np.uint32)


def quad(p0, p1, p2, p3):
    """Two triangles (p0,p1,p2),(p0,p2,p3); normal = (p1-p0)x(p2-p0)."""
    verts = np.asarray([p0, p1, p2, p3], dtype=np.float32)
    tris = np.asarray([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    return verts, tris


# ----------------------------------------------------------------------------------------------
# JSON writer (numbers with <= 9 significant digits; ints without decimal point; floats always with one
# where the reference asserts IsFloat(), SceneParser.cpp:178,192,232)
# ----------------------------------------------------------------------------------------------
def _fmt_floats(a: np.ndarray) -> str:
    a = np.asarray(a, dtype=np.float32).reshape(-1)
    if a.size == 0:
        return ""
    return ",".join(np.char.mod("%.9g", a.astype(np.float64)).tolist())


def _fmt_ints(a: np.ndarray) -> str:
    a = np.asarray(a).reshape(-1)
    if a.size == 0:
        return ""
    return ",".join(np.char.mod("%d", a.astype(np.int64)).tolist())


def _ffloat(x: float) -> str:
    s = "%.9g" % float(np.float32(x))
    if "." not in s and "e" not in s and "n" not in s:
        s += ".0"
    return s


def write_crtscene(path: str, scene: dict) -> str:
    """Serialise a scene dict (see the builders below) to `.crtscene` JSON.  Large number arrays go through the front
    end's native formatter when it is built (same "%.9g" / decimal text, ~30x faster than numpy's char.mod)."""
    native = None
    try:
        from . import front as _front
        native = _front()
    except Exception:
        native = None
    f = open(path, "w")

    def w(text: str) -> None:
        f.write(text)

    def numbers(a, is_float: bool) -> None:
        nonlocal f
        arr = np.ascontiguousarray(np.asarray(a).reshape(-1), dtype=np.float32 if is_float else np.uint32)
        if native is None or arr.size < 4096:
            w(_fmt_floats(arr) if is_float else _fmt_ints(arr))
            return
        f.close()
        fn = native.crtfe_append_f32 if is_float else native.crtfe_append_u32
        if fn(path.encode(), arr.ctypes.data, arr.size) != 0:
            raise IOError("native scene writer failed for " + path)
        f = open(path, "a")

    st = scene["settings"]
    w('{"settings":{"background_color":[%s],"image_settings":{"width":%d,"height":%d,"bucket_size":%d}},'
      % (_fmt_floats(st["background_color"]), st["width"], st["height"], st["bucket_size"]))
    cam = scene["camera"]
    w('"camera":{"matrix":[%s],"position":[%s]},' % (_fmt_floats(cam["matrix"]), _fmt_floats(cam["position"])))
    w('"lights":[%s],' % ",".join('{"intensity":%d,"position":[%s]}' % (int(l["intensity"]), _fmt_floats(l["position"]))
                                  for l in scene["lights"]))
    if "textures" in scene:
        parts = []
        for t in scene["textures"]:
            if t["type"] == "albedo":
                parts.append('{"name":"%s","type":"albedo","albedo":[%s]}' % (t["name"], _fmt_floats(t["albedo"])))
            elif t["type"] == "edges":
                parts.append('{"name":"%s","type":"edges","edge_color":[%s],"inner_color":[%s],"edge_width":%s}'
                             % (t["name"], _fmt_floats(t["edge_color"]), _fmt_floats(t["inner_color"]), _ffloat(t["edge_width"])))
            elif t["type"] == "checker":
                parts.append('{"name":"%s","type":"checker","color_A":[%s],"color_B":[%s],"square_size":%s}'
                             % (t["name"], _fmt_floats(t["color_A"]), _fmt_floats(t["color_B"]), _ffloat(t["square_size"])))
            elif t["type"] == "bitmap":
                parts.append('{"name":"%s","type":"bitmap","file_path":"%s"}' % (t["name"], t["file_path"]))
            else:
                raise ValueError(t["type"])
        w('"textures":[%s],' % ",".join(parts))
    parts = []
    for m in scene["materials"]:
        s = '{"type":"%s"' % m["type"]
        if isinstance(m.get("albedo"), str):
            s += ',"albedo":"%s"' % m["albedo"]
        elif m.get("albedo") is not None:
            s += ',"albedo":[%s]' % _fmt_floats(m["albedo"])
        if m["type"] == "refractive":
            s += ',"ior":%s' % _ffloat(m["ior"])
        s += ',"smooth_shading":%s}' % ("true" if m.get("smooth_shading", False) else "false")
        parts.append(s)
    w('"materials":[%s],' % ",".join(parts))
    w('"objects":[')
    for k, o in enumerate(scene["objects"]):
        if k:
            w(",")
        w('{"material_index":%d,"vertices":[' % o["material_index"])
        numbers(o["vertices"], True)
        w('],')
        if "uvs" in o:
            w('"uvs":[')
            numbers(o["uvs"], True)
            w('],')
        w('"triangles":[')
        numbers(o["triangles"], False)
        w(']}')
    w("]}\n")
    f.close()
    return path


# ----------------------------------------------------------------------------------------------
# tiny PNG writer (RGB8, no PIL dependency at test time) and the closed-form bitmap of config 3
# ----------------------------------------------------------------------------------------------
def write_png_rgb(path: str, img: np.ndarray) -> None:
    h, w, c = img.shape
    assert c == 3 and img.dtype == np.uint8
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw, 6)))
        f.write(chunk(b"IEND", b""))


def pattern_bitmap(size: int = 512) -> np.ndarray:
    """(x*4, y*8, (x^y)*4) mod 256 -- SURVEY 8(d) config 3."""
    y, x = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    img = np.stack([(x * 4) % 256, (y * 8) % 256, ((x ^ y) * 4) % 256], axis=2).astype(np.uint8)
    return img


# ----------------------------------------------------------------------------------------------
# scene builders.  Each returns a dict for write_crtscene.
# ----------------------------------------------------------------------------------------------
def _settings(bg, w, h, buckets=24):
    return {"background_color": bg, "width": w, "height": h, "bucket_size": buckets}


def _obj(mat, verts, tris, uvs=None):
    o = {"material_index": mat, "vertices": verts, "triangles": tris}
    if uvs is not None:
        o["uvs"] = uvs
    return o


def hw07_scene0(width=1920, height=1080, sphere_n=0, buckets=24):
    """Config 1: the HW07 triangle + ground quad (+ optional cube-sphere of 12*sphere_n^2 triangles = scene0b)."""
    objs = [
        _obj(0, np.asarray([(-1.75, -1.75, -3), (1.75, -1.75, -3), (0, 1.75, -3)], dtype=np.float32),
             np.asarray([[0, 1, 2]], dtype=np.uint32)),
        _obj(1, *quad((-6, -1.8, 1), (6, -1.8, 1), (6, -1.8, -12), (-6, -1.8, -12))),
    ]
    if sphere_n:
        objs.append(_obj(0, *cube_sphere(sphere_n, (1.4, -0.9, -2.2), 0.6)))
    return {
        "settings": _settings([0.0, 0.5, 0.0], width, height, buckets),
        "camera": {"matrix": IDENTITY, "position": [0.0, 0.0, 0.0]},
        "lights": [{"intensity": 90, "position": [0.0, 3.0, -1.0]}, {"intensity": 40, "position": [-2.5, 1.0, 0.5]}],
        "materials": [{"type": "diffuse", "albedo": [0.9, 0.2, 0.1], "smooth_shading": False},
                      {"type": "diffuse", "albedo": [0.6, 0.6, 0.6], "smooth_shading": False}],
        "objects": objs,
    }


def hw11_room(width=1920, height=1080, sphere_n=32, buckets=24, depth_note=5):
    """Config 2: Cornell-style room, reflective + refractive (ior 1.5) smooth cube-spheres."""
    a, b, zf, zb = -2.0, 2.0, -2.0, -7.0
    grey, blue, green = [0.7, 0.7, 0.7], [0.1, 0.1, 0.9], [0.1, 0.9, 0.1]
    objs = [
        _obj(0, *quad((a, a, zf), (b, a, zf), (b, a, zb), (a, a, zb))),      # floor, normal +y
        _obj(0, *quad((a, b, zb), (b, b, zb), (b, b, zf), (a, b, zf))),      # ceiling, normal -y
        _obj(0, *quad((a, a, zb), (b, a, zb), (b, b, zb), (a, b, zb))),      # back wall, normal +z
        _obj(1, *quad((a, a, zf), (a, a, zb), (a, b, zb), (a, b, zf))),      # left, normal +x
        _obj(2, *quad((b, a, zb), (b, a, zf), (b, b, zf), (b, b, zb))),      # right, normal -x
        _obj(3, *cube_sphere(sphere_n, (0.9, -0.2, -5.2), 0.7)),
        _obj(4, *cube_sphere(sphere_n, (-0.8, -1.2, -4.0), 0.8)),
    ]
    return {
        "settings": _settings([0.0, 0.0, 0.0], width, height, buckets),
        "camera": {"matrix": IDENTITY, "position": [0.0, 0.0, 0.0]},
        "lights": [{"intensity": 120, "position": [0.0, 1.6, -4.5]}, {"intensity": 40, "position": [0.0, 0.0, -1.0]}],
        "materials": [{"type": "diffuse", "albedo": grey, "smooth_shading": False},
                      {"type": "diffuse", "albedo": blue, "smooth_shading": False},
                      {"type": "diffuse", "albedo": green, "smooth_shading": False},
                      {"type": "reflective", "albedo": [0.9, 0.9, 0.9], "smooth_shading": True},
                      {"type": "refractive", "albedo": None, "ior": 1.5, "smooth_shading": True}],
        "objects": objs,
    }


def hw12_textures(width=1920, height=1080, sphere_n=0, buckets=24, bitmap_path="/crt_pattern.png"):
    """Config 3 (USE_TEXTURES flavour): four textured quads (+ optional textured cube-sphere)."""
    def tq(cx, cy, s=0.75, z=-3.0):
        v, t = quad((cx - s, cy - s, z), (cx + s, cy - s, z), (cx + s, cy + s, z), (cx - s, cy + s, z))
        uv = np.asarray([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)], dtype=np.float32)
        return v, t, uv

    objs = []
    for mat, (cx, cy) in enumerate([(-0.9, 0.85), (0.9, 0.85), (-0.9, -0.85), (0.9, -0.85)]):
        v, t, uv = tq(cx, cy)
        objs.append(_obj(mat, v, t, uv))
    if sphere_n:
        c = (0.0, 0.0, -2.2)
        v, t = cube_sphere(sphere_n, c, 0.45)
        objs.append(_obj(2, v, t, sphere_uvs(v, c)))
    return {
        "settings": _settings([0.05, 0.05, 0.1], width, height, buckets),
        "camera": {"matrix": IDENTITY, "position": [0.0, 0.0, 0.0]},
        "lights": [{"intensity": 60, "position": [0.0, 0.0, 0.0]}, {"intensity": 30, "position": [1.5, 1.0, -1.0]}],
        "textures": [
            {"name": "red", "type": "albedo", "albedo": [0.9, 0.1, 0.1]},
            {"name": "wire", "type": "edges", "edge_color": [0.0, 0.0, 0.0], "inner_color": [0.9, 0.9, 0.2], "edge_width": 0.05},
            {"name": "check", "type": "checker", "color_A": [0.1, 0.1, 0.1], "color_B": [0.9, 0.9, 0.9], "square_size": 0.125},
            {"name": "bmp", "type": "bitmap", "file_path": bitmap_path},
        ],
        "materials": [{"type": "diffuse", "albedo": "red", "smooth_shading": False},
                      {"type": "diffuse", "albedo": "wire", "smooth_shading": False},
                      {"type": "diffuse", "albedo": "check", "smooth_shading": False},
                      {"type": "diffuse", "albedo": "bmp", "smooth_shading": False}],
        "objects": objs,
    }


def hw14_dragon_class(width=3840, height=2160, sphere_n=290, buckets=24):
    """Config 4 (sphere_n=290 -> 1 009 200 triangles) and config 5 (sphere_n=913 -> 10 002 828)."""
    return {
        "settings": _settings([0.0, 0.5, 0.0], width, height, buckets),
        "camera": {"matrix": IDENTITY, "position": [0.0, 0.3, 0.0]},
        "lights": [{"intensity": 300, "position": [3.0, 4.0, 0.0]}, {"intensity": 150, "position": [-3.0, 2.0, -1.0]}],
        "materials": [{"type": "diffuse", "albedo": [0.8, 0.2, 0.2], "smooth_shading": False},
                      {"type": "diffuse", "albedo": [0.6, 0.6, 0.6], "smooth_shading": False}],
        "objects": [
            _obj(0, *cube_sphere(sphere_n, (0.0, 0.0, -4.0), 1.0, displaced=True)),
            _obj(1, *quad((-8, -1.3, 2), (8, -1.3, 2), (8, -1.3, -12), (-8, -1.3, -12))),
        ],
    }


def synthetic_10m(width=1920, height=1080, sphere_n=913, buckets=24):
    """Config 5: same generator, N=913, centre (0,0,-3) so the animation.cpp orbit (radius 5.12 about (0,0,-3)) looks at it."""
    s = hw14_dragon_class(width, height, sphere_n, buckets)
    s["objects"][0] = _obj(0, *cube_sphere(sphere_n, (0.0, 0.0, -3.0), 1.0, displaced=True))
    s["objects"][1] = _obj(1, *quad((-8, -1.3, 5), (8, -1.3, 5), (8, -1.3, -11), (-8, -1.3, -11)))
    s["lights"] = [{"intensity": 300, "position": [3.0, 4.0, 1.0]}, {"intensity": 150, "position": [-3.0, 2.0, -4.0]}]
    return s


def degenerate_uv_scene(width=240, height=135, buckets=1):
    """Small regression scene with zero-area pole triangles on mirror + glass UV spheres."""
    s = hw11_room(width, height, 4, buckets)
    s["objects"][5] = _obj(3, *uv_sphere(8, 12, (0.9, -0.2, -5.2), 0.7))
    s["objects"][6] = _obj(4, *uv_sphere(8, 12, (-0.8, -1.2, -4.0), 0.8))
    return s


def many_meshes(width=128, height=72, count=70, buckets=24):
    """Edge case: more than 64 meshes (the per-ray visited-mesh bitmask of the CUDA core is off, every listing of a mesh in
    a top-level leaf is traversed again, like the reference does) -- `count` small tilted quads on a grid in front of a
    back wall, one mirror among them, two lights."""
    objs = [_obj(1, *quad((-4, -3, -7), (4, -3, -7), (4, 3, -7), (-4, 3, -7)))]
    cols = 10
    for k in range(count - 1):
        cx = -2.7 + 0.6 * (k % cols)
        cy = -1.8 + 0.55 * (k // cols)
        z = -3.0 - 0.25 * ((k * 7) % 9)
        dz = 0.05 * ((k % 5) - 2)
        objs.append(_obj(2 if k == 17 else 0, *quad((cx - 0.25, cy - 0.2, z + dz), (cx + 0.25, cy - 0.2, z - dz),
                                                     (cx + 0.25, cy + 0.2, z - dz), (cx - 0.25, cy + 0.2, z + dz))))
    return {
        "settings": _settings([0.05, 0.1, 0.2], width, height, buckets),
        "camera": {"matrix": IDENTITY, "position": [0.0, 0.0, 0.0]},
        "lights": [{"intensity": 60, "position": [0.0, 2.0, -1.0]}, {"intensity": 30, "position": [-2.0, -1.0, 0.0]}],
        "materials": [{"type": "diffuse", "albedo": [0.8, 0.5, 0.2], "smooth_shading": False},
                      {"type": "diffuse", "albedo": [0.6, 0.6, 0.6], "smooth_shading": False},
                      {"type": "reflective", "albedo": [0.9, 0.9, 0.9], "smooth_shading": False}],
        "objects": objs,
    }


def no_lights(width=64, height=36, buckets=1):
    """Edge case: an empty light list (diffuse surfaces come out (0,0,0), misses the background) on a 1-bucket grid."""
    s = hw07_scene0(width, height, 0, buckets)
    s["lights"] = []
    return s


def empty_scene(width=32, height=18, buckets=1):
    """Edge case: no objects at all -- every pixel is the background colour, no shadow rays."""
    s = hw07_scene0(width, height, 0, buckets)
    s["objects"] = []
    return s


def orbit_cameras(frames: int, radius: float = 5.12, center_z: float = -3.0):
    """Camera path of app/animation.cpp:24-38 with DEG_CHANGE = 360/frames, evaluated in binary32 like the
    reference (M_PIf for the orbit, 22/7 inside Camera::pan, Camera.cpp:10-12,39-48)."""
    f32 = np.float32
    cams = []
    deg = f32(0.0)
    dchange = f32(360.0) / f32(frames)
    pi_f = f32(math.pi)
    for _ in range(frames):
        rad = f32(deg * f32(pi_f / f32(180.0)))
        x = f32(f32(np.sin(rad, dtype=f32)) * f32(radius))
        z = f32(f32(f32(np.cos(rad, dtype=f32)) * f32(radius)) - f32(-center_z))
        dx = f32(x - f32(0))
        dz = f32(z + f32(-center_z))
        look = f32(f32(np.arctan2(dx, dz, dtype=f32)) * f32(f32(180.0) / pi_f))
        r = f32(look * f32(f32(22) / f32(f32(7) * f32(180.0))))
        c, s = f32(np.cos(r, dtype=f32)), f32(np.sin(r, dtype=f32))
        # IDENTITY *= rotateAroundY  (Camera.cpp:39-48, Matrix.h:144-157)
        rot = [c, 0.0, -s, 0.0, 1.0, 0.0, s, 0.0, c]
        cams.append(([float(x), 0.0, float(z)], [float(v) for v in rot]))
        deg = f32(deg + dchange)
    return cams


CONFIGS = {
    "hw07_scene0": lambda **kw: hw07_scene0(**kw),
    "hw07_scene0b": lambda **kw: hw07_scene0(sphere_n=kw.pop("sphere_n", 41), **kw),
    "hw11_room": lambda **kw: hw11_room(**kw),
    "hw12_textures": lambda **kw: hw12_textures(**kw),
    "hw14_dragon_class": lambda **kw: hw14_dragon_class(**kw),
    "synthetic_10M": lambda **kw: synthetic_10m(**kw),
    "degenerate_uv": lambda **kw: degenerate_uv_scene(**kw),
    "many_meshes": lambda **kw: many_meshes(**kw),
    "no_lights": lambda **kw: no_lights(**kw),
    "empty_scene": lambda **kw: empty_scene(**kw),
}


def build(name: str, out_dir: str, **kw) -> str:
    """Write `<out_dir>/<name>.crtscene` (and the bitmap for textured configs); return the path."""
    os.makedirs(out_dir, exist_ok=True)
    scene = CONFIGS[name](**kw)
    if "textures" in scene:
        for t in scene["textures"]:
            if t["type"] == "bitmap":
                # the reference concatenates folder + file_path with no separator (SceneParser.cpp:201)
                write_png_rgb(out_dir + t["file_path"], pattern_bitmap(512))
    return write_crtscene(os.path.join(out_dir, name + ".crtscene"), scene)
