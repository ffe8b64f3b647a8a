"""CPU: the C-ABI libraries load and export every symbol include/*.h declares; host-side logic (rectangle grids, camera,
PPM writer, JSON number parsing, error behaviour without a GPU)."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crt(?:b200|fe)_[a-z0-9_]+)\s*\(", text)))


def test_core_exports_every_declared_symbol(built):
    lib = built.core()
    names = _declared_functions("crtb200.h")
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"libcrtb200.so does not export {n}"
    assert lib.crtb200_abi_version() == 2


def test_front_exports_every_declared_symbol(built):
    lib = built.front()
    names = [n for n in _declared_functions("crtfront.h") if n.startswith("crtfe_")]
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"libcrtfront.so does not export {n}"


def test_no_cpu_fallback(built):
    """Without a usable sm_100 device the product fails loudly instead of rendering on the CPU."""
    n = C.c_int(-1)
    rc = built.core().crtb200_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(built.CrtError):
        built.Context(0)


def test_struct_layouts_match_header(built):
    # sizes the C compiler gives the ABI structs (x86-64): guards the ctypes mirrors in the package
    assert C.sizeof(built.KdNode) == 40 and C.sizeof(built.Mesh) == 36 and C.sizeof(built.Material) == 28
    assert C.sizeof(built.Texture) == 48 and C.sizeof(built.Light) == 16 and C.sizeof(built.Camera) == 48
    assert C.sizeof(built.Rect) == 16 and C.sizeof(built.Hit) == 12 and C.sizeof(built.Options) == 56
    assert C.sizeof(built.Stats) == 144 and C.sizeof(built.Scene) == 224


def test_rectangle_grid_matches_reference_arithmetic(built):
    """RayTracer.cpp:143-152: ny = floor(sqrt(n)), nx = n / ny, w = W / nx, h = H / ny, col = (i*w) % W, row = (i/nx)*h."""
    for (W, H, n) in [(1920, 1080, 24), (3840, 2160, 24), (160, 90, 24), (100, 70, 24), (640, 360, 24), (64, 64, 1), (97, 31, 7)]:
        rects, cnt = built.rectangles(W, H, n, mode=8)
        ny = int(math.sqrt(n)) or 1
        nx = n // ny
        w, h = W // nx, H // ny
        exp = []
        for i in range(n):
            col, row = (i * w) % W, (i // nx) * h
            rl, cl = min(H, row + h), min(W, col + w)
            if row < rl and col < cl:
                exp.append((row, col, cl - col, rl - row))
        got = [(rects[i].row, rects[i].col, rects[i].width, rects[i].height) for i in range(cnt)]
        assert got == exp
    # single-rectangle modes (NoOptimization / AABB / BVH): the whole image
    rects, cnt = built.rectangles(640, 360, 24, mode=7)
    assert cnt == 1 and (rects[0].width, rects[0].height) == (640, 360)


def test_camera_pan_uses_22_over_7(built):
    """Camera::pan (Camera.cpp:10-12,39-48): degrees * (22 / (7 * 180.0f)), IDENTITY *= rotY."""
    cam = built.Camera.make()
    built.camera_pan(cam, 30.0)
    r = np.float32(30.0) * (np.float32(22) / (np.float32(7) * np.float32(180.0)))
    c, s = np.cos(r, dtype=np.float32), np.sin(r, dtype=np.float32)
    exp = np.array([c, 0, -s, 0, 1, 0, s, 0, c], np.float32)
    assert np.array_equal(np.array(list(cam.rotation), np.float32), exp)


def test_orbit_cameras_match_front_end_pan(built, scenes_mod):
    for pos, rot in scenes_mod.orbit_cameras(12):
        cam = built.Camera.make()
        x, z = np.float32(pos[0]), np.float32(pos[2])
        look = np.float32(np.arctan2(x, np.float32(z + np.float32(3.0)), dtype=np.float32) * np.float32(np.float32(180.0) / np.float32(math.pi)))
        built.camera_pan(cam, float(look))
        assert np.allclose(np.array(list(cam.rotation)), np.array(rot), atol=2e-7)


def test_ppm_writer_is_byte_identical_to_reference_format(built, ob, tmp_path):
    rng = np.random.default_rng(3)
    rgb = rng.uniform(-0.2, 1.3, (7, 5, 3)).astype(np.float32)
    rgb[0, 0] = [np.nan, 1.0, 0.999999]
    rgb[1, 1] = [np.inf, -np.inf, 0.0]
    path = str(tmp_path / "o.ppm")
    built.write_ppm(path, rgb)
    q = ob.quantize(rgb)
    exp = "P3\n5 7\n255\n" + "".join("".join("%d %d %d\t" % tuple(q[r, c]) for c in range(5)) + "\n" for r in range(7))
    assert open(path).read() == exp


def test_scene_loader_numbers_and_errors(built, tmp_path):
    scene = ('{"settings":{"background_color":[0.1,0.25,1e-1],"image_settings":{"width":8,"height":4,"bucket_size":2}},'
             '"camera":{"matrix":[1,0,0,0,1,0,0,0,1],"position":[0,0.30000001192092896,-1.5e0]},"lights":[{"intensity":7,"position":[1,2,3]}],'
             '"materials":[{"type":"diffuse","albedo":[0.1,0.2,0.3],"smooth_shading":false}],'
             '"objects":[{"material_index":0,"vertices":[0,0,-1, 1,0,-1, 0,1,-1],"triangles":[0,1,2]}]}')
    (tmp_path / "s.crtscene").write_text(scene)
    sf = built.SceneFile("s.crtscene", str(tmp_path))
    assert (sf.info.width, sf.info.height, sf.info.bucket_size, sf.info.n_triangles) == (8, 4, 2, 1)
    assert np.array_equal(np.array(list(sf.info.background), np.float32), np.array([0.1, 0.25, 0.1], np.float32))
    assert np.float32(sf.info.camera.position[1]) == np.float32(0.3)
    s = sf.flatten().contents
    assert [s.triangle_normal[i] for i in range(3)] == [0.0, 0.0, 1.0]
    with pytest.raises(built.CrtError):
        built.SceneFile("missing.crtscene", str(tmp_path))
    (tmp_path / "bad.crtscene").write_text('{"settings":{}}')
    with pytest.raises(built.CrtError):
        built.SceneFile("bad.crtscene", str(tmp_path))
