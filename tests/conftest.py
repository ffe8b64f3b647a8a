import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "course-assignment-danielhalachev_b200"

# (scene name, builder kwargs, textured flavour, max depth): small variants of the five BASELINE.json configs plus
# the reference-quirk regressions (non-dividing bucket grid = uncovered pixels; degenerate pole triangles = NaN rays)
SMALL_SCENES = {
    "hw07_scene0": dict(builder="hw07_scene0", kw=dict(width=160, height=90, buckets=24), tex=False),
    "hw07_scene0b": dict(builder="hw07_scene0b", kw=dict(width=160, height=96, sphere_n=12, buckets=24), tex=False),
    "hw11_room": dict(builder="hw11_room", kw=dict(width=192, height=108, sphere_n=8, buckets=24), tex=False),
    "hw12_textures": dict(builder="hw12_textures", kw=dict(width=192, height=108, sphere_n=8, buckets=24), tex=True),
    "hw14_small": dict(builder="hw14_dragon_class", kw=dict(width=192, height=108, sphere_n=24, buckets=24), tex=False),
    "degenerate_uv": dict(builder="degenerate_uv", kw=dict(width=48, height=27, buckets=1), tex=False),
    "uncovered": dict(builder="hw11_room", kw=dict(width=100, height=70, sphere_n=4, buckets=24), tex=False),
    # > 64 meshes (no per-ray mesh de-duplication), a mirror among them; no lights at all; a 1x1 image; no objects
    "many_meshes": dict(builder="many_meshes", kw=dict(width=128, height=72, count=70, buckets=24), tex=False),
    "no_lights": dict(builder="no_lights", kw=dict(width=64, height=36, buckets=1), tex=False),
    "one_pixel": dict(builder="hw07_scene0", kw=dict(width=1, height=1, buckets=1), tex=False),
    "empty_scene": dict(builder="empty_scene", kw=dict(width=32, height=18, buckets=1), tex=False),
}


def same_f32(a, b):
    """Element-wise: bit-identical binary32, or both NaN (NaN sign / payload bits depend on x86 operand order and are
    not a property of the algorithm; every NaN quantises to 0, Color.cpp:12-16)."""
    import numpy as np
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def crt():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def scenes_mod():
    return importlib.import_module(PKG + ".scenes")


@pytest.fixture(scope="session")
def ob():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import binding
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def built(crt):
    """Native libraries are built in-tree (nvcc cross-compiles without a GPU)."""
    crt.build()
    return crt


@pytest.fixture(scope="session")
def scene_dir(tmp_path_factory, scenes_mod):
    d = str(tmp_path_factory.mktemp("scenes"))
    for name, spec in SMALL_SCENES.items():
        scene = scenes_mod.CONFIGS[spec["builder"]](**spec["kw"])
        if "textures" in scene:
            for t in scene["textures"]:
                if t["type"] == "bitmap":
                    scenes_mod.write_png_rgb(d + t["file_path"], scenes_mod.pattern_bitmap(64))
        scenes_mod.write_crtscene(os.path.join(d, name + ".crtscene"), scene)
    return d


@pytest.fixture(scope="session")
def loaded(built, scene_dir):
    """name -> (SceneFile, flattened scene pointer, rects, n_rects)"""
    out = {}
    for name in SMALL_SCENES:
        sf = built.SceneFile(name + ".crtscene", scene_dir)
        flat = sf.flatten()
        rects, n = sf.rects(mode=built.MODE_BVH_BUCKETS_THREADPOOL)
        out[name] = (sf, flat, rects, n)
    return out
