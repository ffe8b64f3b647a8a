"""GPU, BASELINE.json's full sizes.  Two kinds of checks:
  * direct: the frame against the UNMODIFIED reference binary (oracle/_ref, prebuilt, travels with the snapshot) run on
    the box's host cores on the same .crtscene -- float RGB bit-identical, PPM identical, ray counts identical;
  * size-independent properties: tile-shard invariance, chunking invariance (with a moving camera, so stale rows would
    show), default traversal == literal walk, quantiser idempotence, background / coverage accounting, frame-to-frame
    determinism, render_frames == render per camera.
"""
import importlib
import os

import numpy as np
import pytest

from conftest import PKG, same_f32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bench_mod():
    return importlib.import_module("bench")


@pytest.fixture(scope="module")
def gpu(built):
    ctx = built.Context(0)
    yield ctx
    ctx.close()


def _load(bench_mod, crt, name, **over):
    f, folder, kw, tex, depth = bench_mod.ensure_scene(name, over)
    sf = crt.SceneFile(f, folder)
    return sf, f, folder, tex, depth


@pytest.mark.parametrize("name", ["hw11_room", "hw12_textures", "hw07_scene0b"])
def test_1080p_frame_identical_to_reference_binary(name, gpu, built, bench_mod, ob, tmp_path):
    """Configs 1-3 at 1920x1080 against crt_ref / crt_ref_tex (a few seconds of CPU each)."""
    sf, f, folder, tex, depth = _load(bench_mod, built, name)
    if not ob.have_reference(tex):
        pytest.skip("oracle/_ref binaries not present")
    ref = ob.run_reference(f, folder, str(tmp_path / name), textured=tex, depth=depth, hits=True)
    gpu.upload(sf.flatten(), keepalive=sf)
    rects, n = sf.rects(mode=built.MODE_BVH_BUCKETS_THREADPOOL)
    rgb, rgb8, hits, st = gpu.render(sf.camera(), built.make_options(max_depth=depth, rects=rects, n_rects=n),
                                     want_rgb8=True, want_hits=True)
    assert np.array_equal(hits["mesh"], ref["hits"]["mesh"]) and np.array_equal(hits["triangle"], ref["hits"]["triangle"])
    h = ref["hits"]["mesh"] >= 0
    assert same_f32(hits["t"][h], ref["hits"]["t"][h]).all()
    same = same_f32(rgb, ref["rgb"])
    assert same.all(), f"{(~same).any(axis=2).sum()} of {same.shape[0] * same.shape[1]} pixels differ"
    assert np.array_equal(rgb8, ob.read_ppm_p3(ref["ppm_path"]))
    r = ref["rays"]
    assert (st["rays_primary"], st["rays_shadow"], st["rays_reflection"], st["rays_refraction"]) == \
           (r["primary"], r["shadow"], r["reflection"], r["refraction"])


def _compare_with_reference(gpu, built, ob, sf, f, folder, depth, prefix, camera=None):
    """Renders the scene file with the UNMODIFIED reference binary (host cores) and with the CUDA core; requires primary
    hit ids + t, float RGB, PPM bytes and the four ray counts to be identical."""
    ref = ob.run_reference(f, folder, prefix, textured=False, depth=depth, hits=True, camera=camera, timeout=1500.0)
    gpu.upload(sf.flatten(), keepalive=sf)
    cam = camera if camera is not None else sf.camera()
    rects, n = sf.rects(mode=built.MODE_BVH_BUCKETS_THREADPOOL)
    rgb, rgb8, hits, st = gpu.render(cam, built.make_options(max_depth=depth, rects=rects, n_rects=n), want_rgb8=True, want_hits=True)
    assert np.array_equal(hits["mesh"], ref["hits"]["mesh"]), f"{(hits['mesh'] != ref['hits']['mesh']).sum()} mesh ids differ"
    assert np.array_equal(hits["triangle"], ref["hits"]["triangle"]), f"{(hits['triangle'] != ref['hits']['triangle']).sum()} triangle ids differ"
    h = ref["hits"]["mesh"] >= 0
    assert same_f32(hits["t"][h], ref["hits"]["t"][h]).all()
    same = same_f32(rgb, ref["rgb"])
    assert same.all(), f"{(~same).any(axis=2).sum()} of {same.shape[0] * same.shape[1]} pixels differ"
    assert np.array_equal(rgb8, ob.read_ppm_p3(ref["ppm_path"]))
    r = ref["rays"]
    assert (st["rays_primary"], st["rays_shadow"], st["rays_reflection"], st["rays_refraction"]) == \
           (r["primary"], r["shadow"], r["reflection"], r["refraction"])
    # the literal walk (traversal 1) gives the same frame as the default (culling + hand-off)
    lit, _, _, _ = gpu.render(cam, built.make_options(max_depth=depth, rects=rects, n_rects=n, traversal=1))
    assert same_f32(lit, rgb).all()
    return ref, st


def test_4k_frame_identical_to_reference_binary(gpu, built, bench_mod, ob, tmp_path):
    """Config 4 at its stated size: hw14_dragon_class, 3840x2160, 1 009 202 triangles, against crt_ref
    (RayTracer::render, RayTracer.cpp:204-298): every pixel, every primary hit, every ray count."""
    if not ob.have_reference(False):
        pytest.skip("oracle/_ref binaries not present")
    sf, f, folder, tex, depth = _load(bench_mod, built, "hw14_dragon_class")
    ref, st = _compare_with_reference(gpu, built, ob, sf, f, folder, depth, str(tmp_path / "hw14_4k"))
    assert ref["width"] == 3840 and ref["height"] == 2160 and ref["triangles"] == 1009202


def test_10m_frame_identical_to_reference_binary(gpu, built, bench_mod, ob, scenes_mod, tmp_path):
    """Config 5 at its stated size: synthetic_10M (10 002 830 triangles), 1920x1080, one camera of the app/animation.cpp
    orbit (:24-38) that is not the identity, against crt_ref.  The reference needs ~2 min to parse and build its trees."""
    if not ob.have_reference(False):
        pytest.skip("oracle/_ref binaries not present")
    sf, f, folder, tex, depth = _load(bench_mod, built, "synthetic_10M")
    pos, rot = scenes_mod.orbit_cameras(60, radius=5.12, center_z=-3.0)[7]
    cam = built.Camera.make(pos, rot)
    assert [float(v) for v in cam.rotation] != [1, 0, 0, 0, 1, 0, 0, 0, 1]
    ref, st = _compare_with_reference(gpu, built, ob, sf, f, folder, depth, str(tmp_path / "syn10m"), camera=cam)
    assert ref["triangles"] == 10002830 and ref["width"] == 1920 and ref["height"] == 1080


def test_4k_1m_triangles_properties(gpu, built, bench_mod):
    """Config 4 (3840x2160, 1 009 202 triangles): properties that do not need the 10 s/frame CPU reference."""
    torch = pytest.importorskip("torch")
    sf, f, folder, tex, depth = _load(bench_mod, built, "hw14_dragon_class")
    W, H = sf.info.width, sf.info.height
    gpu.upload(sf.flatten(), keepalive=sf)
    cam = sf.camera()
    a, a8, ha, sa = gpu.render(cam, built.make_options(max_depth=depth), want_rgb8=True, want_hits=True)
    # accounting: every pixel is a primary ray; every hit on a diffuse mesh shoots one shadow ray per light
    assert sa["rays_primary"] == W * H
    assert sa["rays_shadow"] == int((ha["mesh"] >= 0).sum()) * 2
    bg = np.array(list(sf.info.background), np.float32)
    assert (a[ha["mesh"] < 0] == bg).all()
    assert np.isfinite(a).all() and (a >= 0).all()
    # determinism + chunking invariance + default traversal == literal walk
    b, _, _, sb = gpu.render(cam, built.make_options(max_depth=depth))
    assert same_f32(a, b).all() and sa["rays_total"] == sb["rays_total"]
    gpu.set_queue_budget(256 << 20)
    c, _, _, _ = gpu.render(cam, built.make_options(max_depth=depth))
    gpu.set_queue_budget(16 << 30)
    assert same_f32(a, c).all()
    d, _, hd, _ = gpu.render(cam, built.make_options(max_depth=depth, traversal=1), want_hits=True)
    assert same_f32(a, d).all() and np.array_equal(ha["triangle"], hd["triangle"])
    # tile-shard invariance (the multi-GPU partition emulated on one GPU): 8 shards assemble to the same frame
    world = 8
    items = gpu.shard_items(world)
    slabs = torch.zeros((world, items, 3), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for r in range(world):
        gpu.render_device(cam, built.make_options(max_depth=depth, shard_index=r, shard_count=world), d_rgb=slabs[r].data_ptr(), stream=st)
    out = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    out8 = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    gpu.assemble_shards(slabs.data_ptr(), world, d_rgb=out.data_ptr(), d_rgb8=out8.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert same_f32(out.cpu().numpy(), a).all()
    assert np.array_equal(out8.cpu().numpy(), a8)


def test_animation_frames_match_per_camera_renders_and_oracle(gpu, built, loaded, ob, scenes_mod):
    """app/animation.cpp:24-38 batched: crtb200_render_frames over orbit cameras == one render per camera == oracle."""
    sf, flat, _, _ = loaded["hw14_small"]
    gpu.upload(flat, keepalive=sf)
    cams = [built.Camera.make(p, r) for p, r in scenes_mod.orbit_cameras(6, radius=5.12, center_z=-4.0)]
    opt = built.make_options()
    frames, _, st = gpu.render_frames(cams, opt)
    total = 0
    for k, cam in enumerate(cams):
        one, _, _, s1 = gpu.render(cam, opt)
        assert same_f32(frames[k], one).all()
        total += s1["rays_total"]
        o_rgb, _, _ = ob.render(flat, cam, opt, want_hits=False)
        assert same_f32(one, o_rgb).all()
    assert st["rays_primary"] + st["rays_shadow"] + st["rays_reflection"] + st["rays_refraction"] == total
