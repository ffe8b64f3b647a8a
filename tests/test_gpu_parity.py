"""GPU parity tests proper: the CUDA path, called through the C ABI (libcrtb200.so), against the CPU oracle on the same
inputs and against the committed reference fixtures.  Bar: bit-exact hit ids (mesh, triangle, t), bit-identical float RGB,
identical ray counts and identical traversal work; the north_star 8-bit tolerance (|d| <= 1 on >= 99.9 % of pixels, none
> 4) follows and is asserted too."""
import os

import numpy as np
import pytest

from conftest import ROOT, SMALL_SCENES, same_f32

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def gpu(built):
    try:
        ctx = built.Context(0)
    except built.CrtError as e:  # fail loudly: the product has no CPU fallback
        pytest.fail(f"CUDA core unusable on a GPU box: {e}")
    yield ctx
    ctx.close()


def _covered(sf, rects, n):
    cov = np.zeros((sf.info.height, sf.info.width), bool)
    for i in range(n):
        r = rects[i]
        cov[r.row:r.row + r.height, r.col:r.col + r.width] = True
    return cov


def _assert_pixels(name, rgb, ref_rgb, q, ref_q):
    """Float RGB bit-identical (NaN == NaN) on every pixel -- refractive ones included, since the device evaluates the
    Fresnel powf exactly like glibc (csrc/crt_powf5.h) -- which implies the north_star 8-bit bar (|d| <= 1 on >= 99.9 %,
    none > 4), asserted separately so a float regression still reports how far the image moved."""
    d = np.abs(q.astype(np.int32) - ref_q.astype(np.int32)).max(axis=2)
    assert (d <= 1).mean() >= 0.999 and d.max() <= 4, f"{name}: 8-bit tolerance exceeded (max {d.max()})"
    same = same_f32(rgb, ref_rgb)
    assert same.all(), f"{name}: {(~same).sum()} float components differ (max |d| {np.nanmax(np.abs(rgb - ref_rgb))})"


@pytest.mark.parametrize("name", list(SMALL_SCENES))
def test_render_matches_oracle_and_golden(name, gpu, loaded, ob, crt):
    sf, flat, rects, n = loaded[name]
    gpu.upload(flat, keepalive=sf)
    opt = crt.make_options(rects=rects, n_rects=n, count_work=1)
    rgb, rgb8, hits, st = gpu.render(sf.camera(), opt, want_rgb8=True, want_hits=True)
    o_rgb, o_hits, o_st = ob.render(flat, sf.camera(), opt)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cov = _covered(sf, rects, n)
    # (i) hit ids: bit-exact, including t
    for ref_hits in (o_hits, g["hits"]):
        assert np.array_equal(hits["mesh"][cov], ref_hits["mesh"][cov])
        assert np.array_equal(hits["triangle"][cov], ref_hits["triangle"][cov])
        h = cov & (ref_hits["mesh"] >= 0)
        assert same_f32(hits["t"][h], ref_hits["t"][h]).all()
    # (ii) pixels
    _assert_pixels(name, rgb, o_rgb, rgb8, ob.quantize(o_rgb))
    _assert_pixels(name, rgb, g["rgb"], rgb8, g["ppm"])
    assert np.array_equal(rgb8, ob.quantize(rgb)), "device PPMColor quantiser differs from Color.cpp:12-16"
    # (iii) identical ray sets and, in visit-all counting mode, identical traversal work
    for k in ("rays_primary", "rays_shadow", "rays_reflection", "rays_refraction"):
        assert st[k] == o_st[k], (k, st[k], o_st[k])
    assert [st["rays_primary"], st["rays_shadow"], st["rays_reflection"], st["rays_refraction"]] == list(g["rays"])
    for k in ("node_tests_closest", "triangle_tests_closest", "node_tests_shadow", "triangle_tests_shadow"):
        assert st[k] == o_st[k], (k, st[k], o_st[k])


def test_uncovered_pixels_persist(gpu, loaded, crt):
    """colorBuffer persists across render() calls (RayTracer.h:69): uncovered pixels keep the previous frame."""
    sf, flat, rects, n = loaded["uncovered"]
    gpu.upload(flat, keepalive=sf)
    full, _, _, _ = gpu.render(sf.camera(), crt.make_options())
    part, _, _, _ = gpu.render(sf.camera(), crt.make_options(rects=rects, n_rects=n))
    assert same_f32(full, part).all()
    gpu.upload(flat, keepalive=sf)  # fresh buffer: uncovered pixels are (0,0,0), not background
    part2, _, _, _ = gpu.render(sf.camera(), crt.make_options(rects=rects, n_rects=n))
    cov = _covered(sf, rects, n)
    assert (~cov).any() and (part2[~cov] == 0).all()


def test_primary_rays_bit_exact(gpu, loaded, ob, crt, scenes_mod):
    sf, flat, _, _ = loaded["hw14_small"]
    gpu.upload(flat, keepalive=sf)
    for pos, rot in [((0.0, 0.3, 0.0), (1, 0, 0, 0, 1, 0, 0, 0, 1))] + scenes_mod.orbit_cameras(7)[1:4]:
        cam = crt.Camera.make(pos, rot)
        assert np.array_equal(gpu.generate_rays(cam).view(np.uint32), ob.generate_rays(flat, cam).view(np.uint32))


def test_trace_rays_random_and_nan(gpu, loaded, ob, crt):
    """RayTracer::trace / hasIntersection as plain queries, including NaN / axis-parallel / zero directions."""
    sf, flat, _, _ = loaded["hw11_room"]
    gpu.upload(flat, keepalive=sf)
    rng = np.random.default_rng(7)
    n = 20000
    rays = np.zeros((n, 6), np.float32)
    rays[:, 0:3] = rng.uniform(-1.9, 1.9, (n, 3)) + np.array([0, 0, -4.5])
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3:6] = d
    rays[0:50, 3] = 0.0          # axis-parallel components (|d| < FLT_EPSILON branch of the slab test)
    rays[50:100, 4] = 0.0
    rays[100:110, 3:6] = np.nan  # NaN rays pass every test
    rays[110:120, 3:6] = 0.0
    for rt in (crt.RAY_PRIMARY, crt.RAY_REFLECTION):
        a, b = gpu.trace_rays(rays, rt), ob.trace_rays(flat, rays, rt)
        assert np.array_equal(a["mesh"], b["mesh"]) and np.array_equal(a["triangle"], b["triangle"])
        assert same_f32(a["t"], b["t"]).all()
    dist = rng.uniform(0.1, 6.0, n).astype(np.float32)
    assert np.array_equal(gpu.trace_rays(rays, crt.RAY_SHADOW, dist), ob.trace_rays(flat, rays, crt.RAY_SHADOW, dist))


def test_sharded_render_assembles_to_same_frame(gpu, loaded, crt):
    """Tile sharding (multi-GPU partition) emulated on one GPU: 3 shards rendered separately == full frame."""
    torch = pytest.importorskip("torch")
    sf, flat, _, _ = loaded["hw11_room"]
    gpu.upload(flat, keepalive=sf)
    full, _, _, _ = gpu.render(sf.camera(), crt.make_options())
    world = 3
    items = gpu.shard_items(world)
    slabs = torch.zeros((world, items, 3), dtype=torch.float32, device="cuda")
    for r in range(world):
        gpu.render_device(sf.camera(), crt.make_options(shard_index=r, shard_count=world), d_rgb=slabs[r].data_ptr(),
                          stream=torch.cuda.current_stream().cuda_stream)
    out = torch.zeros((sf.info.height, sf.info.width, 3), dtype=torch.float32, device="cuda")
    gpu.assemble_shards(slabs.data_ptr(), world, d_rgb=out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), full.view(np.uint32))


def test_raytracer_mirror_and_ppm(gpu, loaded, ob, crt, tmp_path):
    """The C++ RayTracer mirror (render(path, options) + exportPPM) end to end: PPM byte-identical to the reference's."""
    sf, flat, rects, n = loaded["hw12_textures"]
    tracer = crt.RayTracer(sf)
    path = str(tmp_path / "out.ppm")
    rgb, st = tracer.render(path, mode=crt.MODE_B200_WAVEFRONT)
    g = np.load(os.path.join(GOLDEN, "hw12_textures.npz"))
    assert same_f32(rgb, g["rgb"]).all()
    assert np.array_equal(ob.read_ppm_p3(path), g["ppm"])
    tracer.close()


def test_chunked_frame_identical(gpu, loaded, crt):
    """A tiny queue budget forces the frame through several chunks; pixels and ray counts must not change."""
    sf, flat, _, _ = loaded["hw11_room"]
    gpu.upload(flat, keepalive=sf)
    a, _, _, sa = gpu.render(sf.camera(), crt.make_options())
    gpu.set_queue_budget(64 << 20)
    b, _, _, sb = gpu.render(sf.camera(), crt.make_options())
    gpu.set_queue_budget(16 << 30)
    assert same_f32(a, b).all()
    assert sa["rays_total"] == sb["rays_total"]


@pytest.mark.parametrize("name", [n for n in SMALL_SCENES if n != "degenerate_uv"])
def test_culled_mode_matches_exact_on_test_scenes(name, gpu, loaded, crt):
    """traversal = 1 (skip subtrees wholly behind the origin / beyond the best hit or the light) is NOT guaranteed to
    see the reference's candidate set; on every clean test scene it must nevertheless reproduce the
    exact mode bit for bit.  (The NaN-triangle scene is excluded: non-finite candidates are position-independent.)"""
    sf, flat, rects, n = loaded[name]
    gpu.upload(flat, keepalive=sf)
    a, _, ha, sa = gpu.render(sf.camera(), crt.make_options(rects=rects, n_rects=n, traversal=0), want_hits=True)
    b, _, hb, sb = gpu.render(sf.camera(), crt.make_options(rects=rects, n_rects=n, traversal=1, count_work=2), want_hits=True)
    assert same_f32(a, b).all()
    assert np.array_equal(ha["mesh"], hb["mesh"]) and np.array_equal(ha["triangle"], hb["triangle"])
    assert sa["rays_total"] == sb["rays_total"]
    with pytest.raises(crt.CrtError):  # visit-all counting is only defined for the exact walk
        gpu.render(sf.camera(), crt.make_options(traversal=1, count_work=1))


def test_dedup_does_less_work_than_visit_all_with_same_pixels(gpu, loaded, crt):
    """Production exact mode traverses a mesh once per ray even when several top-level leaves list it; count_work = 1
    reproduces the reference's repeated traversals (and its counters), count_work = 2 reports the real work."""
    sf, flat, rects, n = loaded["hw11_room"]
    gpu.upload(flat, keepalive=sf)
    a, _, _, visit_all = gpu.render(sf.camera(), crt.make_options(count_work=1))
    b, _, _, real = gpu.render(sf.camera(), crt.make_options(count_work=2))
    assert same_f32(a, b).all()
    assert real["node_tests"] < visit_all["node_tests"] and real["triangle_tests"] < visit_all["triangle_tests"]


@pytest.mark.parametrize("name", list(SMALL_SCENES))
def test_production_path_matches_golden(name, gpu, loaded, crt):
    """count_work = 0 is what bench.py times: any-hit shadow rays, one walk per mesh, MODE 2 loops (thresholded node
    phase + warp-cooperative triangle phase).  Same bar as the counting mode: bit-identical to the reference fixtures."""
    sf, flat, rects, n = loaded[name]
    gpu.upload(flat, keepalive=sf)
    rgb, rgb8, hits, st = gpu.render(sf.camera(), crt.make_options(rects=rects, n_rects=n), want_rgb8=True, want_hits=True)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cov = _covered(sf, rects, n)
    assert np.array_equal(hits["mesh"][cov], g["hits"]["mesh"][cov]) and np.array_equal(hits["triangle"][cov], g["hits"]["triangle"][cov])
    h = cov & (g["hits"]["mesh"] >= 0)
    assert same_f32(hits["t"][h], g["hits"]["t"][h]).all()
    _assert_pixels(name, rgb, g["rgb"], rgb8, g["ppm"])
    assert [st["rays_primary"], st["rays_shadow"], st["rays_reflection"], st["rays_refraction"]] == list(g["rays"])


@pytest.fixture(scope="module")
def gpu_wide(built):
    """A context that walks the opt-in 4-wide layout (CRT_LAYOUT is read by crtb200_create)."""
    os.environ["CRT_LAYOUT"] = "wide"
    try:
        ctx = built.Context(0)
    finally:
        del os.environ["CRT_LAYOUT"]
    yield ctx
    ctx.close()


@pytest.mark.parametrize("name", list(SMALL_SCENES))
def test_wide_layout_matches_golden(name, gpu_wide, loaded, crt):
    """The 4-wide collapse of the reference trees (crt_device.cuh "wide walk") enumerates the same candidates in the same
    order: hit ids, t and float RGB bit-identical to the reference fixtures, NaN-ray scene included."""
    sf, flat, rects, n = loaded[name]
    gpu_wide.upload(flat, keepalive=sf)
    for traversal in (0, 1):
        if traversal == 1 and name == "degenerate_uv":
            continue
        rgb, rgb8, hits, st = gpu_wide.render(sf.camera(), crt.make_options(rects=rects, n_rects=n, traversal=traversal),
                                              want_rgb8=True, want_hits=True)
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        cov = _covered(sf, rects, n)
        assert np.array_equal(hits["mesh"][cov], g["hits"]["mesh"][cov]) and np.array_equal(hits["triangle"][cov], g["hits"]["triangle"][cov])
        _assert_pixels(name, rgb, g["rgb"], rgb8, g["ppm"])
        assert [st["rays_primary"], st["rays_shadow"], st["rays_reflection"], st["rays_refraction"]] == list(g["rays"])


def test_wide_layout_falls_back_when_boxes_do_not_nest(gpu_wide, built, scene_dir, ob, crt):
    """The wide walk is only equivalent to the reference's when every child box lies inside its parent's.  A caller may
    upload any tree through the C ABI: shrink one inner box so its children stick out -> the library must keep walking
    the binary layout, i.e. still agree with the oracle run on the very same (odd) tree."""
    sf = built.SceneFile("hw14_small.crtscene", scene_dir)
    flat = sf.flatten()
    s = flat.contents
    big = max(range(s.n_meshes), key=lambda m: s.meshes[m].n_nodes)
    root = s.mesh_nodes[s.meshes[big].first_node]
    assert root.leaf_count == 0
    root.box_min[0] += 0.25 * (root.box_max[0] - root.box_min[0])  # children keep the old min.x: no longer nested
    gpu_wide.upload(flat, keepalive=sf)
    rgb, _, hits, st = gpu_wide.render(sf.camera(), crt.make_options(), want_hits=True)
    o_rgb, o_hits, o_st = ob.render(flat, sf.camera(), crt.make_options())
    assert np.array_equal(hits["mesh"], o_hits["mesh"]) and np.array_equal(hits["triangle"], o_hits["triangle"])
    assert same_f32(rgb, o_rgb).all()
    assert st["rays_total"] == o_st["rays_total"]


@pytest.fixture(scope="module")
def gpu_steal(built):
    """A context that runs the opt-in range-stealing kernels k_closest_s / k_shadow_s (CRT_STEAL is read by crtb200_create)."""
    os.environ["CRT_STEAL"] = "1"
    try:
        ctx = built.Context(0)
    finally:
        del os.environ["CRT_STEAL"]
    yield ctx
    ctx.close()


@pytest.mark.parametrize("name", list(SMALL_SCENES))
def test_range_stealing_matches_golden(name, gpu_steal, loaded, crt):
    """Long walks hand the far part of their node range to idle lanes of the warp (crt_kernels.cuh, k_*_s): helper
    candidates are folded back in encounter order, so hit ids, t and float RGB stay bit-identical to the reference."""
    sf, flat, rects, n = loaded[name]
    gpu_steal.upload(flat, keepalive=sf)
    rgb, rgb8, hits, st = gpu_steal.render(sf.camera(), crt.make_options(rects=rects, n_rects=n), want_rgb8=True, want_hits=True)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cov = _covered(sf, rects, n)
    assert np.array_equal(hits["mesh"][cov], g["hits"]["mesh"][cov]) and np.array_equal(hits["triangle"][cov], g["hits"]["triangle"][cov])
    h = cov & (g["hits"]["mesh"] >= 0)
    assert same_f32(hits["t"][h], g["hits"]["t"][h]).all()
    _assert_pixels(name, rgb, g["rgb"], rgb8, g["ppm"])
    assert [st["rays_primary"], st["rays_shadow"], st["rays_reflection"], st["rays_refraction"]] == list(g["rays"])


def test_range_stealing_full_size_equals_default(gpu, gpu_steal, built):
    """1920x1080 room with 196 608-triangle mirror and glass spheres (thousands of long walks, depth 5): the stealing
    kernels and the default kernels must agree bit for bit on hits and colours."""
    import importlib
    bench_mod = importlib.import_module("bench")
    f, folder, kw, tex, depth = bench_mod.ensure_scene("hw11_room_128", {})
    sf = built.SceneFile(f, folder)
    flat = sf.flatten()
    outs = []
    for ctx in (gpu, gpu_steal):
        ctx.upload(flat, keepalive=sf)
        rgb, _, hits, st = ctx.render(sf.camera(), built.make_options(max_depth=depth), want_hits=True)
        outs.append((rgb, hits, st))
    assert np.array_equal(outs[0][1]["triangle"], outs[1][1]["triangle"]) and np.array_equal(outs[0][1]["mesh"], outs[1][1]["mesh"])
    assert same_f32(outs[0][1]["t"], outs[1][1]["t"]).all()
    assert same_f32(outs[0][0], outs[1][0]).all()
    assert outs[0][2]["rays_total"] == outs[1][2]["rays_total"]


@pytest.fixture(scope="module")
def gpu_long(built):
    """A context that suspends every shadow walk after 4 node-phase iterations and finishes it in k_shadow_long, one warp
    per ray (CRT_LONG_BUDGET is read by crtb200_create; the default, 0, never suspends)."""
    os.environ["CRT_LONG_BUDGET"] = "4"
    try:
        ctx = built.Context(0)
    finally:
        del os.environ["CRT_LONG_BUDGET"]
    yield ctx
    ctx.close()


@pytest.mark.parametrize("name", list(SMALL_SCENES))
def test_long_walk_second_pass_matches_golden(name, gpu_long, loaded, crt):
    """The warp-per-ray walk covers the rest of a suspended shadow walk in a different order (LIFO of subtrees, 32 boxes
    per iteration); by the nesting property the set of tested leaves is the same, so every pixel stays bit-identical."""
    sf, flat, rects, n = loaded[name]
    gpu_long.upload(flat, keepalive=sf)
    for traversal in (0, 1):
        if traversal == 1 and name == "degenerate_uv":
            continue
        rgb, rgb8, hits, st = gpu_long.render(sf.camera(), crt.make_options(rects=rects, n_rects=n, traversal=traversal),
                                              want_rgb8=True, want_hits=True)
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        _assert_pixels(name, rgb, g["rgb"], rgb8, g["ppm"])
        assert [st["rays_primary"], st["rays_shadow"], st["rays_reflection"], st["rays_refraction"]] == list(g["rays"])
