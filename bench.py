#!/usr/bin/env python3
"""bench.py -- Mrays/s and ms/frame of the per-pixel hot path on N B200s (contract: see README / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one frame of the named workload: every primary, shadow, reflection and refraction ray the reference would
trace for it (SURVEY.md 8(d) definition).  Default workload = BASELINE.json config 4, `hw14_dragon_class`: 1 009 200-triangle
displaced cube-sphere + ground quad, 3840x2160, 2 point lights -- "the largest KD-tree scene" the north_star's target
is quoted on that fits one GPU.

  value      whole-job Mrays/s with the scene resident in HBM and the framebuffer left on the device
             (crtb200_render_device); per-step CUDA events on the launching stream, L2 flushed between steps
  e2e        the same metric through crtb200_render with HOST buffers: camera + options in, the float colour buffer
             (what RayTracer::render returns) copied back to pinned host memory inside the timed region
  roofline   dominant traversal launch group (k_shadow + its k_coop, or k_closest): the DRAM bytes it really moves
             (ncu dram__bytes_read + write of the same sources, profiles/ncu_counters.json) / its CUDA-event time in
             THIS run, vs the measured HBM copy peak; L2 / L1 hit rates, issue-slot use and warp execution efficiency
             from the same capture; the SURVEY 8(d) byte model (visit-all and executed) under `byte_model`
  config5    synthetic_10M (10 M triangles) on the same GPU: a static 1080p frame and the 60-frame orbit
  cpu_baseline  the UNMODIFIED reference (oracle/_ref/crt_ref) on this box's host cores, one full frame, its pixels and
             ray counts compared with ours; --impl reference: steps + warmup full frames of the same binary
  literal_walk  the reference's visit-all itinerary (traversal = 1) beside the default, with the pixel difference (0)

N > 1 (torchrun), scene replicated per GPU:
  tiles (default)                 one frame, 8x4-pixel tiles dealt round-robin.  Up to 4 GPUs every rank's store kernel writes
                                  its tiles straight into rank 0's frame (peer memory over NVLink, CUDA IPC; one tiny NCCL
                                  all-reduce as completion barrier); beyond, compact float slabs are gathered to rank 0 over
                                  NCCL and scattered into the frame (--tile-transport peer | gather forces one).  Strong
                                  scaling: total work fixed.  Both frames are checked bit-equal to the single-GPU frame before
                                  timing (tiles_check); the transport not used is timed as `tiles_other_transport`, the frames
                                  partition as `frames_mode`.  e2e: every rank's store kernel writes its tiles into ONE pinned
                                  host frame shared by the ranks (zero-copy over each GPU's own PCIe link).
  --parallelism frames            every rank renders one frame of the sequence per step, PPMColor bytes gathered to rank 0
                                  over NCCL overlapped with the next render.  Weak scaling: per-GPU work fixed.
  --animation F                   config 5's orbit: frame f on rank f % N, frames gathered to rank 0.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "course-assignment-danielhalachev_b200"

WORKLOADS = {
    # name: (builder, kwargs, textured, max_depth)
    "hw14_dragon_class": ("hw14_dragon_class", dict(width=3840, height=2160, sphere_n=290), False, 5),
    "hw07_scene0b": ("hw07_scene0b", dict(width=1920, height=1080, sphere_n=41), False, 5),
    "hw11_room": ("hw11_room", dict(width=1920, height=1080, sphere_n=32), False, 5),
    "hw11_room_128": ("hw11_room", dict(width=1920, height=1080, sphere_n=128), False, 5),
    "hw12_textures": ("hw12_textures", dict(width=1920, height=1080, sphere_n=92), True, 5),
    "synthetic_10M": ("synthetic_10M", dict(width=1920, height=1080, sphere_n=913), False, 5),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def scene_cache_dir() -> str:
    d = os.environ.get("CRT_BENCH_CACHE", "/tmp/crtb200_bench")
    os.makedirs(d, exist_ok=True)
    return d


def ensure_scene(name: str, overrides: dict) -> tuple[str, str, dict, bool, int]:
    """Writes the workload's .crtscene (deterministic generator) once; returns (file, folder, kwargs, textured, depth)."""
    scenes = importlib.import_module(PKG + ".scenes")
    builder, kw, tex, depth = WORKLOADS[name]
    kw = dict(kw)
    kw.update({k: v for k, v in overrides.items() if v})
    tag = name + "_" + "_".join(f"{k}{v}" for k, v in sorted(kw.items()))
    folder = scene_cache_dir()
    path = os.path.join(folder, tag + ".crtscene")
    if not os.path.exists(path):
        t = time.time()
        scene = scenes.CONFIGS[builder](**kw)
        if "textures" in scene:
            for tx in scene["textures"]:
                if tx["type"] == "bitmap":
                    scenes.write_png_rgb(folder + tx["file_path"], scenes.pattern_bitmap(512))
        tmp = path + f".tmp{os.getpid()}"
        scenes.write_crtscene(tmp, scene)
        os.replace(tmp, path)
        log(f"[bench] generated {path} in {time.time() - t:.1f}s")
    return tag + ".crtscene", folder, kw, tex, depth


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.rows = []
        self.proc = None
        self.dev = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(rays: int, node_tests: int, tri_tests: int) -> int:
    """bytes(ray) = 32*n_node + 52*n_tri + 64."""
    return 32 * node_tests + 52 * tri_tests + 64 * rays


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref) on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def run_reference_full(scene_file, folder, tex, depth, repeats, out_prefix="-"):
    """crt_ref on the FULL frame of the workload's .crtscene, `repeats` renders (tree build untimed, like MEASURE_TIME)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import binding as ob
    if not ob.have_reference(tex):
        return None
    out = ob.run_reference(scene_file, folder, out_prefix, textured=tex, depth=depth, hits=False, ppm=False, repeat=repeats, timeout=3300.0)
    out["rays_total"] = sum(out["rays"].values())
    out["sample"] = (f"{repeats} full {out['width']}x{out['height']} frame(s) of the workload ({out['triangles']} triangles), oracle/_ref/crt_ref "
                     f"(the unmodified reference), mode BVHBucketsThreadPool on {out['threads']} host threads, MEASURE_TIME seconds; "
                     f"KD build {out['build_s']:.1f}s untimed")
    return out


def reference_arm(args):
    """The reference's own CPU path on the SAME config as our arm: every step is one full frame of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    scene_file, folder, kw, tex, depth = ensure_scene(args.workload, dict(width=args.width, height=args.height))
    # bounded run: full frames only (never a smaller frame); if K + W frames cannot fit the budget, fewer frames are
    # timed and the line says how many
    repeats = args.steps + args.warmup
    budget = float(os.environ.get("CRT_REFERENCE_BUDGET_S", "1200"))
    cal = run_reference_full(scene_file, folder, tex, depth, 1)
    if cal is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/crt_ref not built"}))
        return 0
    per_frame = cal["render_s"]
    timed_repeats = max(1, min(repeats, int(budget / max(per_frame, 1e-3))))
    out = run_reference_full(scene_file, folder, tex, depth, timed_repeats) if timed_repeats > 1 else cal
    warm = min(args.warmup, max(0, timed_repeats - 1))
    times = out["render_all_s"][warm:]
    sec = sum(times) / len(times)
    v = out["rays_total"] / sec / 1e6
    sample = out["sample"] + f"; timed {len(times)} step(s) after {warm} warm-up"
    line = {
        "impl": "reference", "metric": "Mrays/s (primary+secondary)", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "width": out["width"], "height": out["height"], "triangles": out["triangles"],
                   "max_depth": depth, "rays_per_step": out["rays_total"], "sample": sample},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": out["threads"], "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rays_per_step": out["rays_total"], "steps_timed": len(times),
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def source_sha16() -> str:
    """Hash of the CUDA sources: ties the ncu counters in profiles/ncu_counters.json to the binary that is timed."""
    import hashlib
    h = hashlib.sha256()
    for f in ("crt_kernels.cuh", "crt_device.cuh", "crtb200_core.cu", "crt_powf5.h"):
        h.update(open(os.path.join(ROOT, PKG, "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def ncu_counters(workload: str, kernel: str):
    """ncu --set full counters of `kernel` on `workload`, captured from this very source (tools/ncu_counters.py writes
    the file from the .ncu-rep); None when the capture is of other sources."""
    p = os.path.join(ROOT, "profiles", "ncu_counters.json")
    try:
        d = json.load(open(p))
    except Exception:
        return None
    if d.get("source_sha16") != source_sha16():
        return None
    return d.get(workload, {}).get(kernel)


def roofline_record(workload, cst, cst2, seq, ms_per_step, W, H):
    """Dominant traversal launch group (K2 = k_closest + its k_coop, or K3 = k_shadow + its k_coop) against the HBM peak.
    achieved = DRAM bytes the group really moves (ncu dram__bytes_read + write of the same sources, per frame) / its
    CUDA-event time measured in THIS run; the SURVEY 8(d) byte model (visit-all and executed) is reported beside it."""
    peak, peak_src = measured_hbm_peak()
    c_ms = statistics.mean(s["closest_ms"] for s in seq)
    s_ms = statistics.mean(s["shadow_ms"] for s in seq)
    closest_rays = cst["rays_primary"] + cst["rays_reflection"] + cst["rays_refraction"]
    b_closest = algorithmic_bytes(closest_rays, cst["node_tests_closest"], cst["triangle_tests_closest"])
    b_shadow = algorithmic_bytes(cst["rays_shadow"], cst["node_tests_shadow"], cst["triangle_tests_shadow"])
    a_closest = algorithmic_bytes(closest_rays, cst2["node_tests_closest"], cst2["triangle_tests_closest"])
    a_shadow = algorithmic_bytes(cst2["rays_shadow"], cst2["node_tests_shadow"], cst2["triangle_tests_shadow"])
    dom = ("k_closest", b_closest, c_ms, a_closest) if c_ms >= s_ms else ("k_shadow", b_shadow, s_ms, a_shadow)
    ncu = ncu_counters(workload, dom[0])
    visit_all_gbs = dom[1] / (dom[2] * 1e-3) / 1e9
    executed_gbs = dom[3] / (dom[2] * 1e-3) / 1e9
    rec = {"bound": "hbm", "kernel": dom[0], "peak": peak, "unit": "GB/s", "peak_source": peak_src, "kernel_ms": dom[2],
           "closest_ms": c_ms, "shadow_ms": s_ms, "coop_closest_ms": statistics.mean(s["coop_closest_ms"] for s in seq),
           "coop_shadow_ms": statistics.mean(s["coop_shadow_ms"] for s in seq),
           "sequential_frame_ms": statistics.mean(s["device_ms"] for s in seq),
           "byte_model": {"visit_all_bytes_per_launch": dom[1], "visit_all_gbs": visit_all_gbs, "visit_all_frac": visit_all_gbs / peak,
                          "executed_bytes_per_launch": dom[3], "executed_gbs": executed_gbs, "executed_frac": executed_gbs / peak,
                          "note": "SURVEY 8(d): 32 B/AABB test + 52 B/triangle test + 64 B/ray.  visit_all = the reference's visit-all work "
                                  "(count_work=1); executed = the tests this kernel really performs (any-hit shadow rays, one walk per mesh, "
                                  "conservative culling: all exact).  These are L1/L2-level logical bytes, not HBM traffic."},
           "node_tests_per_ray": cst["node_tests"] / max(1, cst["rays_total"]),
           "triangle_tests_per_ray": cst["triangle_tests"] / max(1, cst["rays_total"]),
           "executed_node_tests_per_ray": cst2["node_tests"] / max(1, cst2["rays_total"]),
           "frame_GBps_all_kernels_executed": (a_closest + a_shadow + 12 * W * H) / (ms_per_step * 1e-3) / 1e9}
    if ncu:
        achieved = ncu["dram_bytes"] / (dom[2] * 1e-3) / 1e9
        rec.update({"achieved": achieved, "frac": achieved / peak, "traffic": ncu["dram_bytes"],
                    "limiter": "L2 latency / issue rate of a dependent pointer chase (not HBM): see l2_hit, issue_slot_util, warp_exec_eff",
                    "l2_hit": ncu.get("l2_hit"), "l1_hit": ncu.get("l1_hit"), "issue_slot_util": ncu.get("issue_slot_util"),
                    "warp_exec_eff": ncu.get("warp_exec_eff"), "achieved_occupancy": ncu.get("achieved_occupancy"),
                    "ncu_source_sha16": source_sha16(), "ncu_kernel_ms": ncu.get("kernel_ms"), "ncu_file": ncu.get("file")})
    else:
        rec.update({"achieved": executed_gbs, "frac": executed_gbs / peak, "traffic": None,
                    "limiter": "no ncu capture of these exact sources is committed: achieved falls back to the executed byte model (logical bytes)"})
    return rec


def config5_record(crt, torch, args, dev):
    """BASELINE.json config 5 on this GPU: synthetic_10M (10 002 830 triangles, the HBM-streaming regime), 1920x1080 --
    a static frame, and the 60-frame orbit of app/animation.cpp:24-38 through the batched C-ABI call."""
    scenes = importlib.import_module(PKG + ".scenes")
    t0 = time.time()
    scene_file, folder, kw, tex, depth = ensure_scene("synthetic_10M", {})
    sf = crt.SceneFile(scene_file, folder)
    flat = sf.flatten()
    ctx = crt.Context(dev.index)
    ctx.upload(flat, keepalive=sf)
    t_setup = time.time() - t0
    W, H = sf.info.width, sf.info.height
    stream = torch.cuda.current_stream().cuda_stream
    frame = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    opt = crt.make_options(max_depth=depth, traversal=args.traversal)
    ms = []
    for k in range(3 + 5):
        flush.fill_(k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ctx.render_device(sf.camera(), opt, d_rgb=frame.data_ptr(), stream=stream)
        b.record()
        torch.cuda.synchronize()
        if k >= 3:
            ms.append(a.elapsed_time(b))
    st = ctx.last_stats()
    static = {"ms_per_frame": statistics.mean(ms), "value": st["rays_total"] / (statistics.mean(ms) * 1e-3) / 1e6, "unit": "Mrays/s",
              "rays_per_frame": st["rays_total"], "frames_timed": len(ms)}
    F = 60
    cams = [crt.Camera.make(p, r) for p, r in scenes.orbit_cameras(F, radius=5.12, center_z=-3.0)]
    rays = 0
    for cam in cams:  # untimed: ray count of the sequence (also the warm-up)
        ctx.render_device(cam, opt, d_rgb=frame.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        rays += ctx.last_stats()["rays_total"]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for cam in cams:
        ctx.render_device(cam, opt, d_rgb=frame.data_ptr(), stream=stream)
    b.record()
    torch.cuda.synchronize()
    seq_ms = a.elapsed_time(b)
    # e2e: crtb200_render_frames, PPMColor frames into pinned host memory (what exportPPM consumes), copies inside
    host8 = torch.empty((F, H, W, 3), dtype=torch.uint8).pin_memory()
    t = time.perf_counter()
    ctx.render_frames_into(cams, opt, rgb8_out=host8.numpy())
    e2e_ms = (time.perf_counter() - t) * 1e3
    orbit = {"frames": F, "ms_per_sequence": seq_ms, "ms_per_frame": seq_ms / F, "value": rays / (seq_ms * 1e-3) / 1e6, "unit": "Mrays/s",
             "rays_per_sequence": rays,
             "e2e": {"value": rays / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_sequence": e2e_ms, "ms_per_frame": e2e_ms / F,
                     "d2h_bytes_per_sequence": F * W * H * 3, "api": "crtb200_render_frames(host cameras -> pinned host PPMColor frames)"}}
    ctx.close()
    return {"workload": "synthetic_10M", "width": W, "height": H, "triangles": int(sf.info.n_triangles), "max_depth": depth,
            "setup_s": t_setup, "static_frame": static, "orbit_60": orbit,
            "note": "scene working set (~0.6 GB of nodes / references / triangles) exceeds the 126 MB L2: the HBM-streaming regime"}


def ours_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    crt = importlib.import_module(PKG)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local_rank if world > 1 else 0)

    if rank == 0:
        scene_file, folder, kw, tex, depth = ensure_scene(args.workload, dict(width=args.width, height=args.height))
    if world > 1:
        dist.barrier()
    if rank != 0:
        scene_file, folder, kw, tex, depth = ensure_scene(args.workload, dict(width=args.width, height=args.height))

    t0 = time.time()
    sf = crt.SceneFile(scene_file, folder)
    t_parse = time.time() - t0
    flat = sf.flatten()
    ctx = crt.Context(dev.index)
    t1 = time.time()
    ctx.upload(flat, keepalive=sf)
    t_upload = time.time() - t1
    W, H = sf.info.width, sf.info.height
    cam = sf.camera()
    if rank == 0:
        log(f"[bench] {args.workload}: {sf.info.n_triangles} triangles, {W}x{H}; parse {t_parse:.2f}s, KD build+flatten "
            f"{sf.build_seconds:.2f}s, upload {t_upload:.2f}s")

    rects, n_rects = sf.rects(mode=crt.MODE_B200_WAVEFRONT)
    stream = torch.cuda.current_stream().cuda_stream
    frame = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    frame8 = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # one counting pass (untimed): ray counts and visit-all traversal work = the algorithmic bytes; and a few frames
    # with the chunks strictly sequential (concurrency 1) so the per-kernel CUDA-event timers are not inflated by overlap
    cst = cst2 = None
    seq = []
    if world == 1:
        _, _, _, cst = ctx.render(cam, crt.make_options(max_depth=depth, rects=rects, n_rects=n_rects, count_work=1), want_rgb=False)
        # and the tests the kernels really execute (any-hit shadow rays, one walk per mesh, conservative culling: all exact)
        _, _, _, cst2 = ctx.render(cam, crt.make_options(max_depth=depth, rects=rects, n_rects=n_rects, count_work=2, traversal=args.traversal), want_rgb=False)
        ctx.set_concurrency(1)
        for k in range(4):
            flush.fill_(k)
            torch.cuda.synchronize()
            _, _, _, s1 = ctx.render(cam, crt.make_options(max_depth=depth, rects=rects, n_rects=n_rects, traversal=args.traversal), want_rgb=False)
            if k:
                seq.append(s1)
    ctx.set_concurrency(args.concurrency)

    parallelism = args.parallelism
    if parallelism == "auto":
        parallelism = "tiles"  # BASELINE.json config 4: "tile-sharded over 1/2/4/8 GPUs" -- one frame, strong scaling
    opt = crt.make_options(max_depth=depth, rects=rects, n_rects=n_rects, traversal=args.traversal)
    sharded = None
    peer = None
    tiles_check = None
    single_frame_host = None
    mg = None
    if world > 1:
        mg = importlib.import_module(PKG + ".multigpu")
        sharded = mg.ShardedRenderer(crt, ctx, torch, dist, dev)  # NCCL gather of slabs + assembly on rank 0
        peer = mg.PeerStoreRenderer(crt, ctx, torch, dist, dev, W, H)  # every rank stores into rank 0's frame (CUDA IPC)
        if rank == 0:
            pframe, pframe8 = peer.frame_tensors()
        # both assembled N-GPU frames must be bit-equal to the frame one GPU renders alone (rank 0 renders all three)
        sharded.render(cam, max_depth=depth, traversal=args.traversal, frame=frame if rank == 0 else None, frame8=frame8 if rank == 0 else None)
        peer.render(cam, max_depth=depth, traversal=args.traversal)
        torch.cuda.synchronize()
        if rank == 0:
            single = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
            single8 = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)
            ctx.render_device(cam, opt, d_rgb=single.data_ptr(), d_rgb8=single8.data_ptr(), stream=stream)
            torch.cuda.synchronize()

            def differing(a):
                return int(((a.view(torch.int32) != single.view(torch.int32)) & ~(torch.isnan(a) & torch.isnan(single))).any(dim=2).sum().item())
            tiles_check = {"pixels": W * H,
                           "peer_store": {"pixels_differing_from_single_gpu_frame": differing(pframe), "rgb8_equal": bool(torch.equal(pframe8, single8))},
                           "nccl_gather": {"pixels_differing_from_single_gpu_frame": differing(frame), "rgb8_equal": bool(torch.equal(frame8, single8))}}
            single_frame_host = single.cpu().numpy()  # (kept for the check of the zero-copy host frame of the e2e leg)
            del single, single8
            for k in ("peer_store", "nccl_gather"):
                if tiles_check[k]["pixels_differing_from_single_gpu_frame"] or not tiles_check[k]["rgb8_equal"]:
                    raise SystemExit(f"[bench] tile-sharded frame ({k}) differs from the single-GPU frame: {tiles_check}")
        dist.barrier()

    # frames partition (weak scaling, kept as an extra key at N > 1): every rank renders one frame of a static-camera
    # sequence per step, PPMColor frames gathered to rank 0 with one NCCL gather, double-buffered
    bufs8 = gflat = gls = None
    works = [None, None]
    if world > 1:
        bufs8 = [torch.zeros((H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
        gflat = [torch.zeros((world, H, W, 3), dtype=torch.uint8, device=dev) if rank == 0 else None for _ in range(2)]
        gls = [[gflat[b][i] for i in range(world)] for b in range(2)] if rank == 0 else [None, None]
    host_frame = torch.empty((H, W, 3), dtype=torch.float32).pin_memory() if (world > 1 and rank == 0) else None

    # auto: peer stores up to 4 GPUs, the NCCL gather beyond (run r2fs / r2bc, ms per 4K frame peer / gather: N = 2
    # 1.77 / 1.81, N = 4 1.28 / 1.29, N = 8 1.03-1.05 / 0.95-1.02: eight writers into one GPU's frame lose to NCCL's transfer)
    use_peer = args.tile_transport == "peer" or (args.tile_transport == "auto" and world <= 4)

    def step_tiles_gather(to_host=False):
        sharded.render(cam, max_depth=depth, traversal=args.traversal, frame=frame if rank == 0 else None,
                       frame8=frame8 if rank == 0 else None)
        if to_host and rank == 0:
            host_frame.copy_(frame, non_blocking=True)

    def step_tiles_peer(to_host=False):
        peer.render(cam, max_depth=depth, traversal=args.traversal)
        if to_host and rank == 0:
            host_frame.copy_(pframe, non_blocking=True)

    step_tiles = step_tiles_peer if use_peer else step_tiles_gather

    def step_frames(k):
        b = k & 1
        if works[b] is not None:
            works[b].wait()
        ctx.render_device(cam, opt, d_rgb=frame.data_ptr(), d_rgb8=bufs8[b].data_ptr(), stream=stream)
        works[b] = dist.gather(bufs8[b], gls[b], dst=0, async_op=True)

    def drain_frames():
        for b in range(2):
            if works[b] is not None:
                works[b].wait()
                works[b] = None

    def step_single():
        ctx.render_device(cam, opt, d_rgb=frame.data_ptr(), d_rgb8=frame8.data_ptr(), stream=stream)

    tiles_mode = world > 1 and parallelism == "tiles"
    frames_mode = world > 1 and parallelism == "frames"

    def timed_per_step(step_fn, steps):
        """K steps bracketed by a barrier + synchronize on both sides; per step: L2 flush (outside the events), event pair
        around the step.  Everything is queued asynchronously: with N > 1 the ranks re-align on the device at the end of
        every step (the step's own collective), so a step's events contain no host-side barrier.  Returns per-step ms of
        this rank and the last step's stats."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        for k in range(steps):
            flush.fill_(k & 0xFF)  # L2 flush (256 MiB > 126 MB L2), outside the timed events
            ev[k][0].record()
            step_fn()
            ev[k][1].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return [a.elapsed_time(b) for a, b in ev], [ctx.last_stats()]

    def timed_frames(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        a.record()
        for k in range(steps):
            step_frames(k)
        drain_frames()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    main_step = step_tiles if tiles_mode else (step_single if world == 1 else None)
    for k in range(max(args.warmup, 3)):
        if frames_mode:
            step_frames(k)
        else:
            main_step()
    if frames_mode:
        drain_frames()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    if frames_mode:
        per = timed_frames(args.steps)
        step_ms = [per] * args.steps
        kstats = [ctx.last_stats()]
    else:
        step_ms, kstats = timed_per_step(main_step, args.steps)
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(sum(step_ms))
    ms_per_step = total_ms / args.steps
    # ray count of one step: identical every step; tiles: sum of the shards; frames: the N frames of a step
    rays_step = sum_over_ranks(kstats[-1]["rays_total"])
    value = rays_step / (ms_per_step * 1e-3) / 1e6

    # N > 1 end to end: the same K steps, every step's gathered frame copied to pinned host memory on rank 0 inside the
    # timed region (all ranks take part: the gather is a collective); device events, max over ranks
    e2e_multi_ms = None
    frames_extra = None
    other_extra = None
    e2e_zero_copy = False
    if tiles_mode:
        # the ranks' store kernels write straight into one pinned host frame shared by the ranks (each GPU over its own
        # PCIe link); falls back to "assemble on rank 0's GPU, then one device-to-host copy" when the registration fails
        shared = mg.SharedHostFrame(torch, dist, W, H)  # (collective; never raises on one rank alone)
        if not shared.usable:
            log("[bench] shared pinned host frame unavailable; e2e copies the assembled frame from rank 0")
            shared.close()
            shared = None
        if shared is not None:
            peer.render_to_host(cam, shared, max_depth=depth, traversal=args.traversal)
            torch.cuda.synchronize()
            dist.barrier()
            if rank == 0:
                import numpy as np
                a32, b32 = shared.array.view(np.uint32), single_frame_host.view(np.uint32)
                diff = int(((a32 != b32) & ~(np.isnan(shared.array) & np.isnan(single_frame_host))).any(axis=2).sum())
                tiles_check["host_zero_copy"] = {"pixels_differing_from_single_gpu_frame": diff}
                if diff:
                    raise SystemExit(f"zero-copy host frame differs from the single-GPU frame on {diff} pixels")
            e2e_zero_copy = True

        def step_e2e():
            if shared is not None:
                peer.render_to_host(cam, shared, max_depth=depth, traversal=args.traversal)
            else:
                step_tiles(to_host=True)

        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for k in range(2):
            step_e2e()
        dist.barrier()
        torch.cuda.synchronize()
        a.record()
        for k in range(args.steps):
            step_e2e()
        b2.record()
        torch.cuda.synchronize()
        dist.barrier()
        e2e_multi_ms = max_over_ranks(a.elapsed_time(b2)) / args.steps
        if shared is not None:
            shared.close()
        # extra key: the other transport of the same tile split
        other = step_tiles_gather if use_peer else step_tiles_peer
        for k in range(3):
            other()
        o_ms, _ = timed_per_step(other, args.steps)
        o_per = max_over_ranks(sum(o_ms)) / args.steps
        other_extra = {"transport": "nccl_gather" if use_peer else "peer_store", "value": rays_step / (o_per * 1e-3) / 1e6, "unit": "Mrays/s",
                       "ms_per_step": o_per}
        # extra key: the frames partition on the same ranks (weak scaling: N frames per step)
        for k in range(3):
            step_frames(k)
        drain_frames()
        per = max_over_ranks(timed_frames(args.steps))
        rays_frames = sum_over_ranks(ctx.last_stats()["rays_total"])
        frames_extra = {"value": rays_frames / (per * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": per, "scaling": "weak",
                        "note": f"frames partition: {world} frames per step (one per GPU, static camera), PPMColor frames gathered to rank 0 over NCCL, overlapped"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- e2e through crtb200_render with pinned host buffers (N = 1) / host copy of the gathered frame (N > 1) ----
    if world == 1:
        host_rgb = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
        e2e_ms = []
        for k in range(args.steps + 1):
            flush.fill_(k & 0xFF)
            torch.cuda.synchronize()
            t = time.perf_counter()
            ctx.render(cam, opt, rgb_out=host_rgb.numpy())
            dt = (time.perf_counter() - t) * 1e3
            if k:
                e2e_ms.append(dt)
        e2e = {"value": rays_step / (statistics.mean(e2e_ms) * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": statistics.mean(e2e_ms),
               "h2d_bytes_per_step": 48 + 40 + 16 * n_rects, "d2h_bytes_per_step": W * H * 12,
               "api": "crtb200_render(host camera/options -> pinned host float RGB)"}
    elif tiles_mode:
        e2e = {"value": rays_step / (e2e_multi_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_multi_ms,
               "h2d_bytes_per_step": (48 + 40 + 16 * n_rects) * world, "d2h_bytes_per_step": W * H * 12,
               "api": ("crtb200_render_device shard per rank; every rank's store kernel writes its tiles straight into ONE pinned host float frame "
                       "shared by the ranks (POSIX shm + cudaHostRegister: zero-copy over each GPU's own PCIe link), NCCL all-reduce as barrier"
                       if e2e_zero_copy else
                       "crtb200_render_device shard per rank, stored straight into rank 0's frame (CUDA IPC over NVLink) -> pinned host float RGB on rank 0"
                       if use_peer else "crtb200_render_device shard per rank -> NCCL gather -> crtb200_assemble_shards -> pinned host float RGB on rank 0")}
    else:
        e2e = None

    roofline = roofline_record(args.workload, cst, cst2, seq, ms_per_step, W, H) if cst is not None else None

    literal = None
    if world == 1 and args.traversal == 0:
        # the reference's literal visit-all itinerary (traversal = 1) beside the default, and their pixel difference (0)
        try:
            ex = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
            li = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
            ctx.render_device(cam, opt, d_rgb=ex.data_ptr(), stream=stream)
            l_opt = crt.make_options(max_depth=depth, rects=rects, n_rects=n_rects, traversal=1)
            for _ in range(3):
                ctx.render_device(cam, l_opt, d_rgb=li.data_ptr(), stream=stream)
            lms = []
            for k in range(min(args.steps, 5)):
                flush.fill_(k & 0xFF)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ctx.render_device(cam, l_opt, d_rgb=li.data_ptr(), stream=stream)
                b.record()
                torch.cuda.synchronize()
                lms.append(a.elapsed_time(b))
            diff = ((ex.view(torch.int32) != li.view(torch.int32)) & ~(torch.isnan(ex) & torch.isnan(li))).any(dim=2)
            literal = {"value": rays_step / (statistics.mean(lms) * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": statistics.mean(lms),
                       "pixels_differing_from_default": int(diff.sum().item()), "pixels": W * H,
                       "note": "traversal=1: the reference's visit-all itinerary with no culling and no hand-off (round-1 default)"}
            del ex, li
        except Exception as e:
            literal = {"error": str(e)}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            with tempfile.TemporaryDirectory() as td:
                ref = run_reference_full(scene_file, folder, tex, depth, 1, out_prefix=os.path.join(td, "ref"))
                if ref is not None:
                    ours = host_rgb.numpy()
                    same = (ours.view(np.uint32) == ref["rgb"].view(np.uint32)) | (np.isnan(ours) & np.isnan(ref["rgb"]))
                    cpu_baseline = {"value": ref["rays_total"] / ref["render_s"] / 1e6, "unit": "Mrays/s", "cores": ref["threads"],
                                    "kind": "reference", "sample": ref["sample"], "seconds": ref["render_s"],
                                    "rays": ref["rays_total"], "rays_equal_ours": bool(ref["rays_total"] == rays_step),
                                    "pixels_differing_from_ours": int((~same).any(axis=2).sum()), "pixels": W * H}
            if cpu_baseline is None:
                sys.path.insert(0, os.path.join(ROOT, "oracle"))
                import binding as ob
                nb = max(1, n_rects // 16)  # bounded sample for the port: 1/16 of the frame's bucket grid
                t = time.perf_counter()
                _, _, ost = ob.render(flat, cam, crt.make_options(max_depth=depth, rects=rects, n_rects=nb), want_hits=False)
                dt = time.perf_counter() - t
                cpu_baseline = {"value": ost["rays_total"] / dt / 1e6, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{nb} of {n_rects} buckets of the frame, OpenMP over rows", "seconds": dt}
        except Exception as e:  # the baseline is a report, never a reason to lose the bench line
            cpu_baseline = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}

    config5 = None
    if world == 1 and not args.no_config5:
        try:
            ctx.close()
            del frame, frame8
            config5 = config5_record(crt, torch, args, dev)
        except Exception as e:
            config5 = {"error": str(e)}

    per_frame_launches = kstats[-1]["kernel_launches"]
    launches = (per_frame_launches + (1 if (tiles_mode and not use_peer) else 0)) * (world if frames_mode else 1)
    line = {
        "metric": "Mrays/s (primary+secondary)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak" if (frames_mode or world == 1) else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "width": W, "height": H, "triangles": int(sf.info.n_triangles), "max_depth": depth,
                   "traversal": ("default: conservative culling + tail hand-off (results identical to the reference)" if args.traversal == 0
                                 else "literal visit-all itinerary"),
                   "parallelism": (f"frames{world}: one frame per GPU per step, rgb8 gathered to rank 0 (NCCL, overlapped)" if frames_mode
                                   else (f"tiles{world}: one frame, 8x4 tiles round-robin over {world} GPUs, every GPU's store kernel writes into rank 0's frame "
                                         f"(peer memory over NVLink, CUDA IPC; one NCCL all-reduce as completion barrier)" if use_peer else
                                         f"tiles{world}: one frame, 8x4 tiles round-robin over {world} GPUs, float slabs gathered to rank 0 (NCCL) and assembled")
                                   if world > 1 else "single"),
                   "l2": ("each frame streams ~1 GB of queues through the 126 MB L2 (inputs larger than L2)" if frames_mode
                          else "flushed between steps (256 MiB fill)"),
                   "rays_per_step": rays_step, "chunks_in_flight": args.concurrency,
                   "shadow_rays_zero_term_per_step_rank0": kstats[-1].get("shadow_rays_zero_term", 0),
                   "walks_handed_to_k_coop_per_step_rank0": kstats[-1].get("handoff_closest", 0) + kstats[-1].get("handoff_shadow", 0)},
        "clocks": clocks, "gpu_launches": launches * args.steps, "step_ms": step_ms,
    }
    if e2e:
        line["e2e"] = e2e
    if roofline:
        line["roofline"] = roofline
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    if literal:
        line["literal_walk"] = literal
    if tiles_check:
        line["tiles_check"] = tiles_check
    if other_extra:
        line["tiles_other_transport"] = other_extra
    if frames_extra:
        line["frames_mode"] = frames_extra
    if config5:
        line["config5"] = config5
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# config 5: camera-animation sequence (app/animation.cpp:24-38), frames round-robin over the GPUs
# ---------------------------------------------------------------------------------------------------------------------
def animation_arm(args):
    """A step = the whole F-frame orbit sequence.  Frame f is rendered by rank f % N (scene replicated); after every
    round of N frames the PPMColor frames are gathered to rank 0 over NCCL, double-buffered so the gather of round r
    overlaps the render of round r + 1.  value = rays of all F frames / sequence time; e2e adds the copy of every
    gathered frame to pinned host memory (what the reference's exportPPM consumes)."""
    import torch
    import torch.distributed as dist

    crt = importlib.import_module(PKG)
    scenes = importlib.import_module(PKG + ".scenes")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if world > 1 else 0)
    torch.cuda.set_device(dev)
    if rank == 0:
        scene_file, folder, kw, tex, depth = ensure_scene(args.workload, dict(width=args.width, height=args.height))
    if world > 1:
        dist.barrier()
    if rank != 0:
        scene_file, folder, kw, tex, depth = ensure_scene(args.workload, dict(width=args.width, height=args.height))
    sf = crt.SceneFile(scene_file, folder)
    flat = sf.flatten()
    ctx = crt.Context(dev.index)
    ctx.upload(flat, keepalive=sf)
    ctx.set_concurrency(args.concurrency)
    W, H = sf.info.width, sf.info.height
    F = args.animation
    center_z = -3.0 if args.workload == "synthetic_10M" else -4.0
    cams = [crt.Camera.make(p, r) for p, r in scenes.orbit_cameras(F, radius=5.12, center_z=center_z)]
    rects, n_rects = sf.rects(mode=crt.MODE_B200_WAVEFRONT)
    opt = crt.make_options(max_depth=depth, rects=rects, n_rects=n_rects, traversal=args.traversal)
    stream = torch.cuda.current_stream().cuda_stream
    rounds = (F + world - 1) // world
    frame = torch.zeros((H, W, 3), dtype=torch.float32, device=dev)
    bufs8 = [torch.zeros((H, W, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    gflat = [torch.zeros((world, H, W, 3), dtype=torch.uint8, device=dev) if rank == 0 else None for _ in range(2)]
    host8 = [torch.empty((world, H, W, 3), dtype=torch.uint8).pin_memory() if rank == 0 else None for _ in range(2)]
    works = [None, None]

    # untimed pass: ray count of every frame this rank owns (also the warm-up)
    rays_mine = 0
    for f in range(rank, F, world):
        ctx.render_device(cams[f], opt, d_rgb=frame.data_ptr(), d_rgb8=bufs8[0].data_ptr(), stream=stream)
        torch.cuda.synchronize()
        rays_mine += ctx.last_stats()["rays_total"]
    rays_t = torch.tensor([rays_mine], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(rays_t, op=dist.ReduceOp.SUM)
    rays_seq = int(rays_t.item())

    copy_stream = torch.cuda.Stream(device=dev)

    def finish(b, to_host):
        if works[b] is not None:
            if works[b] is not True:
                works[b].wait()
            if to_host and rank == 0:  # side stream: the copy engine works while the SMs render the next round
                copy_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(copy_stream):
                    host8[b].copy_(gflat[b], non_blocking=True)
            works[b] = None

    def sequence(to_host):
        for r in range(rounds):
            b = r & 1
            finish(b, to_host)
            f = r * world + rank
            if f < F:
                ctx.render_device(cams[f], opt, d_rgb=frame.data_ptr(), d_rgb8=bufs8[b].data_ptr(), stream=stream)
            if to_host and rank == 0:
                torch.cuda.current_stream().wait_stream(copy_stream)  # gflat[b] is about to be overwritten
            if world > 1:
                works[b] = dist.gather(bufs8[b], [gflat[b][i] for i in range(world)] if rank == 0 else None, dst=0, async_op=True)
            else:
                gflat[b][0].copy_(bufs8[b])
                works[b] = True
        finish(0, to_host)
        finish(1, to_host)
        if to_host and rank == 0:
            torch.cuda.current_stream().wait_stream(copy_stream)

    def timed(to_host):
        out = []
        for _ in range(args.steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a.record()
            sequence(to_host)
            b.record()
            torch.cuda.synchronize()
            out.append(a.elapsed_time(b))
        t = torch.tensor([sum(out)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / args.steps, out

    for _ in range(max(1, min(args.warmup, 2))):
        sequence(False)
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    ms, step_ms = timed(False)
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms, _ = timed(True)
    if rank == 0:
        launches = 0
        st = ctx.last_stats()
        launches = st["kernel_launches"] * F + (rounds if world > 1 else 0)
        line = {
            "metric": "Mrays/s (primary+secondary)", "value": rays_seq / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(1, min(args.warmup, 2)) + 1, "ms_per_step": ms, "ms_per_frame": ms / F,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "width": W, "height": H, "triangles": int(sf.info.n_triangles), "max_depth": depth,
                       "animation_frames": F, "camera_path": "app/animation.cpp orbit, radius 5.12, 360/F degrees per frame",
                       "traversal": "exact (reference visit-all order)" if args.traversal == 0 else "fast (ordered+culled)",
                       "parallelism": f"frames round-robin over {world} GPU(s), rgb8 frames gathered to rank 0 (NCCL, overlapped)",
                       "l2": "every frame has a new camera and streams its queues through L2; scene working set >= L2 for synthetic_10M",
                       "rays_per_step": rays_seq},
            "clocks": clocks, "gpu_launches": launches * args.steps, "step_ms": step_ms,
            "e2e": {"value": rays_seq / (e2e_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_ms, "ms_per_frame": e2e_ms / F,
                    "h2d_bytes_per_step": (48 + 40 + 16 * n_rects) * F, "d2h_bytes_per_step": rounds * world * W * H * 3,
                    "api": "crtb200_render_device per frame -> NCCL gather -> pinned host PPMColor frames on rank 0"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hw14_dragon_class", choices=list(WORKLOADS))
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--traversal", type=int, default=0)
    ap.add_argument("--parallelism", default="auto", choices=["auto", "frames", "tiles"],
                    help="N > 1: tiles (default) = one frame split by 8x4 tiles (strong scaling, config 4); frames = one frame per GPU per step (weak)")
    ap.add_argument("--no-config5", action="store_true", help="skip the synthetic_10M sub-record (N = 1)")
    ap.add_argument("--tile-transport", default="auto", choices=["auto", "gather", "peer"],
                    help="tiles, N > 1: peer = every rank's store kernel writes its tiles into rank 0's frame with 16-byte stores "
                         "(CUDA IPC over NVLink; one NCCL all-reduce as completion barrier); gather = compact slabs, one NCCL gather, "
                         "crtb200_assemble_shards; auto (default) = peer up to 4 GPUs, gather beyond.  The other one is timed as an extra key")
    ap.add_argument("--concurrency", type=int, default=2, help="chunk streams of a host-bound frame (e2e); the library default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--animation", type=int, default=0, help="F > 0: a step is the F-frame orbit animation (config 5), frames round-robin over GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("MASTER_PORT", "29511"), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.animation > 0:
        return animation_arm(args)
    return ours_arm(args)


if __name__ == "__main__":
    sys.exit(main())
