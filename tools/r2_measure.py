#!/usr/bin/env python3
"""Round-2 A/B matrix in one process (tools only): device ms per frame (median of N, concurrency 1, L2 flushed) for
  traversal   0 = default (conservative culling + tail hand-off)   1 = literal visit-all walk
  hand-off    CRT_TAIL_ITERS settings (read by crtb200_create, so one context per setting); -1 = off
  shards      whole frame and shard 0 of 2 / 4 / 8 (the per-rank work of a tile-sharded N-GPU frame)

  python tools/r2_measure.py --workloads hw14_dragon_class,synthetic_10M,hw11_room --tails 32,64,-1
"""
import argparse
import importlib
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def ctx_with_env(crt, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return crt.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="hw14_dragon_class")
    ap.add_argument("--tails", default="32,-1", help="CRT_TAIL_ITERS values; -1 = hand-off off")
    ap.add_argument("--shards", default="1,8")
    ap.add_argument("--frames", type=int, default=7)
    ap.add_argument("--json", default="")
    ap.add_argument("--chunks", default="", help="CRT_DEVICE_CHUNKS values to sweep (device-only frames split over k streams)")
    args = ap.parse_args()
    import torch
    crt = importlib.import_module(bench.PKG)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    rows = []
    for w in args.workloads.split(","):
        f, folder, kw, tex, depth = bench.ensure_scene(w, {})
        sf = crt.SceneFile(f, folder)
        flat = sf.flatten()
        configs = [("literal", 1, None)] + [(f"default tail {t}", 0, t) for t in args.tails.split(",")]
        if args.chunks:
            configs = [(f"{lab} chunks {ch}", trav, tail, ch) for (lab, trav, tail) in configs[1:] for ch in args.chunks.split(",")]
        else:
            configs = [(lab, trav, tail, "") for (lab, trav, tail) in configs]
        for label, trav, tail, chunks in configs:
            env = {}
            if chunks:
                os.environ["CRT_DEVICE_CHUNKS"] = chunks
            if tail is not None:  # "floor" or "floor:cap" or "floor:cap:start"
                parts = tail.split(":")
                env = {"CRT_TAIL_ITERS": parts[0]}
                if len(parts) > 1:
                    env["CRT_TAIL_CAP"] = parts[1]
                if len(parts) > 2:
                    env["CRT_TAIL_START"] = parts[2]
                if len(parts) > 3 and parts[3]:
                    env["CRT_TAIL_SMALL"] = parts[3]
                if len(parts) > 4:
                    env["CRT_TAIL_START_CLOSEST"] = parts[4]
            ctx = ctx_with_env(crt, env)
            ctx.upload(flat, keepalive=sf)
            ctx.set_concurrency(1)
            for shards in [int(x) for x in args.shards.split(",")]:
                opt = crt.make_options(max_depth=depth, traversal=trav, shard_index=0, shard_count=shards)
                if shards > 1:
                    out = torch.zeros((ctx.shard_items(shards), 3), dtype=torch.float32, device="cuda")
                else:
                    out = torch.zeros((sf.info.height, sf.info.width, 3), dtype=torch.float32, device="cuda")
                ms, cms, sms, ccm, csm = [], [], [], [], []
                for k in range(args.frames + 2):
                    flush.fill_(k & 0xFF)
                    torch.cuda.synchronize()
                    ctx.render_device(sf.camera(), opt, d_rgb=out.data_ptr(), stream=stream)
                    torch.cuda.synchronize()
                    st = ctx.last_stats()
                    if k >= 2:
                        ms.append(st["device_ms"])
                        cms.append(st["closest_ms"])
                        sms.append(st["shadow_ms"])
                        ccm.append(st["coop_closest_ms"])
                        csm.append(st["coop_shadow_ms"])
                row = {"workload": w, "config": label, "shards": shards, "device_ms": statistics.median(ms), "min_ms": min(ms),
                       "closest_ms": statistics.median(cms), "shadow_ms": statistics.median(sms), "rays": st["rays_total"],
                       "coop_closest_ms": statistics.median(ccm), "coop_shadow_ms": statistics.median(csm),
                       "handoff_closest": st["handoff_closest"], "handoff_shadow": st["handoff_shadow"],
                       "mrays_s": st["rays_total"] / statistics.median(ms) / 1e3}
                rows.append(row)
                print("%-18s %-16s shards %d  frame %7.3f ms (min %7.3f)  closest %6.3f (coop %6.3f, %7d walks)  shadow %6.3f (coop %6.3f, %7d walks)  %8.1f Mrays/s" %
                      (w, label, shards, row["device_ms"], row["min_ms"], row["closest_ms"], row["coop_closest_ms"], row["handoff_closest"],
                       row["shadow_ms"], row["coop_shadow_ms"], row["handoff_shadow"], row["mrays_s"]), flush=True)
                del out
            ctx.close()
    if args.json:
        with open(args.json, "w") as fh:
            json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()
