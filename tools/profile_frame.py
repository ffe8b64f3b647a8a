#!/usr/bin/env python3
"""Renders a few frames of one bench workload and prints per-kernel timings: the short command ncu wraps.

  python tools/profile_frame.py [--workload hw14_dragon_class] [--frames 3] [--width W --height H] [--traversal 0|1]
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="hw14_dragon_class")
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--traversal", type=int, default=0)
    ap.add_argument("--count", type=int, default=0)
    ap.add_argument("--concurrency", type=int, default=4)
    ap.add_argument("--shards", type=int, default=1, help="render shard 0 of N (per-rank work of an N-GPU run)")
    args = ap.parse_args()
    crt = importlib.import_module(bench.PKG)
    f, folder, kw, tex, depth = bench.ensure_scene(args.workload, dict(width=args.width, height=args.height))
    sf = crt.SceneFile(f, folder)
    ctx = crt.Context(0)
    ctx.upload(sf.flatten(), keepalive=sf)
    ctx.set_concurrency(args.concurrency)
    rects, n = sf.rects()
    if args.shards > 1:
        import torch
        opt = crt.make_options(max_depth=depth, traversal=args.traversal, count_work=args.count, shard_index=0, shard_count=args.shards)
        slab = torch.zeros((ctx.shard_items(args.shards), 3), dtype=torch.float32, device="cuda")
    else:
        opt = crt.make_options(max_depth=depth, rects=rects, n_rects=n, traversal=args.traversal, count_work=args.count)
    for i in range(args.frames):
        if args.shards > 1:
            ctx.render_device(sf.camera(), opt, d_rgb=slab.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            st = ctx.last_stats()
        else:
            _, _, _, st = ctx.render(sf.camera(), opt, want_rgb=False)
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()}))
    ctx.close()


if __name__ == "__main__":
    main()
