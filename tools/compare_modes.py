#!/usr/bin/env python3
"""Exact (traversal 0) vs culled (traversal 1) on the bench workloads: pixel / hit-id differences and frame time."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["hw07_scene0b", "hw11_room", "hw12_textures", "hw14_dragon_class"])
    ap.add_argument("--frames", type=int, default=3)
    args = ap.parse_args()
    crt = importlib.import_module(bench.PKG)
    for w in args.workloads:
        f, folder, kw, tex, depth = bench.ensure_scene(w, {})
        sf = crt.SceneFile(f, folder)
        ctx = crt.Context(0)
        ctx.upload(sf.flatten(), keepalive=sf)
        rects, n = sf.rects()
        out = {}
        for mode in (0, 1):
            opt = crt.make_options(max_depth=depth, rects=rects, n_rects=n, traversal=mode)
            best = 1e9
            for i in range(args.frames):
                rgb, rgb8, hits, st = ctx.render(sf.camera(), opt, want_rgb8=True, want_hits=True)
                best = min(best, st["device_ms"])
            opt2 = crt.make_options(max_depth=depth, rects=rects, n_rects=n, traversal=mode, count_work=2)
            _, _, _, st2 = ctx.render(sf.camera(), opt2, want_rgb=False)
            out[mode] = (rgb, rgb8, hits, st, best, st2)
        a, b = out[0], out[1]
        same = (a[0].view(np.uint32) == b[0].view(np.uint32)) | (np.isnan(a[0]) & np.isnan(b[0]))
        px = (~same).any(axis=2)
        d8 = np.abs(a[1].astype(int) - b[1].astype(int)).max(axis=2)
        hid = (a[2]["mesh"] != b[2]["mesh"]) | (a[2]["triangle"] != b[2]["triangle"])
        print(json.dumps({
            "workload": w, "pixels": int(px.size), "float_px_diff": int(px.sum()), "u8_px_diff": int((d8 > 0).sum()), "u8_max": int(d8.max()),
            "u8_gt1": int((d8 > 1).sum()), "u8_gt4": int((d8 > 4).sum()), "hit_id_diff": int(hid.sum()),
            "rays_exact": a[3]["rays_total"], "rays_culled": b[3]["rays_total"],
            "ms_exact": round(a[4], 3), "ms_culled": round(b[4], 3), "speedup": round(a[4] / b[4], 2),
            "tests_exact": [a[5]["node_tests"], a[5]["triangle_tests"]], "tests_culled": [b[5]["node_tests"], b[5]["triangle_tests"]]}))
        ctx.close()


if __name__ == "__main__":
    main()
