#!/usr/bin/env python3
"""e2e frame time (crtb200_render into pinned host memory) for a few chunking settings.  usage: e2e_time.py [workload]"""
import importlib
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "hw14_dragon_class"
    crt = importlib.import_module(bench.PKG)
    f, folder, kw, tex, depth = bench.ensure_scene(wl, dict(width=0, height=0))
    sf = crt.SceneFile(f, folder)
    flat = sf.flatten()
    rects, n = sf.rects()
    host = torch.empty((sf.info.height, sf.info.width, 3), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    staggers = os.environ.get("E2E_STAGGERS", "0,0.1,0.15,0.2,0.3").split(",")
    per_sets = [int(x) for x in os.environ.get("E2E_PER_SET", "2,3").split(",")]
    concs = [int(x) for x in os.environ.get("E2E_SETS", "3,4,6").split(",")]
    for stagger in staggers:
        os.environ["CRT_CHUNK_STAGGER"] = stagger
        for per_set in per_sets:
            os.environ["CRT_HOST_CHUNKS_PER_SET"] = str(per_set)
            for conc in concs:
                ctx = crt.Context(0)
                ctx.upload(flat, keepalive=sf)
                ctx.set_concurrency(conc)
                opt = crt.make_options(max_depth=depth, rects=rects, n_rects=n)
                ts = []
                for k in range(9):
                    flush.fill_(k)
                    torch.cuda.synchronize()
                    t = time.perf_counter()
                    ctx.render(sf.camera(), opt, rgb_out=host.numpy())
                    ts.append((time.perf_counter() - t) * 1e3)
                print(f"{wl} stagger {stagger} chunks/set {per_set} sets {conc}: e2e {statistics.median(ts[2:]):.3f} ms (min {min(ts[2:]):.3f})", flush=True)
                ctx.close()


if __name__ == "__main__":
    main()
