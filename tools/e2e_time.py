#!/usr/bin/env python3
"""e2e frame time (crtb200_render into pinned host memory) for a few chunking settings (tools only).
usage: e2e_time.py [workload]     env: E2E_BAND_STREAM=on,off  E2E_PER_SET=0,1,2,3 (0 = library default)  E2E_SETS=1,2,3,4"""
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
    for band in os.environ.get("E2E_BAND_STREAM", "on,off").split(","):
        if band == "off":
            os.environ["CRT_NO_BAND_STREAM"] = "1"
        else:
            os.environ.pop("CRT_NO_BAND_STREAM", None)
        for per_set in [int(x) for x in os.environ.get("E2E_PER_SET", "0,1,2").split(",")]:
            if per_set:
                os.environ["CRT_HOST_CHUNKS_PER_SET"] = str(per_set)
            else:
                os.environ.pop("CRT_HOST_CHUNKS_PER_SET", None)
            for conc in [int(x) for x in os.environ.get("E2E_SETS", "2,3,4").split(",")]:
                ctx = crt.Context(0)
                ctx.upload(flat, keepalive=sf)
                ctx.set_concurrency(conc)
                opt = crt.make_options(max_depth=depth, rects=rects, n_rects=n)
                ts = []
                for k in range(11):
                    flush.fill_(k)
                    torch.cuda.synchronize()
                    t = time.perf_counter()
                    ctx.render(sf.camera(), opt, rgb_out=host.numpy())
                    ts.append((time.perf_counter() - t) * 1e3)
                print(f"{wl} band stream {band} chunks/set {per_set or 'default'} sets {conc}: e2e {statistics.median(ts[2:]):.3f} ms (min {min(ts[2:]):.3f})", flush=True)
                ctx.close()


if __name__ == "__main__":
    main()
