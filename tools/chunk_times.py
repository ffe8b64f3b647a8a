#!/usr/bin/env python3
"""Timeline of a host-bound frame (tools only): when each chunk is stored and when its band has reached the host.
usage: chunk_times.py [workload]   (CRT_CHUNK_STAGGER / CRT_HOST_CHUNKS_PER_SET / E2E_SETS as in e2e_time.py)"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "hw14_dragon_class"
crt = importlib.import_module(bench.PKG)
f, folder, kw, tex, depth = bench.ensure_scene(wl, dict(width=0, height=0))
sf = crt.SceneFile(f, folder)
flat = sf.flatten()
rects, n = sf.rects()
host = torch.empty((sf.info.height, sf.info.width, 3), dtype=torch.float32).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ctx = crt.Context(0)
ctx.upload(flat, keepalive=sf)
ctx.set_concurrency(int(os.environ.get("E2E_SETS", "4")))
opt = crt.make_options(max_depth=depth, rects=rects, n_rects=n)
for k in range(5):
    flush.fill_(k)
    torch.cuda.synchronize()
    if k == 4:
        os.environ["CRT_CHUNK_TIMES"] = "1"
    t = time.perf_counter()
    ctx.render(sf.camera(), opt, rgb_out=host.numpy())
    print(f"frame {k}: e2e {(time.perf_counter() - t) * 1e3:.3f} ms", flush=True)
ctx.close()
