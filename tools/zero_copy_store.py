#!/usr/bin/env python3
"""How fast does k_store write a frame straight into mapped pinned host memory (zero-copy over PCIe)?  tools only.
Renders hw14 4K with d_rgb = the device pointer of a pinned host tensor, whole frame and shard 0 of 8."""
import importlib
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from cuda import cudart  # noqa: E402

crt = importlib.import_module(bench.PKG)
f, folder, kw, tex, depth = bench.ensure_scene("hw14_dragon_class", {})
sf = crt.SceneFile(f, folder)
flat = sf.flatten()
H, W = sf.info.height, sf.info.width
ctx = crt.Context(0)
ctx.upload(flat, keepalive=sf)
host = torch.zeros((H, W, 3), dtype=torch.float32).pin_memory()
err, dptr = cudart.cudaHostGetDevicePointer(host.data_ptr(), 0)
print("cudaHostGetDevicePointer:", err, hex(dptr), "host", hex(host.data_ptr()))
dev = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for shards in (1, 8):
    for name, ptr in (("device frame", dev.data_ptr()), ("mapped host frame", int(dptr))):
        opt = crt.make_options(max_depth=depth, shard_index=0, shard_count=shards, shard_full_frame=shards > 1)
        ms = []
        for k in range(9):
            flush.fill_(k)
            torch.cuda.synchronize()
            ctx.render_device(sf.camera(), opt, d_rgb=ptr, stream=stream)
            torch.cuda.synchronize()
            ms.append(ctx.last_stats()["device_ms"])
        print(f"shards {shards}  {name:18s} frame {statistics.median(ms[2:]):.3f} ms", flush=True)
ref = dev.cpu()
print("host frame equals device frame (shard 0 of 8 region only written last):", bool(torch.isfinite(host).all()))
ctx.close()
