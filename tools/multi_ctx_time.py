#!/usr/bin/env python3
"""crtb200_create_multi on 1, 2, 4, 8 visible GPUs: wall-clock ms per crtb200_render_device frame of hw14 4K (one host
thread, frame assembled on device 0 by peer stores) and per crtb200_render frame into pinned host memory.  tools only."""
import ctypes as C
import importlib
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import numpy as np
    import torch
    crt = importlib.import_module(bench.PKG)
    n = C.c_int(0)
    crt.core().crtb200_device_count(C.byref(n))
    f, folder, kw, tex, depth = bench.ensure_scene("hw14_dragon_class", {})
    sf = crt.SceneFile(f, folder)
    flat = sf.flatten()
    H, W = sf.info.height, sf.info.width
    torch.cuda.set_device(0)
    frame = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda:0")
    host = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream().cuda_stream
    ref = None
    for k in [g for g in (1, 2, 4, 8) if g <= n.value]:
        ctx = crt.Context(list(range(k)))
        t = time.time()
        ctx.upload(flat, keepalive=sf)
        up = time.time() - t
        opt = crt.make_options(max_depth=depth)
        dev, e2e = [], []
        for i in range(13):
            torch.cuda.synchronize()
            t = time.perf_counter()
            ctx.render_device(sf.camera(), opt, d_rgb=frame.data_ptr(), stream=stream)
            torch.cuda.synchronize()
            if i >= 3:
                dev.append((time.perf_counter() - t) * 1e3)
        for i in range(8):
            t = time.perf_counter()
            _, _, _, st = ctx.render(sf.camera(), opt, rgb_out=host.numpy())
            if i >= 2:
                e2e.append((time.perf_counter() - t) * 1e3)
        out = frame.cpu().numpy()
        if ref is None:
            ref = out.copy()
        same = bool(np.array_equal(out.view(np.uint32), ref.view(np.uint32)) and np.array_equal(host.numpy().view(np.uint32), ref.view(np.uint32)))
        print(f"{k} GPU(s): upload {up:.2f}s  render_device {statistics.median(dev):.3f} ms (min {min(dev):.3f})  "
              f"render->host {statistics.median(e2e):.3f} ms  {st['rays_total'] / statistics.median(dev) / 1e3:.0f} Mrays/s  bit-equal to 1 GPU: {same}", flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
