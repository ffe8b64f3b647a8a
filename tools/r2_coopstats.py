#!/usr/bin/env python3
"""k_coop debug counters (needs a -DCRT_COOP_STATS=1 build via CRT_CORE_LIB): walks / iterations / box tests / triangle
tests of the hand-off pass beside the executed totals of the same (shard of a) frame.  tools only."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from r2_measure import ctx_with_env  # noqa: E402


def main():
    import torch
    crt = importlib.import_module(bench.PKG)
    stream = torch.cuda.current_stream().cuda_stream
    for w in sys.argv[1].split(","):
        f, folder, kw, tex, depth = bench.ensure_scene(w, {})
        sf = crt.SceneFile(f, folder)
        flat = sf.flatten()
        for tail in sys.argv[2].split(","):
            ctx = ctx_with_env(crt, {"CRT_TAIL_ITERS": tail})
            ctx.upload(flat, keepalive=sf)
            ctx.set_concurrency(1)
            for shards in (1, 8):
                out = torch.zeros((max(ctx.shard_items(shards), sf.info.height * sf.info.width), 3), dtype=torch.float32, device="cuda")
                cnt = crt.make_options(max_depth=depth, count_work=2, shard_index=0, shard_count=shards)
                ctx.render_device(sf.camera(), cnt, d_rgb=out.data_ptr(), stream=stream)
                torch.cuda.synchronize()
                c = ctx.last_stats()
                print(f"{w} tail {tail} shards {shards}: executed totals closest {c['node_tests_closest']} box / {c['triangle_tests_closest']} tri, "
                      f"shadow {c['node_tests_shadow']} box / {c['triangle_tests_shadow']} tri, rays {c['rays_total']}", flush=True)
                opt = crt.make_options(max_depth=depth, shard_index=0, shard_count=shards)
                for k in range(3):
                    sys.stderr.flush()
                    ctx.render_device(sf.camera(), opt, d_rgb=out.data_ptr(), stream=stream)
                    torch.cuda.synchronize()
                    s = ctx.last_stats()
                print(f"   frame {s['device_ms']:.3f} ms closest {s['closest_ms']:.3f} (coop {s['coop_closest_ms']:.3f}) shadow {s['shadow_ms']:.3f} (coop {s['coop_shadow_ms']:.3f})", flush=True)
                del out
            ctx.close()


if __name__ == "__main__":
    main()
