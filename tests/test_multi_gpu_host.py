"""CPU (gloo, world_size 2): the host side of the tile-sharded multi-GPU path -- shard maps, equal-size slabs, gather to
rank 0 and assembly -- with the CPU oracle supplying the pixel values (no GPU needed)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, scene_dir, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    crt = importlib.import_module(PKG)
    mg = importlib.import_module(PKG + ".multigpu")
    import binding as ob
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sf = crt.SceneFile("hw11_room.crtscene", scene_dir)
    flat = sf.flatten()
    W, H = sf.info.width, sf.info.height
    full, _, _ = ob.render(flat, sf.camera(), crt.make_options(), want_hits=False, threads=2)
    row, col, valid = mg.shard_pixel_map(W, H, rank, world)
    slab = np.zeros((mg.shard_items(W, H, world), 3), np.float32)
    slab[valid] = full[row[valid], col[valid]]
    t = torch.from_numpy(slab)
    gl = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gl, dst=0)
    if rank == 0:
        frame = mg.assemble_host(np.stack([g.numpy() for g in gl]), W, H)
        q.put(bool(np.array_equal(frame.view(np.uint32), full.view(np.uint32))))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_maps_partition_the_frame(crt):
    mg = importlib.import_module(PKG + ".multigpu")
    for (w, h, world) in [(192, 108, 2), (100, 70, 3), (1920, 1080, 8), (33, 5, 4)]:
        seen = np.zeros((h, w), np.int32)
        for r in range(world):
            row, col, valid = mg.shard_pixel_map(w, h, r, world)
            assert row.shape[0] == mg.shard_items(w, h, world)
            np.add.at(seen, (row[valid], col[valid]), 1)
        assert (seen == 1).all()


def test_gloo_gather_assembles_frame(built, scene_dir):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, scene_dir, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
