"""Tile-sharded multi-GPU rendering: one process per GPU, scene replicated, the frame's 8x4-pixel tiles
dealt round-robin to the ranks, per-rank compact slabs gathered to rank 0 over NCCL (torch.distributed), scattered
into the frame by crtb200_assemble_shards.  No exchange step exists inside a frame, so the only collective is the gather.

The reference has no multi-device path (single process, std::thread over buckets, RayTracer.cpp:114-202); the unit
sharded here is the same one it parallelises over -- independent pixels of one frame.
"""
from __future__ import annotations

import numpy as np

TILE_W, TILE_H = 8, 4  # must match item_pixel() in csrc/crt_kernels.cuh


def shard_items(width: int, height: int, shard_count: int) -> int:
    tiles = ((width + TILE_W - 1) // TILE_W) * ((height + TILE_H - 1) // TILE_H)
    return ((tiles + shard_count - 1) // shard_count) * 32


def shard_pixel_map(width: int, height: int, shard_index: int, shard_count: int):
    """Host mirror of item_pixel(): for every slab item of a shard its (row, col) and validity."""
    tiles_x = (width + TILE_W - 1) // TILE_W
    n_tiles = tiles_x * ((height + TILE_H - 1) // TILE_H)
    items = shard_items(width, height, shard_count)
    i = np.arange(items, dtype=np.int64)
    tile = (i >> 5) * shard_count + shard_index
    lane = i & 31
    col = (tile % tiles_x) * TILE_W + (lane & 7)
    row = (tile // tiles_x) * TILE_H + (lane >> 3)
    valid = (tile < n_tiles) & (col < width) & (row < height)
    return row, col, valid


def assemble_host(slabs: np.ndarray, width: int, height: int) -> np.ndarray:
    """numpy equivalent of crtb200_assemble_shards (used by the CPU tests of the gather logic)."""
    world = slabs.shape[0]
    frame = np.zeros((height, width, slabs.shape[2]), dtype=slabs.dtype)
    for r in range(world):
        row, col, valid = shard_pixel_map(width, height, r, world)
        frame[row[valid], col[valid]] = slabs[r][valid]
    return frame


class ShardedRenderer:
    """Per-rank helper: render this rank's shard, gather on rank 0, assemble.  `ctx` is a crt.Context with the scene
    uploaded on this rank's GPU; `dist` is an initialised torch.distributed NCCL group (world_size = shard count)."""

    def __init__(self, crt, ctx, torch, dist, device):
        self.crt, self.ctx, self.torch, self.dist = crt, ctx, torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.items = ctx.shard_items(self.world)
        self.slab = torch.zeros((self.items, 3), dtype=torch.float32, device=device)
        self.slabs = torch.zeros((self.world, self.items, 3), dtype=torch.float32, device=device) if self.rank == 0 else None
        self.gather_list = list(self.slabs.unbind(0)) if self.rank == 0 else None

    def render(self, camera, max_depth: int = 5, traversal: int = 0, frame=None, frame8=None) -> None:
        """Asynchronous on the current torch stream.  frame / frame8: full-frame device tensors on rank 0."""
        stream = self.torch.cuda.current_stream().cuda_stream
        opt = self.crt.make_options(max_depth=max_depth, shard_index=self.rank, shard_count=self.world, traversal=traversal)
        self.ctx.render_device(camera, opt, d_rgb=self.slab.data_ptr(), stream=stream)
        self.dist.gather(self.slab, self.gather_list, dst=0)
        if self.rank == 0:
            self.ctx.assemble_shards(self.slabs.data_ptr(), self.world, d_rgb=frame.data_ptr() if frame is not None else 0,
                                     d_rgb8=frame8.data_ptr() if frame8 is not None else 0, stream=stream)
