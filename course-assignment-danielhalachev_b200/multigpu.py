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


class PeerStoreRenderer:
    """Per-rank helper for the tile split WITHOUT a gather: rank 0 owns the frame (float RGB + PPMColor bytes) and exports
    it with CUDA IPC; every rank renders its shard with `shard_full_frame`, so its store kernel writes the pixels straight
    into rank 0's frame over NVLink / NVSwitch.  One tiny NCCL all-reduce per frame is the completion barrier between
    the ranks' streams.  The frame tensors must be whole cudaMalloc allocations (allocated here, before anything else
    shares their block)."""

    def __init__(self, crt, ctx, torch, dist, device, width: int, height: int):
        self.crt, self.ctx, self.torch, self.dist, self.device = crt, ctx, torch, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.frame = self.frame8 = None
        self._mapped = []
        handles = torch.zeros((2, 64), dtype=torch.uint8, device=device)
        if self.rank == 0:
            # cudaMalloc directly (not torch's caching allocator): an IPC handle names a whole allocation
            n = width * height * 3
            self._own = [crt.ipc_alloc(device.index, n * 4), crt.ipc_alloc(device.index, n)]
            self.d_rgb, self.d_rgb8 = self._own
            import numpy as np
            h = np.stack([np.frombuffer(crt.ipc_export(p), dtype=np.uint8) for p in self._own])
            handles.copy_(torch.from_numpy(h.copy()))
        dist.broadcast(handles, src=0)
        if self.rank != 0:
            hb = handles.cpu().numpy()
            self.d_rgb = crt.ipc_open(device.index, bytes(hb[0]))
            self.d_rgb8 = crt.ipc_open(device.index, bytes(hb[1]))
            self._mapped = [self.d_rgb, self.d_rgb8]
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.width, self.height = width, height

    def frame_tensors(self):
        """rank 0: (float HxWx3, uint8 HxWx3) views of the frame the ranks store into"""
        assert self.rank == 0
        t = self.torch

        class _Arr:  # minimal __cuda_array_interface__ holder
            def __init__(s, ptr, shape, typestr):
                s.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}
        f = t.as_tensor(_Arr(self.d_rgb, (self.height, self.width, 3), "<f4"), device=self.device)
        b = t.as_tensor(_Arr(self.d_rgb8, (self.height, self.width, 3), "|u1"), device=self.device)
        return f, b

    def render(self, camera, max_depth: int = 5, traversal: int = 0) -> None:
        """Asynchronous on the current torch stream; after it (on rank 0's stream) the frame is complete."""
        stream = self.torch.cuda.current_stream().cuda_stream
        opt = self.crt.make_options(max_depth=max_depth, shard_index=self.rank, shard_count=self.world, traversal=traversal,
                                    shard_full_frame=True)
        self.ctx.render_device(camera, opt, d_rgb=self.d_rgb, d_rgb8=self.d_rgb8, stream=stream)
        self.dist.all_reduce(self._flag)  # completion barrier: rank 0's stream continues when every rank has stored

    def render_to_host(self, camera, host_frame: "SharedHostFrame", max_depth: int = 5, traversal: int = 0) -> None:
        """The same split, but every rank's store kernel writes its tiles into `host_frame` (pinned host memory shared by
        the ranks) over its own PCIe link.  After the call (on every rank's stream, behind the all-reduce) the host frame
        is complete."""
        stream = self.torch.cuda.current_stream().cuda_stream
        opt = self.crt.make_options(max_depth=max_depth, shard_index=self.rank, shard_count=self.world, traversal=traversal,
                                    shard_full_frame=True)
        self.ctx.render_device(camera, opt, d_rgb=host_frame.device_ptr, stream=stream)
        self.dist.all_reduce(self._flag)

    def close(self):
        for p in self._mapped:
            self.crt.ipc_close(self.device.index, p)
        self._mapped = []
        if self.rank == 0 and getattr(self, "_own", None):
            for p in self._own:
                self.crt.ipc_free(self.device.index, p)
            self._own = None


class SharedHostFrame:
    """A float RGB frame in pinned host memory that every rank of one node can write from its GPU: rank 0 creates a POSIX
    shared-memory segment, every rank maps it and registers it with CUDA (cudaHostRegister, portable + mapped).  With
    `PeerStoreRenderer.render_to_host` each rank's store kernel then writes its tiles straight into this frame over its
    own GPU's PCIe link (zero-copy), so an N-GPU frame reaches the host without a pass through rank 0's GPU and its
    single link.  `array` is the numpy view (H, W, 3); `device_ptr` the address the kernels use."""

    def __init__(self, torch, dist, width: int, height: int):
        from multiprocessing import shared_memory
        self.torch, self.dist = torch, dist
        self.rank = dist.get_rank()
        self.nbytes = width * height * 3 * 4
        self.shm = None
        self.array = None
        self.registered = False
        self.host_ptr = self.device_ptr = 0
        # no rank may leave this constructor early: every step below is followed by a collective
        name = [None]
        if self.rank == 0:
            try:
                self.shm = shared_memory.SharedMemory(create=True, size=self.nbytes)
                name[0] = self.shm.name
            except Exception:
                self.shm = None
        dist.broadcast_object_list(name, src=0)
        if name[0] is not None:
            try:
                if self.rank != 0:
                    self.shm = shared_memory.SharedMemory(name=name[0])
                    try:  # Python < 3.13 registers attached segments with the resource tracker as if this process owned
                        from multiprocessing import resource_tracker  # them and warns about a "leak" at exit
                        resource_tracker.unregister(self.shm._name, "shared_memory")
                    except Exception:
                        pass
                self.array = np.ndarray((height, width, 3), dtype=np.float32, buffer=self.shm.buf)
                self.host_ptr = self.array.ctypes.data
                rc = int(torch.cuda.cudart().cudaHostRegister(self.host_ptr, self.nbytes, 1 | 2))  # portable | mapped
                self.registered = rc == 0
                self.device_ptr = self.host_ptr  # unified addressing: registered host memory keeps its address on x86-64 ...
                if self.registered:
                    try:  # ... but ask when the runtime binding is there
                        from cuda.bindings import runtime as cudart
                        err, dptr = cudart.cudaHostGetDevicePointer(self.host_ptr, 0)
                        if int(err) == 0:
                            self.device_ptr = int(dptr)
                    except Exception:
                        pass
            except Exception:
                self.registered = False
        ok = torch.tensor([1 if self.registered else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        self.usable = bool(ok.item())

    def close(self):
        if getattr(self, "registered", False):
            self.torch.cuda.cudart().cudaHostUnregister(self.host_ptr)
            self.registered = False
        if getattr(self, "shm", None) is not None:
            self.array = None
            try:
                self.shm.close()
                if self.rank == 0:
                    self.shm.unlink()
            except Exception:
                pass
            self.shm = None


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
