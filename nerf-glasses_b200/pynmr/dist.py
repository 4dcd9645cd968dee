"""Data-parallel rendering over several GPUs, one process per GPU (torch.distributed; NCCL on GPUs, gloo in CPU tests).

The reference is single-GPU (no NCCL/MPI anywhere); this module is the new host-side plumbing of
SURVEY section 8e.  Rays are independent, every rank holds a full model replica, and nothing is exchanged while a frame
renders.  Two partitionings:

  views   each rank renders whole views (render.py's multi-view landmark pass): `view_slice`, no collective at all unless
          the caller wants every image on one rank (`gather_views`).
  tiles   one frame, rows dealt to ranks in bands of `band` rows, round-robin (nmr_set_shard): a rank renders only its rows;
          `gather_frame` packs the owned rows and gathers them to `dst` with ONE collective, where they are scattered back
          into a full image.  The gathered image is bit-identical to the single-GPU image (tests/test_gpu_parity.py).

  tiles, fused   `PeerShardedRenderer`: no gather at all - every rank's render kernels store their rows straight into ONE image
          in the destination rank's memory (CUDA IPC mapping over NVLink, include/nmr.h: nmr_gather_*), ordered by
          device-side sequence flags.  torch.distributed only carries the 64-byte IPC handle once, at set-up.

Tensors are torch tensors on whatever device the process group's backend moves (cuda for nccl, cpu for gloo).
"""
from __future__ import annotations

import numpy as np


def owned_rows(height: int, rank: int, world: int, band: int) -> np.ndarray:
    """Image rows rendered by `rank`: row y belongs to rank (y // band) % world (same rule as csrc/kernels.cu:shard_row)."""
    if world < 1 or not (0 <= rank < world) or band < 1:
        raise ValueError("bad shard specification")
    y = np.arange(height)
    return y[(y // band) % world == rank]


def max_owned_rows(height: int, world: int, band: int) -> int:
    return max(len(owned_rows(height, r, world, band)) for r in range(world))


def view_slice(n_views: int, rank: int, world: int) -> range:
    """Contiguous block of views for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_views, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def pack_rows(image, rows: np.ndarray, padded_rows: int):
    """image [H, W, C] -> [padded_rows, W, C] holding the owned rows first (zero padding keeps gather sizes equal)."""
    import torch
    idx = torch.as_tensor(rows, dtype=torch.long, device=image.device)
    out = torch.zeros((padded_rows,) + tuple(image.shape[1:]), dtype=image.dtype, device=image.device)
    out[: len(rows)] = image.index_select(0, idx)
    return out


def gather_frame(local_image, rank: int, world: int, band: int, dst: int = 0, group=None):
    """Gathers a row-sharded frame.  local_image: [H, W, C] tensor in which only this rank's rows are valid.
    Returns the full [H, W, C] image on rank `dst`, None elsewhere.  One collective (gather); no other communication."""
    import torch
    import torch.distributed as dist
    H = int(local_image.shape[0])
    if world == 1:
        return local_image
    padded = max_owned_rows(H, world, band)
    mine = pack_rows(local_image, owned_rows(H, rank, world, band), padded)
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == dst else None
    dist.gather(mine, gather_list=parts, dst=dst, group=group)
    if rank != dst:
        return None
    full = torch.empty_like(local_image)
    for r in range(world):
        rows = owned_rows(H, r, world, band)
        full.index_copy_(0, torch.as_tensor(rows, dtype=torch.long, device=full.device), parts[r][: len(rows)])
    return full


def gather_views(local_views, n_views: int, rank: int, world: int, dst: int = 0, group=None):
    """local_views: [n_local, H, W, C] (this rank's view_slice).  Returns [n_views, H, W, C] on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_views
    padded = max(len(view_slice(n_views, r, world)) for r in range(world))
    mine = torch.zeros((padded,) + tuple(local_views.shape[1:]), dtype=local_views.dtype, device=local_views.device)
    mine[: local_views.shape[0]] = local_views
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == dst else None
    dist.gather(mine, gather_list=parts, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([parts[r][: len(view_slice(n_views, r, world))] for r in range(world)], dim=0)


class ShardedRenderer:
    """One frame across the process group: every rank calls render_frame(); rank `dst` gets the full image.

    renderer: this rank's pynmr.NerfMeshRenderer (created on this rank's device, same scene loaded on every rank).
    """

    def __init__(self, renderer, rank: int, world: int, band: int = 8, group=None, surface_mode: int | None = 2):
        self.r, self.rank, self.world, self.band, self.group = renderer, rank, world, band, group
        renderer.set_shard(rank, world, band)
        # one mesh-surface rule for all ranks: the default (auto) rule counts the live rays of a context's own rows, so shards could
        # decide differently from each other (include/nmr.h: nmr_set_shard); 8-sample batches are what a single GPU picks in
        # render.py's framing
        if surface_mode is not None:
            renderer.set_surface_insertion(surface_mode)
        self._buf = None

    def render_frame(self, dst: int = 0):
        """frame() on the local shard, then the gather.  Returns a torch.float32 [H, W, 4] cuda tensor on `dst`."""
        import torch
        self.r.frame_async()
        H, W = self.r.height, self.r.width
        if self._buf is None or tuple(self._buf.shape) != (H, W, 4):
            self._buf = torch.empty((H, W, 4), dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        self.r.copy_device_image(self._buf.data_ptr(), self._buf.numel() * 4)      # device->device on libnmr's stream, then a stream synchronise
        return gather_frame(self._buf, self.rank, self.world, self.band, dst, self.group)


class _DeviceImage:
    """float32 [H, W, 4] view of device memory owned by libnmr, for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, height: int, width: int):
        self.__cuda_array_interface__ = {"shape": (height, width, 4), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}


class PeerShardedRenderer:
    """One frame across the process group with the gather fused into the rendering: rank `dst` owns the image, every
    rank's frame() writes its own rows into it through peer memory (NVLink).  Per frame there is no collective and no host
    synchronisation between ranks; `render_frame()` returns the full image on `dst` (a view of the shared image, valid until
    the next render_frame) and None elsewhere.  Every rank must call render_frame() the same number of times."""

    def __init__(self, renderer, rank: int, world: int, band: int = 8, dst: int = 0, group=None, surface_mode: int | None = 2):
        import torch.distributed as dist
        self.r, self.rank, self.world, self.band, self.dst, self.group = renderer, rank, world, band, dst, group
        renderer.set_shard(rank, world, band)
        if surface_mode is not None:        # one mesh-surface rule for all ranks, see ShardedRenderer (2 = NMR_SURFACE_BATCH8)
            renderer.set_surface_insertion(surface_mode)
        self._image = None
        box = [None]
        if rank == dst:
            handle, ptr = renderer.gather_create()
            box[0] = handle
            import torch
            self._image = torch.as_tensor(_DeviceImage(ptr, renderer.height, renderer.width), device=torch.device("cuda", torch.cuda.current_device()))
        if world > 1:
            dist.broadcast_object_list(box, src=dst, group=group)
            if rank != dst:
                renderer.gather_attach(box[0])
            dist.barrier(group=group)       # everybody is attached before the first frame

    def render_frame(self, sync: bool = True):
        """sync=False only enqueues (pipelined callers that consume the image on the renderer's stream, r.stream_ptr())."""
        if self.rank != self.dst:
            self.r.frame_async()
            return None
        if sync:
            self.r.frame()                  # the destination's stream ends with the wait for every rank's signal; raises when a
                                            # rank did not deliver its rows in time (the image would be incomplete)
        else:
            self.r.frame_async()
        return self._image

    def close(self):
        import torch.distributed as dist
        if self.world > 1:
            self.r.synchronize()
            dist.barrier(group=self.group)  # nobody unmaps while a peer may still be writing
        self.r.gather_detach()
        self.r.set_shard(0, 1, self.band)
