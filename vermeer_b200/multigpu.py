"""Multi-GPU plumbing: one process per GPU, image partitioned by interleaved 32x32 tiles, scene replicated, and ONE
exchange per frame. There is no data-path collective while rendering: every (pixel, iteration) sample is an
independent function of (x, y, iter, framescramble[pixel]) (core/render.go:89-125) and writes only its own pixel (:127-129).

The exchange itself lives in the library (csrc/comm.cu: vg_comm_init / vg_gather_frame — NCCL send/recv on the context's
stream between the library's own pack and scatter kernels). `init_library_comm` below is the small amount of host glue a
launcher needs: rank 0 creates the NCCL id, the host distributes it (here over torch.distributed, whatever its backend), every
rank calls vg_comm_init.

`FrameGather` is the same exchange written with torch collectives. It is NOT on the product path: the CPU test suite uses it
(gloo, world_size 2) as an executable statement of what the gathered frame must be, against the library's own tile lists.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .partition import owned_pixels


def init_library_comm(dev, rank: int, world: int):
    """Collective over torch.distributed's default group: vg_comm_unique_id on rank 0 -> broadcast -> vg_comm_init."""
    if world == 1:
        dev.comm_init(0, 1, None)
        return
    obj = [dev.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    dev.comm_init(rank, world, obj[0])


class FrameGather:
    """Torch twin of vg_gather_frame (tests only). Pre-computes the index lists once; `gather(fb)` is one all_gather plus one
    index_copy."""

    def __init__(self, xres: int, yres: int, rank: int, world: int, device):
        self.xres, self.yres, self.rank, self.world = xres, yres, rank, world
        self.device = torch.device(device)
        lists = [owned_pixels(xres, yres, r, world) for r in range(world)]
        self.counts = [len(l) for l in lists]
        self.pad = max(self.counts) if self.counts else 0
        self.own = torch.as_tensor(lists[rank], dtype=torch.long, device=self.device)
        # scatter index for the gathered (world, pad) layout; padding slots point at pixel 0 and are masked out
        idx = np.zeros((world, self.pad), np.int64)
        mask = np.zeros((world, self.pad), bool)
        for r, l in enumerate(lists):
            idx[r, :len(l)] = l
            mask[r, :len(l)] = True
        self.scatter_idx = torch.as_tensor(idx.reshape(-1)[mask.reshape(-1)], dtype=torch.long, device=self.device)
        self.valid = torch.as_tensor(np.nonzero(mask.reshape(-1))[0], dtype=torch.long, device=self.device)
        self.send = torch.zeros((self.pad, 3), dtype=torch.float32, device=self.device)
        self.recv = torch.zeros((world, self.pad, 3), dtype=torch.float32, device=self.device)

    def gather(self, fb: torch.Tensor) -> torch.Tensor:
        """fb: this rank's full-frame (yres*xres, 3) buffer (non-owned pixels are ignored). Returns the complete frame on
        every rank. Owned pixels are copied, never summed, so the result is bit-identical to a single-GPU render."""
        fb = fb.view(-1, 3)
        self.send[: self.counts[self.rank]] = fb.index_select(0, self.own)
        if self.world > 1:
            dist.all_gather_into_tensor(self.recv.view(-1), self.send.view(-1))
        else:
            self.recv[0] = self.send
        out = torch.empty((self.yres * self.xres, 3), dtype=torch.float32, device=self.device)
        out.index_copy_(0, self.scatter_idx, self.recv.view(-1, 3).index_select(0, self.valid))
        return out.view(self.yres, self.xres, 3)

    @property
    def bytes_per_rank(self) -> int:
        return self.pad * 12
