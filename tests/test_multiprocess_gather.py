"""N>1 host logic on CPU: two gloo processes partition a frame by interleaved tiles, fill their owned pixels and gather.
The gathered frame must equal the single-process frame bit for bit (ownership is disjoint, nothing is summed)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _expected(xres, yres):
    p = np.arange(xres * yres, dtype=np.float32)
    return np.stack([p, p * 0.5 + 1.0, -p], -1).reshape(yres, xres, 3)


def _worker(rank, world, port, xres, yres, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vermeer_b200.multigpu import FrameGather
    from vermeer_b200.partition import owned_pixels
    g = FrameGather(xres, yres, rank, world, "cpu")
    fb = torch.full((yres * xres, 3), float("nan"))           # non-owned pixels must never leak into the result
    own = owned_pixels(xres, yres, rank, world)
    fb[torch.as_tensor(own)] = torch.from_numpy(_expected(xres, yres).reshape(-1, 3)[own])
    full = g.gather(fb)
    if rank == 0:
        np.save(out, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process(tmp_path):
    xres, yres, world = 200, 140, 2
    out = str(tmp_path / "frame.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, xres, yres, out), nprocs=world, join=True)
    got = np.load(out)
    assert np.array_equal(got.view(np.uint32), _expected(xres, yres).view(np.uint32))


def test_single_rank_gather_is_identity():
    from vermeer_b200.multigpu import FrameGather
    xres, yres = 70, 50
    g = FrameGather(xres, yres, 0, 1, "cpu")
    fb = torch.from_numpy(_expected(xres, yres).reshape(-1, 3).copy())
    assert torch.equal(g.gather(fb).view(-1, 3), fb)
