"""Image partition across GPUs: the reference's 32x32 tiles (core/render.go:196-199) dealt to ranks by
(tx + ty*k) % world with k odd, so neither rows nor columns of tiles alias onto one rank.
Python twin of the formula in csrc/render.cu (prepare()); used by the multi-GPU gather and by tests."""
from __future__ import annotations

import numpy as np


def tile_stride(tiles_x: int) -> int:
    return tiles_x + 1 if tiles_x % 2 == 0 else tiles_x


def owned_pixels(xres: int, yres: int, rank: int, world: int) -> np.ndarray:
    """Full-frame pixel indices (row-major) owned by `rank`, in the device's tile-major order."""
    tiles_x, tiles_y = (xres + 31) // 32, (yres + 31) // 32
    k = tile_stride(tiles_x)
    out = []
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            if (tx + ty * k) % world != rank:
                continue
            ys = np.arange(ty * 32, min(ty * 32 + 32, yres))
            xs = np.arange(tx * 32, min(tx * 32 + 32, xres))
            out.append((ys[:, None] * xres + xs[None, :]).reshape(-1))
    return np.concatenate(out) if out else np.zeros(0, np.int64)
