"""Image partition across GPUs: the reference's 32x32 tiles (core/render.go:196-199) dealt to ranks by
(tx + ty*k) % world with k >= tilesX the smallest stride coprime with `world`, so that neither rows nor columns of tiles
alias onto one rank for any world size.

Python twin of csrc/comm.cu (owned_pixels / partition_stride); tests check the two against each other through the library's
host-only entry point vg_owned_pixels."""
from __future__ import annotations

from math import gcd

import numpy as np


def tile_stride(tiles_x: int, world: int = 1) -> int:
    k = tiles_x
    while gcd(k, world) != 1:
        k += 1
    return k


def owned_pixels(xres: int, yres: int, rank: int, world: int, pixel_block: bool = True) -> np.ndarray:
    """Full-frame pixel indices (row-major) owned by `rank`, in the device's path order: tile by tile, 8x4 pixel blocks,
    Morton order inside a block (bits of l, low to high: x0 y0 x1 y1 x2)."""
    tiles_x, tiles_y = (xres + 31) // 32, (yres + 31) // 32
    k = tile_stride(tiles_x, world)
    if pixel_block:
        b, l = np.meshgrid(np.arange(32), np.arange(32), indexing="ij")
        lx = (l & 1) | ((l >> 1) & 2) | ((l >> 2) & 4)
        ly = ((l >> 1) & 1) | ((l >> 2) & 2)
        ox, oy = ((b & 3) * 8 + lx).reshape(-1), ((b >> 2) * 4 + ly).reshape(-1)
    else:
        j, i = np.meshgrid(np.arange(32), np.arange(32), indexing="ij")
        ox, oy = i.reshape(-1), j.reshape(-1)
    out = []
    for ty in range(tiles_y):
        for tx in range(tiles_x):
            if (tx + ty * k) % world != rank:
                continue
            x, y = tx * 32 + ox, ty * 32 + oy
            keep = (x < xres) & (y < yres)
            out.append(y[keep] * xres + x[keep])
    return np.concatenate(out).astype(np.int64) if out else np.zeros(0, np.int64)
