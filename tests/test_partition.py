"""Tile ownership: the library's host-only entry point vg_owned_pixels against its Python twin, and the
properties the multi-GPU gather relies on. No GPU needed."""
import numpy as np
import pytest


@pytest.mark.parametrize("xres,yres", [(1920, 1080), (3840, 2160), (200, 140), (33, 70), (512, 512)])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 6, 8])
def test_owned_pixels_partition_the_frame(built_library, xres, yres, world):
    from vermeer_b200.host import owned_pixels as lib_owned
    from vermeer_b200.partition import owned_pixels as py_owned
    seen = np.zeros(xres * yres, np.int32)
    counts = []
    for rank in range(world):
        a = lib_owned(xres, yres, rank, world)
        assert np.array_equal(a, py_owned(xres, yres, rank, world)), (rank, world)
        seen[a] += 1
        counts.append(len(a))
        # the same SET whatever the order inside a tile
        assert np.array_equal(np.sort(a), np.sort(lib_owned(xres, yres, rank, world, pixel_block=False)))
    assert (seen == 1).all()                                      # disjoint and complete
    if xres * yres >= 512 * 512:
        assert max(counts) - min(counts) <= 0.15 * (xres * yres / world) + 2 * 1024   # balanced up to a couple of tiles


@pytest.mark.parametrize("world", [2, 3, 5, 6, 8])
def test_no_rank_owns_whole_tile_columns(built_library, world):
    """The row-to-row skew is coprime with the world size: a rank's tiles never line up in columns (advisor finding, round 1)."""
    from vermeer_b200.host import owned_pixels as lib_owned
    from vermeer_b200.partition import tile_stride
    for tiles_x in (30, 60, 15, 9, 10):
        xres, yres = tiles_x * 32, 8 * 32
        k = tile_stride(tiles_x, world)
        assert np.gcd(k, world) == 1 and k >= tiles_x
        a = lib_owned(xres, yres, 0, world)
        tx, ty = (a % xres) // 32, (a // xres) // 32
        cols_row0 = set(tx[ty == 0].tolist())
        cols_row1 = set(tx[ty == 1].tolist())
        if tiles_x >= world:
            assert cols_row0 != cols_row1


def test_morton_blocks_are_compact(built_library):
    """Inside a tile the path order is 8x4 blocks in Morton order: 2, 4, 8, 16 and 32 consecutive pixels form 2x1, 2x2, 4x2, 4x4
    and 8x4 blocks, which is what lets a warp of B pixels x 32/B iterations stay compact (render.cu: path_index)."""
    from vermeer_b200.host import owned_pixels as lib_owned
    a = lib_owned(64, 64, 0, 1)
    x, y = a % 64, a // 64
    for n, (w, h) in {2: (2, 1), 4: (2, 2), 8: (4, 2), 16: (4, 4), 32: (8, 4)}.items():
        for g in range(0, len(a), n):
            assert x[g:g + n].max() - x[g:g + n].min() == w - 1 and y[g:g + n].max() - y[g:g + n].min() == h - 1
