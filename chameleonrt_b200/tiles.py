"""Image-tile sharding across GPUs: host-side layout logic.

The reference already splits the image into independent 64x64 tiles across CPU threads
(backends/embree/render_embree.cpp:172-192, tile id = ty * ntiles_x + tx). The CUDA backend
shards the same tiles across ranks — rank r owns the tiles with ``tile_id % world_size == r`` —
and every rank keeps its tiles' accumulation resident in a tile-local buffer of
``num_local_tiles * 4096`` pixels. Inside a tile, pixels are ordered in 8x4 blocks (one warp =
one 8x4 screen rectangle). This module restates that mapping in numpy so the gather/assembly
plumbing can be tested on CPU (gloo) and so hosts can de-tile gathered buffers without a GPU.
It must agree with ``local_pixel_coords`` in csrc/kernels.cuh.
"""
from __future__ import annotations

import numpy as np

TILE = 64
TILE_PIXELS = TILE * TILE


def num_tiles(fb_width: int, fb_height: int):
    ntx = fb_width // TILE + (1 if fb_width % TILE else 0)
    nty = fb_height // TILE + (1 if fb_height % TILE else 0)
    return ntx, nty


def local_tile_ids(fb_width: int, fb_height: int, rank: int, world_size: int) -> np.ndarray:
    ntx, nty = num_tiles(fb_width, fb_height)
    t = np.arange(ntx * nty, dtype=np.uint32)
    return t[t % world_size == rank]


def max_local_tiles(fb_width: int, fb_height: int, world_size: int) -> int:
    ntx, nty = num_tiles(fb_width, fb_height)
    return (ntx * nty + world_size - 1) // world_size


def local_pixel_coords(fb_width: int, fb_height: int, rank: int, world_size: int):
    """(x, y, valid) for every local pixel index of ``rank``."""
    ntx, _ = num_tiles(fb_width, fb_height)
    tiles = local_tile_ids(fb_width, fb_height, rank, world_size)
    lp = np.arange(len(tiles) * TILE_PIXELS, dtype=np.uint32)
    lt, within = lp // TILE_PIXELS, lp % TILE_PIXELS
    tile = tiles[lt] if len(tiles) else lt
    blk, lane = within >> 5, within & 31
    x = (tile % ntx) * TILE + (blk & 7) * 8 + (lane & 7)
    y = (tile // ntx) * TILE + (blk >> 3) * 4 + (lane >> 3)
    valid = (x < fb_width) & (y < fb_height)
    return x.astype(np.int64), y.astype(np.int64), valid


def to_local(full: np.ndarray, rank: int, world_size: int) -> np.ndarray:
    """Row-major (h, w, ...) frame -> this rank's tile-local buffer (npx_local, ...)."""
    h, w = full.shape[:2]
    x, y, valid = local_pixel_coords(w, h, rank, world_size)
    out = np.zeros((len(x),) + full.shape[2:], dtype=full.dtype)
    out[valid] = full[y[valid], x[valid]]
    return out


def assemble(chunks, fb_width: int, fb_height: int, world_size: int) -> np.ndarray:
    """Inverse of ``to_local``: chunks[r] is rank r's tile-local buffer (extra padding ignored)."""
    first = np.asarray(chunks[0])
    full = np.zeros((fb_height, fb_width) + first.shape[1:], dtype=first.dtype)
    for r in range(world_size):
        x, y, valid = local_pixel_coords(fb_width, fb_height, r, world_size)
        c = np.asarray(chunks[r])[: len(x)]
        full[y[valid], x[valid]] = c[valid]
    return full
