"""Frame-end gather of image tiles across ranks (one process per GPU, torch.distributed).

The render path needs no communication during a frame (pixels are independent, the scene is
replicated — SURVEY.md §8e). The single exchange step is the frame-end gather of every rank's
accumulated tiles to rank 0:

* backend "nccl": the tile-local device buffers of ``RenderCUDA`` are wrapped as torch CUDA
  tensors (zero copy) and moved with ``dist.gather`` over NVLink/NVSwitch; rank 0 scatters each
  gathered chunk into the full frame with the ``k_assemble`` kernel (``crtc_assemble_rank``).
* backend "gloo" (CPU tests): the same plumbing with numpy buffers and ``tiles.assemble``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import tiles


class _DevPtr:
    """Minimal __cuda_array_interface__ holder so torch can alias a raw device pointer."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2, "strides": None,
        }


def device_tensor(ptr: int, nbytes: int, device):
    import torch

    return torch.as_tensor(_DevPtr(ptr, nbytes), device=device)


_GATHER_CACHE = {}


def gather_frame_cuda(renderer, group=None, dst: int = 0):
    """NCCL gather of this rank's tile-local accum + img buffers to ``dst`` and assembly there.

    Returns True on ``dst`` (full frame now readable with ``renderer.read_accum()/read_img()``).
    All ranks must call; everything is stream-ordered on the current torch stream (which must be
    the renderer's stream), no host synchronisation. When every rank owns the same number of
    tiles (e.g. 1280x720: 240 tiles over 1/2/4/8 ranks) the renderer's buffers are sent in place;
    otherwise they are padded to the largest per-rank tile count.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = torch.device("cuda", renderer.device)
    accum_ptr, img_ptr, nloc = renderer.local_buffers()
    max_tiles = tiles.max_local_tiles(renderer.width, renderer.height, world)
    ntx, nty = tiles.num_tiles(renderer.width, renderer.height)
    uniform = (ntx * nty) % world == 0
    a_bytes, i_bytes = tiles.TILE_PIXELS * 12, tiles.TILE_PIXELS * 4
    key = (id(renderer), accum_ptr, img_ptr, world, max_tiles)
    st = _GATHER_CACHE.get(key)
    if st is None:
        st = {}
        if uniform:
            st["send_a"] = device_tensor(accum_ptr, nloc * a_bytes, dev)
            st["send_i"] = device_tensor(img_ptr, nloc * i_bytes, dev)
        else:
            st["send_a"] = torch.zeros(max_tiles * a_bytes, dtype=torch.uint8, device=dev)
            st["send_i"] = torch.zeros(max_tiles * i_bytes, dtype=torch.uint8, device=dev)
            st["src_a"] = device_tensor(accum_ptr, nloc * a_bytes, dev) if nloc else None
            st["src_i"] = device_tensor(img_ptr, nloc * i_bytes, dev) if nloc else None
        if rank == dst:
            st["recv_a"] = [torch.empty(max_tiles * a_bytes, dtype=torch.uint8, device=dev) for _ in range(world)]
            st["recv_i"] = [torch.empty(max_tiles * i_bytes, dtype=torch.uint8, device=dev) for _ in range(world)]
        _GATHER_CACHE.clear()
        _GATHER_CACHE[key] = st
    if not uniform and nloc:
        st["send_a"][: nloc * a_bytes].copy_(st["src_a"])
        st["send_i"][: nloc * i_bytes].copy_(st["src_i"])
    dist.gather(st["send_a"], st.get("recv_a"), dst=dst, group=group)
    dist.gather(st["send_i"], st.get("recv_i"), dst=dst, group=group)
    if rank != dst:
        return False
    for r in range(world):
        renderer.assemble_rank(r, world, st["recv_a"][r].data_ptr(), st["recv_i"][r].data_ptr())
    return True


def gather_frame_numpy(local_accum: np.ndarray, local_img: np.ndarray, fb_width: int, fb_height: int,
                       group=None, dst: int = 0):
    """gloo twin of ``gather_frame_cuda`` for CPU tests: returns (accum, img) on dst, else None."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    max_px = tiles.max_local_tiles(fb_width, fb_height, world) * tiles.TILE_PIXELS
    a = np.zeros((max_px, 3), np.float32)
    i = np.zeros((max_px,), np.int32)
    a[: len(local_accum)] = local_accum
    i[: len(local_img)] = local_img.view(np.int32)
    ta, ti = torch.from_numpy(a), torch.from_numpy(i)
    ra = [torch.empty_like(ta) for _ in range(world)] if rank == dst else None
    ri = [torch.empty_like(ti) for _ in range(world)] if rank == dst else None
    dist.gather(ta, ra, dst=dst, group=group)
    dist.gather(ti, ri, dst=dst, group=group)
    if rank != dst:
        return None
    accum = tiles.assemble([t.numpy() for t in ra], fb_width, fb_height, world)
    img = tiles.assemble([t.numpy().view(np.uint32) for t in ri], fb_width, fb_height, world)
    return accum, img
