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


def gather_frame_cuda(renderer, group=None, dst: int = 0):
    """NCCL gather of this rank's tile-local accum + img buffers to ``dst`` and assembly there.

    Returns True on ``dst`` (full frame now readable with ``renderer.read_accum()/read_img()``).
    All ranks must call. Pads to the largest per-rank tile count so chunks are uniform.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = torch.device("cuda", renderer.device)
    accum_ptr, img_ptr, nloc = renderer.local_buffers()
    max_tiles = tiles.max_local_tiles(renderer.width, renderer.height, world)
    a_bytes, i_bytes = tiles.TILE_PIXELS * 12, tiles.TILE_PIXELS * 4
    send = torch.zeros(max_tiles * (a_bytes + i_bytes), dtype=torch.uint8, device=dev)
    if nloc:
        send[: nloc * a_bytes].copy_(device_tensor(accum_ptr, nloc * a_bytes, dev))
        send[max_tiles * a_bytes: max_tiles * a_bytes + nloc * i_bytes].copy_(device_tensor(img_ptr, nloc * i_bytes, dev))
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return False
    torch.cuda.current_stream(dev).synchronize()
    for r in range(world):
        base = recv[r].data_ptr()
        renderer.assemble_rank(r, world, base, base + max_tiles * a_bytes)
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
