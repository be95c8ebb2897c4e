"""Frame-end gather of image tiles across ranks (one process per GPU, torch.distributed).

The render path needs no communication during a frame (pixels are independent, the scene is
replicated — SURVEY.md §8e). The single exchange step is the frame-end gather of every rank's
accumulated tiles to rank 0:

* backend "nccl": the tile-local device buffers of ``RenderCUDA`` are wrapped as torch CUDA
  tensors (zero copy) and moved with ``dist.gather`` over NVLink/NVSwitch; rank 0 scatters each
  gathered chunk into the full frame with the ``k_assemble`` kernel (``crtc_assemble_rank``).
* or no gather at all (``PeerFrame``): rank 0 exports its full-frame buffers as CUDA IPC handles, the other ranks
  map them, and every rank's frame-end resolve kernel stores its pixels straight into rank 0's frame over NVLink,
  followed by a completion flag written the same way; rank 0's stream waits for the flags. No collective per frame.
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


class FrameGatherer:
    """Frame-end gather of one renderer's tiles to rank ``dst`` (NCCL) and assembly there.

    ``submit()`` stages this rank's tile-local accum + img buffers (a device-to-device copy, so the
    next frame may overwrite them) and starts the gather asynchronously on NCCL's stream;
    ``finish()`` makes the renderer's stream wait for it and, on ``dst``, scatters every rank's
    chunk into the full frame (``k_assemble``). ``submit()`` of the next frame finishes the previous
    one first, so in a frame loop the transfer of frame f overlaps the rendering of frame f+1;
    call ``finish()`` right after ``submit()`` when the assembled frame is needed immediately.
    Everything is stream-ordered on the current torch stream (= the renderer's stream).
    """

    def __init__(self, renderer, group=None, dst: int = 0):
        import torch
        import torch.distributed as dist

        # `dst` is a rank of `group` (like self.rank); dist.gather wants the global rank
        self.r, self.group, self.dst = renderer, group, dst
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.dst_global = dist.get_global_rank(group, dst) if group is not None else dst
        self.dev = torch.device("cuda", renderer.device)
        # The staging copies and k_assemble are ordered against the frame only if they run on the renderer's stream:
        # make the current torch stream the renderer's (a default-constructed RenderCUDA owns a non-blocking stream
        # that the current torch stream would otherwise race)
        # (no CUDA runtime = the CPU dry run of tests/test_bench_contract.py over the emulated renderer)
        if torch.cuda.is_available():
            renderer.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
        accum_ptr, img_ptr, nloc = renderer.local_buffers()
        self.nloc = nloc
        max_tiles = tiles.max_local_tiles(renderer.width, renderer.height, self.world)
        a_bytes, i_bytes = tiles.TILE_PIXELS * 12, tiles.TILE_PIXELS * 4
        self.src_a = device_tensor(accum_ptr, nloc * a_bytes, self.dev) if nloc else None
        self.src_i = device_tensor(img_ptr, nloc * i_bytes, self.dev) if nloc else None
        self.send_a = torch.zeros(max_tiles * a_bytes, dtype=torch.uint8, device=self.dev)
        self.send_i = torch.zeros(max_tiles * i_bytes, dtype=torch.uint8, device=self.dev)
        self.recv_a = self.recv_i = None
        if self.rank == dst:
            self.recv_a = [torch.empty_like(self.send_a) for _ in range(self.world)]
            self.recv_i = [torch.empty_like(self.send_i) for _ in range(self.world)]
        self.work = None

    def submit(self):
        import torch.distributed as dist

        self.finish()
        if self.nloc:
            self.send_a[: self.src_a.numel()].copy_(self.src_a)
            self.send_i[: self.src_i.numel()].copy_(self.src_i)
        self.work = [dist.gather(self.send_a, self.recv_a, dst=self.dst_global, group=self.group, async_op=True),
                     dist.gather(self.send_i, self.recv_i, dst=self.dst_global, group=self.group, async_op=True)]

    def finish(self) -> bool:
        """Returns True on ``dst`` when a frame was assembled."""
        if self.work is None:
            return False
        for w in self.work:
            w.wait()  # the current stream waits for NCCL; no host synchronisation
        self.work = None
        if self.rank != self.dst:
            return False
        for r in range(self.world):
            self.r.assemble_rank(r, self.world, self.recv_a[r].data_ptr(), self.recv_i[r].data_ptr())
        return True


class PeerFrame:
    """Frame assembly WITHOUT a gather and without a collective: the resolve kernel of every rank writes its tiles directly
    into the assembling rank's full frame through peer-mapped memory (``crtc_export_frame`` / ``crtc_import_frame``;
    st.global over NVLink, fused into the kernel that produces the pixels) and then publishes a completion flag the same
    way; the assembling rank's stream waits for the flags (``crtc_frame_wait``). torch.distributed is used ONCE, to ship
    the 128 handle bytes. Same interface as ``FrameGatherer``: ``submit()`` does nothing (the data is already on its way
    when the frame's last kernel runs), ``finish()`` orders ``dst``'s stream after every rank's stores of the last frame —
    after it, ``read_accum`` / ``read_img`` on ``dst`` see the assembled frame (they also wait by themselves). Every
    rank must enqueue the same sequence of frames. Call again after ``initialize`` (a resize)."""

    def __init__(self, renderer, group=None, dst: int = 0):
        import torch
        import torch.distributed as dist

        self.r, self.group, self.dst = renderer, group, dst
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        # NCCL moves device tensors; with gloo (two processes sharing one GPU in the tests) the 128 bytes travel as a
        # host tensor
        nccl = dist.get_backend(group) == "nccl"
        handles = torch.zeros(128, dtype=torch.uint8, device=torch.device("cuda", renderer.device) if nccl else "cpu")
        if self.rank == dst:
            handles.copy_(torch.frombuffer(bytearray(renderer.export_frame()), dtype=torch.uint8))
        dist.broadcast(handles, src=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
        if self.rank != dst:
            renderer.import_frame(handles.cpu().numpy().tobytes())
        dist.barrier(group=group)  # nobody renders into the frame before everybody has mapped it
        self.pending = False

    def submit(self):
        self.pending = True

    def finish(self) -> bool:
        """Returns True on ``dst`` when a frame is complete there (stream-ordered, no host wait)."""
        if not self.pending:
            return False
        self.pending = False
        if self.rank != self.dst:
            return False
        self.r.frame_wait()
        return True


_GATHERERS = {}


def gather_frame_cuda(renderer, group=None, dst: int = 0) -> bool:
    """Blocking convenience form: gather this frame now and assemble it on ``dst`` (True there)."""
    key = (id(renderer), renderer.width, renderer.height)
    g = _GATHERERS.get(key)
    if g is None:
        _GATHERERS.clear()
        g = _GATHERERS[key] = FrameGatherer(renderer, group, dst)
    g.submit()
    return g.finish()


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
