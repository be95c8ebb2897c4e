"""Launched by torchrun (one process per GPU): renders a frame sharded by tiles over all ranks, assembles it on rank 0
and checks it is bit-identical to a single-GPU render of the same frame (SURVEY.md §8e determinism check).
  default       tiles gathered over NCCL, scattered by k_assemble (FrameGatherer)
  --peer        no gather: every rank's resolve kernel writes into rank 0's frame through CUDA IPC (PeerFrame)
  --one-gpu     all ranks on cuda:0 with the gloo backend (NCCL refuses two ranks per GPU): lets a single-GPU box
                exercise the cross-process peer mapping; implies --peer"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from chameleonrt_b200 import ArcballCamera, RenderCUDA
from chameleonrt_b200.distributed import FrameGatherer, PeerFrame
from chameleonrt_b200.scenes import sponza_like

ap = argparse.ArgumentParser()
ap.add_argument("width", type=int, nargs="?", default=640)
ap.add_argument("height", type=int, nargs="?", default=360)
ap.add_argument("--peer", action="store_true")
ap.add_argument("--one-gpu", action="store_true")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
if args.one_gpu:
    local, args.peer = 0, True
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if args.one_gpu:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
w, h = args.width, args.height
scene, cam = sponza_like(spp=2, detail=0.4, tex_size=128)
c = ArcballCamera(cam["eye"], cam["center"], cam["up"])
r = RenderCUDA(local, max_depth=5, rank=rank, world_size=world, stream=stream.cuda_stream)
r.initialize(w, h)
r.set_scene(scene)
gatherer = PeerFrame(r) if args.peer else FrameGatherer(r)
for f in range(3):
    st = r.render(c.eye(), c.dir(), c.up(), cam["fov_y"], f == 0, False)
    gatherer.submit()  # FrameGatherer: frame f's transfer overlaps frame f+1's rendering; PeerFrame: nothing to do
gatherer.finish()
t = torch.tensor([float(st.num_rays)], dtype=torch.float64, device="cpu" if args.one_gpu else dev)
dist.all_reduce(t)
if rank == 0:
    got_accum, got_img = r.read_accum(), r.read_img()
    single = RenderCUDA(local, max_depth=5, stream=stream.cuda_stream)
    single.initialize(w, h)
    single.set_scene(scene)
    for f in range(3):
        s1 = single.render(c.eye(), c.dir(), c.up(), cam["fov_y"], f == 0, True)
    ok = (got_accum.view(np.uint32) == single.read_accum().view(np.uint32)).all() and (got_img == single.read_img()).all()
    ok = ok and int(t.item()) == s1.num_rays
    print("MGPU_OK" if ok else "MGPU_MISMATCH", world, "peer" if args.peer else "gather", int(t.item()), s1.num_rays, flush=True)
dist.barrier()
if args.peer and rank != 0:
    r.import_frame(None)  # unmap before rank 0 frees the frame
dist.barrier()
dist.destroy_process_group()
