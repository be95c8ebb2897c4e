"""The N>1 path on CPU: world_size-2 gloo processes shard a frame by tiles, gather to rank 0
and assemble — the same plumbing the NCCL path uses (chameleonrt_b200/distributed.py)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np
    import torch.distributed as dist
    from chameleonrt_b200 import tiles
    from chameleonrt_b200.distributed import gather_frame_numpy
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% sys.argv[1], rank=int(sys.argv[2]), world_size=2)
    w, h = 200, 130   # ragged: 4x3 tiles, the last row/column partial, 12 tiles -> 6 per rank
    rank = dist.get_rank()
    yy, xx = np.mgrid[0:h, 0:w]
    full_accum = np.stack([xx * 1.0, yy * 2.0, xx * 0.5 + yy], -1).astype(np.float32)
    full_img = (xx + yy * w).astype(np.uint32)
    local_accum = tiles.to_local(full_accum, rank, 2)
    local_img = tiles.to_local(full_img, rank, 2)
    assert len(local_accum) == len(tiles.local_tile_ids(w, h, rank, 2)) * 4096
    out = gather_frame_numpy(local_accum, local_img, w, h)
    if rank == 0:
        accum, img = out
        assert (accum == full_accum).all() and (img == full_img).all()
        print("GATHER_OK")
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()
""") % ROOT


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [subprocess.Popen([sys.executable, str(script), str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]
