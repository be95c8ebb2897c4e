"""N > 1 on real GPUs (skipped unless the box has at least two): tile sharding + NCCL gather is
bit-identical to one GPU. The CPU/gloo twin of this plumbing is tests/test_distributed_cpu.py."""
import os
import socket
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("size,peer", [((640, 360), False), ((200, 130), False), ((640, 360), True)])
def test_tile_sharded_render_matches_single_gpu(built, size, peer):
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "scripts", "mgpu_check.py"), str(size[0]), str(size[1])] + (["--peer"] if peer else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_OK" in r.stdout, r.stdout[-2000:]
