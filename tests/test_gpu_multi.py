"""The NCCL tile assembly on real GPUs (needs at least two; skipped on a one-GPU box, where
tests/test_gpu_render.py::test_packed_tiles_gather_and_scatter covers the device halves and
tests/test_multirank_cpu.py the protocol): rt_render_multi and rayito_b200::raytraceMulti() over
all visible GPUs give rank 0 the single-GPU image bit for bit."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tile_assembly_over_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible: the NCCL gather needs two or more")
    n = min(n, 8)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "multi_check.py")]
    proc = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(proc.stdout[-4000:])
    assert proc.returncode == 0 and "MULTI_CHECK_OK" in proc.stdout
