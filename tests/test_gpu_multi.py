"""The N > 1 path on HARDWARE: torchrun with one rank per GPU over NCCL (skipped on a one-GPU box; run it with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).  The gloo run of tests/test_sharding_gloo.py covers
the same partitioning logic on CPU."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_solve_and_device_gather_over_nccl():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip(f"needs >= 2 GPUs for an NCCL run (found {ngpu})")
    world = 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tests", "multi_gpu_worker.py")], capture_output=True, text=True, timeout=900,
                       cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"MULTI_GPU_OK {world}" in r.stdout
