"""The multi-GPU check (tests/multi_gpu_check.py: slice equivalence, the statistics exchange through NCCL, the library's own
NVLink peer-memory kernel and the in-kernel publisher of the PD law) as a pytest case: launched under torchrun on two GPUs when
the box has them, skipped on a one-GPU box (the driver's GPU tier)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_slices_and_statistics_exchange():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "multi-GPU check OK on 2 GPUs" in r.stdout
