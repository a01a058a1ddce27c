"""GPU, >= 2 devices: the data-parallel MLE step with the bucketed, overlapped all-reduce (tgan_allreduce_bucket, NCCL
bound at run time) against the flat all-reduce and the single-process run.  Skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_bucketed_allreduce_matches_flat_and_single_process():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "dp_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DP_WORKER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
