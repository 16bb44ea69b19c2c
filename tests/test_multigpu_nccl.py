"""Multi-GPU (NCCL) correctness of the data-parallel step, as a -m gpu test: skipped on a one-GPU box, run under
torchrun on two GPUs otherwise (tests/multigpu/check_nccl_step.py: sharded pairs + in-backward all-reduce on the
packed gradient table == full-batch gradient / world; bounded by-rows route; owner-computes optimizer step).
The same check runs inside `bench.py --gpus N` (its `parity_check` record), which is what the driver's scaling
run sees; the host-side logic is covered on CPU with gloo in tests/test_distributed_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_nccl_step_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "multigpu", "check_nccl_step.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "nccl step check ok on 2 GPUs" in r.stdout
