"""Multi-GPU training step (SURVEY.md 4 "Distributed"): parameters after one data-parallel step == the oracle's Adam on
the mean of the per-rank oracle gradients, replicas bit-identical, for every all-reduce mode (NCCL, fused P2P kernel,
NVLS).  Needs >= 2 GPUs (skipped on a single-GPU box); launched through torchrun on 127.0.0.1."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_data_parallel_step_matches_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(root, "tests", "_dist_train_worker.py")]
    p = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and p.stdout.count("DIST_OK") == n, (p.stdout[-3000:], p.stderr[-3000:])
