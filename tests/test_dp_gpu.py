"""N-GPU data-parallel parity (needs >= 2 visible GPUs; skipped on a 1-GPU box): runs
scripts/dp_parity.py under torchrun/NCCL -- parameters after 3 steps on N ranks (one all-reduce per
step, 1/N folded into the fused AdamW) must equal the single-process result on the same N samples."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_rank_step_equals_single_process():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "dp_parity.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "dp_parity OK" in out.stdout
