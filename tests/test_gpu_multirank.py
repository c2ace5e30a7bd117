"""The N > 1 path on real hardware: world_size 2 (or more) over NCCL, one process per GPU under torchrun — sharded upload,
sharded pair stage, the three exchange steps and the final union-find — against the CPU oracle on every rank.  Skipped on a
box with one GPU (the single-GPU emulation of the same C-ABI stages is tests/test_gpu_sharded_emulated.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_run_under_nccl_matches_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "mg_worker.py")],
                         capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert out.returncode == 0 and "MG_WORKER_OK" in out.stdout, (out.stdout[-3000:], out.stderr[-3000:])
