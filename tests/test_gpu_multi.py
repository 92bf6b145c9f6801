"""Runs tests/dist_gpu_check.py under torchrun when at least two GPUs are visible."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fused_sweep,port", [("1", "29517"), ("2", "29518"), ("0", "29519")],
                         ids=["default", "fused_sweep_on_shards", "four_kernels"])
def test_two_gpu_row_sharded_parity(fused_sweep, port):
    """default: four-kernel chain with in-kernel collectives over NVLink peer memory; B2A_FUSED_SWEEP=2: the fused
    orthogonalisation kernel with the all-reduces inside its grid barriers (both verified on 2 x B200)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    env = dict(os.environ, B2A_FUSED_SWEEP=fused_sweep)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert "DIST_GPU_CHECK_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
