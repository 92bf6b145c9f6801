"""Runs tests/dist_gpu_check.py under torchrun when at least two GPUs are visible."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


VARIANTS = {
    # default: staged copy-engine exchange + owner-blocked mat-vec + fused orthogonalisation kernel
    "default": {},
    "staged_four_kernels": {"B2A_FUSED_SWEEP": "0"},
    "staged_one_side_stream": {"B2A_XCHG_STREAMS": "1"},
    "staged_without_owner_blocks": {"B2A_OWNER_BLOCKS": "0"},
    # round-1 exchange: x pushed from inside the normalising kernel
    "push_fused_sweep": {"B2A_XCHG": "0", "B2A_FUSED_SWEEP": "2"},
    "push_four_kernels": {"B2A_XCHG": "0", "B2A_FUSED_SWEEP": "0"},
    "nccl_collectives": {"B2A_NO_PEER": "1"},
}


def run_check(world, port, env_extra, args=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gpu_check.py"), *args]
    env = dict(os.environ, **env_extra)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_two_gpu_row_sharded_parity(variant):
    """H, Schur vectors, eigenvalues and breakdown handling of the row-sharded path against the oracle on the whole
    matrix, for every exchange / orthogonalisation variant (tests/dist_gpu_check.py)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    res = run_check(2, 29517 + sorted(VARIANTS).index(variant), VARIANTS[variant])
    assert "DIST_GPU_CHECK_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_baseline_size_parity_on_all_gpus(world):
    """bench.py's matrix at 1e6 rows per GPU on `world` GPUs against the oracle on the whole matrix."""
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs >= {world} GPUs")
    res = run_check(world, 29540 + world, {}, ("--big", "1000000"))
    assert "DIST_GPU_BIG_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
