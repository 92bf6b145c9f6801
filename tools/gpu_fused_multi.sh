#!/bin/bash
# 2 GPUs: parity of the fused sweep kernel on row-sharded workspaces (B2A_FUSED_SWEEP=2) and of the default path.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B2A_FUSED_SWEEP=2 timeout 150 $TR --master-port 29531 tests/dist_gpu_check.py > gpurun_out/multi_fused_check.log 2>&1
echo "dist check (fused) rc=$?"; grep -E "world=|DIST_GPU_CHECK_OK|AssertionError" gpurun_out/multi_fused_check.log | tail -6
timeout 150 $TR --master-port 29532 tests/dist_gpu_check.py > gpurun_out/multi_default_check.log 2>&1
echo "dist check (default) rc=$?"; grep -E "world=|DIST_GPU_CHECK_OK|AssertionError" gpurun_out/multi_default_check.log | tail -6
