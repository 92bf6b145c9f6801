#!/bin/bash
# 2 GPUs: parity of the fused sweep kernel on row-sharded workspaces (B2A_FUSED_SWEEP=2), then the bench both ways.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B2A_FUSED_SWEEP=2 timeout 150 $TR --master-port 29531 tests/dist_gpu_check.py > gpurun_out/multi_fused_check.log 2>&1
echo "dist check (fused) rc=$?"; grep -E "world=|DIST_GPU_CHECK_OK|Error|error" gpurun_out/multi_fused_check.log | tail -6
timeout 120 $TR --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench2_default.json 2> gpurun_out/bench2_default.err
echo "bench2 default rc=$?"; cut -c1-330 gpurun_out/bench2_default.json
B2A_FUSED_SWEEP=2 timeout 120 $TR --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench2_fused.json 2> gpurun_out/bench2_fused.err
echo "bench2 fused rc=$?"; cut -c1-330 gpurun_out/bench2_fused.json; tail -2 gpurun_out/bench2_fused.err
