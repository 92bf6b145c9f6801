#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B2A_OWNER_GROUP=1 timeout 240 $TR --master-port 29801 tests/dist_gpu_check.py > gpurun_out/r2c_dist_check_n2_groups.log 2>&1
grep -E "world=|OK|Error|error" gpurun_out/r2c_dist_check_n2_groups.log | cut -c1-300
timeout 240 $TR --master-port 29802 tests/dist_gpu_check.py > gpurun_out/r2c_dist_check_n2_default.log 2>&1
grep -E "world=|OK|Error|error" gpurun_out/r2c_dist_check_n2_default.log | cut -c1-300
for v in "B2A_OWNER_GROUP=1" "B2A_BENCH_X=1"; do
  env B2A_BENCH_QUICK=1 $v timeout 120 $TR --master-port 29803 bench.py --gpus 2 --steps 6 --warmup 3 2>/dev/null | grep "^{" | cut -c1-700
done
