#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rotate_mma.py tests/test_gpu_kernels.py -m gpu -x -q -k "rotate" 2>&1 | tail -5 > gpurun_out/r2_rot_tests.log
for c in 1 2; do
  B2A_ROT_CTAS=$c timeout 300 python tools/rotbench.py --fp64-peak 37.1 --json gpurun_out/r2_rotbench_ctas$c.json > gpurun_out/r2_rotbench_ctas$c.log 2>&1
done
tail -3 gpurun_out/r2_rot_tests.log; grep dmma gpurun_out/r2_rotbench_ctas1.log; echo; grep dmma gpurun_out/r2_rotbench_ctas2.log
