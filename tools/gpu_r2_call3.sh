#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_gpu_tests_b.log
timeout 300 python tools/spmvbench.py 1e6 16 --esweep > gpurun_out/r2_spmv_esweep_cfg2.log 2>&1
SWEEP_LPR=2,4 timeout 300 python tools/spmvbench.py 160 --laplace --esweep > gpurun_out/r2_spmv_esweep_lap160.log 2>&1
SWEEP_LPR=4,8 timeout 300 python tools/spmvbench.py 2e6 20 --complex --esweep > gpurun_out/r2_spmv_esweep_c64.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -3 gpurun_out/r2_gpu_tests_b.log; cat gpurun_out/r2_spmv_esweep_cfg2.log; cat gpurun_out/r2_spmv_esweep_lap160.log; cat gpurun_out/r2_spmv_esweep_c64.log; head -c 600 gpurun_out/r2_bench_b.json
