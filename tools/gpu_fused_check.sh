#!/bin/bash
# One gpurun call: parity of the fused sweep kernel, then the bench with and without it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout 420 python -m pytest tests/test_gpu_fused_sweep.py -x -q > gpurun_out/fused_tests.log 2>&1
echo "fused tests rc=$?"; tail -5 gpurun_out/fused_tests.log
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default rc=$?"; cut -c1-400 gpurun_out/bench_default.json
B2A_FUSED_SWEEP=1 timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err
echo "bench fused rc=$?"; cut -c1-400 gpurun_out/bench_fused.json; tail -3 gpurun_out/bench_fused.err
