#!/bin/bash
# round 2, GPU call 1 (one GPU): parity suite, FP64 probe, rotation micro-benchmark, bench line, ncu of the rotation
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_gpu_tests.log
timeout 60 tools/fp64_peak.bin > gpurun_out/r2_fp64_peak.json 2>&1
FP=$(python -c "import json;print(json.load(open('gpurun_out/r2_fp64_peak.json'))['dmma_tflops'])" 2>/dev/null || echo 40)
timeout 300 python tools/rotbench.py --fp64-peak $FP --json gpurun_out/r2_rotbench.json > gpurun_out/r2_rotbench.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rotate_mma -s 2 -c 1 -f -o gpurun_out/r2_rot_mma python tools/rotbench.py --only 0 > gpurun_out/r2_ncu_rot.log 2>&1
tail -5 gpurun_out/r2_gpu_tests.log; cat gpurun_out/r2_fp64_peak.json; cat gpurun_out/r2_rotbench.log | tail -12
