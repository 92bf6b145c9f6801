#!/bin/bash
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_gpu_tests_final.log
tail -4 gpurun_out/r2_gpu_tests_final.log
