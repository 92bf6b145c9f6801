#!/bin/bash
# One gpurun call: parity of the fused sweep kernel, then sweep timings of the variants.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused_sweep.py -x -q > gpurun_out/fused_tests.log 2>&1
echo "fused tests rc=$?"; tail -5 gpurun_out/fused_tests.log
timeout 400 python tools/sweepbench.py 2>&1 | tee gpurun_out/sweepbench2.log
