#!/bin/bash
mkdir -p gpurun_out
B2A_TRACE=1 timeout 150 python -X faulthandler -m pytest tests/test_gpu_shift_invert.py -q -s > gpurun_out/r2_debug_shift.log 2>&1
grep -v "^\[b2a\]   chunk" gpurun_out/r2_debug_shift.log | tail -30 | cut -c1-300
