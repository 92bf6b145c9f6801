#!/bin/bash
mkdir -p gpurun_out
for n in 1e6 2e6 4e6 8e6; do
  for mb in 0 32 16; do
    echo "n=$n B2A_SPMV_BLOCK_MB=$mb" 
    B2A_SPMV_BLOCK_MB=$mb timeout 200 python tools/spmvbench.py $n 16 2>&1 | tail -1
  done
done > gpurun_out/r2_spmv_xsize.log 2>&1
cat gpurun_out/r2_spmv_xsize.log
