#!/bin/bash
# round 2, second 8-GPU call: owner GROUPS (32 MB column blocks in arrival order) - parity, variants, cfg 5, bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 8 --master-port 29701 tests/dist_gpu_check.py > gpurun_out/r2d_dist_check_n8.log 2>&1
grep -E "world=|OK|Error" gpurun_out/r2d_dist_check_n8.log | cut -c1-300
: > gpurun_out/r2d_variants_n8.txt
port=29710
for v in "B2A_BENCH_X=default" "B2A_OWNER_BLOCKS=0" "B2A_OWNER_GROUP=2"; do
  port=$((port+1))
  env B2A_BENCH_QUICK=1 $v timeout 120 $TR --nproc-per-node 8 --master-port $port bench.py --gpus 8 --steps 6 --warmup 3 2>/dev/null | grep "^{" >> gpurun_out/r2d_variants_n8.txt
done
port=$((port+1))
CUDA_VISIBLE_DEVICES=0,1,2,3 B2A_BENCH_QUICK=1 B2A_OWNER_GROUP=2 timeout 120 $TR --nproc-per-node 4 --master-port $port bench.py --gpus 4 --steps 6 --warmup 3 2>/dev/null | grep "^{" >> gpurun_out/r2d_variants_n8.txt
port=$((port+1))
CUDA_VISIBLE_DEVICES=0,1,2,3 B2A_BENCH_QUICK=1 timeout 120 $TR --nproc-per-node 4 --master-port $port bench.py --gpus 4 --steps 6 --warmup 3 2>/dev/null | grep "^{" >> gpurun_out/r2d_variants_n8.txt
python - <<'PY'
import json
for l in open('gpurun_out/r2d_variants_n8.txt'):
    d=json.loads(l); e={k:v for k,v in d['env'].items() if k!='B2A_BENCH_QUICK'}
    print('N',d['n_gpus'],e,'ms/solve',round(d['ms_per_step'],2),'frac',d['hbm_frac_aggregate'],d['kernels_us'])
PY
timeout 500 $TR --nproc-per-node 8 --master-port 29733 tools/cfg_dist_bench.py cfg5 --residual --restarts 200 > gpurun_out/r2d_cfg5_n8.json 2> gpurun_out/r2d_cfg5_n8.err
grep "^{" gpurun_out/r2d_cfg5_n8.json | cut -c1-1300; tail -3 gpurun_out/r2d_cfg5_n8.err | cut -c1-300
timeout 200 $TR --nproc-per-node 8 --master-port 29731 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2d_bench_n8.json 2> gpurun_out/r2d_bench_n8.err
grep "^{" gpurun_out/r2d_bench_n8.json | cut -c1-400
