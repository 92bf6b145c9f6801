#!/bin/bash
# round 2, 8-GPU call: parity at N = 8 (small + BASELINE size), bench at N = 4 / 8 (+ round-1 exchange at N = 8),
# BASELINE cfg 5 at its true size (north star), e2e phases at N = 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
timeout 240 $TR --nproc-per-node 8 --master-port 29701 tests/dist_gpu_check.py > gpurun_out/r2_dist_check_n8.log 2>&1
grep -E "world=|OK" gpurun_out/r2_dist_check_n8.log | cut -c1-300
timeout 400 $TR --nproc-per-node 8 --master-port 29702 tests/dist_gpu_check.py --big 1000000 > gpurun_out/r2_dist_big_n8.log 2>&1
grep -E "^\{|OK" gpurun_out/r2_dist_big_n8.log | cut -c1-700
summ() {
python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[2]) if l.startswith('{')][0])
    k=d['roofline']['kernels']
    print(sys.argv[1], 'ms/solve', round(d['ms_per_step'],2), 'value', round(d['value']), 'hbm_frac_agg', d['hbm_frac_aggregate'], 'e2e ms', round(d['e2e']['ms_per_step'],1), 'resid', d.get('residual_AQ_QR'), {n:(v['launches'], v['avg_us']) for n,v in k.items()})
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
timeout 200 $TR --nproc-per-node 8 --master-port 29703 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
summ "N=8 default" gpurun_out/r2_bench_n8.json
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 200 $TR --nproc-per-node 4 --master-port 29704 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
summ "N=4 default" gpurun_out/r2_bench_n4.json
B2A_XCHG=0 timeout 200 $TR --nproc-per-node 8 --master-port 29705 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8_push.json 2> gpurun_out/r2_bench_n8_push.err
summ "N=8 round-1 exchange" gpurun_out/r2_bench_n8_push.json
timeout 500 $TR --nproc-per-node 8 --master-port 29706 tools/cfg_dist_bench.py cfg5 --residual --restarts 200 > gpurun_out/r2_cfg5_n8.json 2> gpurun_out/r2_cfg5_n8.err
grep "^{" gpurun_out/r2_cfg5_n8.json | cut -c1-1800; tail -3 gpurun_out/r2_cfg5_n8.err | cut -c1-300
timeout 200 $TR --nproc-per-node 8 --master-port 29707 tools/e2e_phases.py > gpurun_out/r2_e2e_phases_n8.log 2>&1
grep "^{" gpurun_out/r2_e2e_phases_n8.log
