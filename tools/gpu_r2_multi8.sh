#!/bin/bash
# round 2, 8-GPU call: parity at N = 8 (small + BASELINE size), exchange variants at N = 4 / 8, full bench lines,
# BASELINE cfg 5 at its true size (north star), e2e phases at N = 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
timeout 240 $TR --nproc-per-node 8 --master-port 29701 tests/dist_gpu_check.py > gpurun_out/r2_dist_check_n8.log 2>&1
grep -E "world=|OK" gpurun_out/r2_dist_check_n8.log | cut -c1-300
timeout 300 $TR --nproc-per-node 8 --master-port 29702 tests/dist_gpu_check.py --big 1000000 --dump gpurun_out/r2_big_n8.npz > gpurun_out/r2_dist_big_n8.log 2>&1
grep -E "^\{|DUMPED|Error|error" gpurun_out/r2_dist_big_n8.log | cut -c1-700
: > gpurun_out/r2_variants_n48.txt
port=29710
for N in 8; do
  for v in "B2A_OWNER_FUSED=1" "B2A_OWNER_FUSED=0" "B2A_OWNER_BLOCKS=0" "B2A_XCHG=0"; do
    port=$((port+1))
    env B2A_BENCH_QUICK=1 $v timeout 120 $TR --nproc-per-node $N --master-port $port bench.py --gpus $N --steps 6 --warmup 3 2>/dev/null | grep "^{" >> gpurun_out/r2_variants_n48.txt
  done
done
BEST=$(python - <<'PY'
import json,sys
best=None
for l in open('gpurun_out/r2_variants_n48.txt'):
    d=json.loads(l); e={k:v for k,v in d['env'].items() if k!='B2A_BENCH_QUICK'}
    print('N',d['n_gpus'],e,'ms/solve',round(d['ms_per_step'],2),'frac',d['hbm_frac_aggregate'],d['kernels_us'], file=sys.stderr)
    if d.get('converged') and (best is None or d['ms_per_step']<best[0]): best=(d['ms_per_step'],e)
print(' '.join(f'{k}={v}' for k,v in (best[1] if best else {}).items()))
PY
)
echo "best variant at N=8: $BEST"
export $BEST
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
timeout 200 $TR --nproc-per-node 8 --master-port 29731 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
summ "N=8 default" gpurun_out/r2_bench_n8.json
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 200 $TR --nproc-per-node 4 --master-port 29732 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
summ "N=4 default" gpurun_out/r2_bench_n4.json
timeout 500 $TR --nproc-per-node 8 --master-port 29733 tools/cfg_dist_bench.py cfg5 --residual --restarts 200 > gpurun_out/r2_cfg5_n8.json 2> gpurun_out/r2_cfg5_n8.err
grep "^{" gpurun_out/r2_cfg5_n8.json | cut -c1-1800; tail -3 gpurun_out/r2_cfg5_n8.err | cut -c1-300
timeout 200 $TR --nproc-per-node 8 --master-port 29735 tools/e2e_phases.py > gpurun_out/r2_e2e_phases_n8.log 2>&1
grep "^{" gpurun_out/r2_e2e_phases_n8.log
