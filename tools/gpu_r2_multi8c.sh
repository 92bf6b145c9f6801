#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > gpurun_out/r2e_variants_n8.txt
port=29810
for v in "B2A_XCHG_STREAMS=7" "B2A_XCHG_STREAMS=3" "B2A_XCHG_STREAMS=1" "B2A_XCHG_STREAMS=7 B2A_OWNER_BLOCKS=0"; do
  port=$((port+1))
  env B2A_BENCH_QUICK=1 $v timeout 100 $TR --nproc-per-node 8 --master-port $port bench.py --gpus 8 --steps 5 --warmup 2 2>/dev/null | grep "^{" >> gpurun_out/r2e_variants_n8.txt
done
python - <<'PY'
import json
for l in open('gpurun_out/r2e_variants_n8.txt'):
    d=json.loads(l); e={k:v for k,v in d['env'].items() if k!='B2A_BENCH_QUICK'}
    print('N',d['n_gpus'],e,'ms/solve',round(d['ms_per_step'],2),'frac',d['hbm_frac_aggregate'],d['kernels_us'])
PY
