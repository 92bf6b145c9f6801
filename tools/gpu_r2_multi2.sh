#!/bin/bash
# round 2, 2-GPU call: parity of every exchange variant, BASELINE-size parity, bench variants, e2e phases
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_multi2_tests.log
tail -4 gpurun_out/r2_multi2_tests.log
i=0
for v in "" "B2A_OWNER_FUSED=0" "B2A_AR_LL=0" "B2A_XCHG_STREAMS=1" "B2A_OWNER_BLOCKS=0" "B2A_XCHG=0"; do
  i=$((i+1))
  env $v timeout 150 $TR --master-port $((29610+i)) bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2b_bench_n2_v$i.json 2> gpurun_out/r2b_bench_n2_v$i.err
  python - "$v" gpurun_out/r2b_bench_n2_v$i.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[2]) if l.startswith('{')][0])
    k=d['roofline']['kernels']
    print(sys.argv[1] or 'default', 'ms/solve', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'resid', d.get('residual_AQ_QR'), {n:(v['launches'], v['avg_us']) for n,v in k.items()})
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
timeout 200 $TR --master-port 29641 tools/e2e_phases.py > gpurun_out/r2_e2e_phases_n2.log 2>&1
grep "^{" gpurun_out/r2_e2e_phases_n2.log
