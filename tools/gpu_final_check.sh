#!/bin/bash
# Round-end validation on one B200: full GPU test-suite, bench, ncu launch list + full captures, sanitizer.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/final_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/final_gpu_tests.log
timeout 200 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/final_bench.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv \
  --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:cgs_sweep -s 38 -c 2 \
  -o gpurun_out/prof_cgs_sweep -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_sweep.log 2>&1
echo "ncu sweep rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:spmv_csr -s 30 -c 1 \
  -o gpurun_out/prof_spmv -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_spmv.log 2>&1
echo "ncu spmv rc=$?"
timeout 150 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/final_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/final_memcheck.log
timeout 150 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/final_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/final_racecheck.log
ls -la gpurun_out | head -30
