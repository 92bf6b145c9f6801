#!/bin/bash
# round 2, final single-GPU evidence: parity suite, smoke, bench line, launch list, ncu captures
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_gpu_tests_final.log
tail -12 gpurun_out/r2_gpu_tests_final.log
timeout 100 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
head -c 200 gpurun_out/r2_bench_final.json; echo
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
timeout 80 ncu --set full --clock-control none --import-source on -k regex:rotate_mma -s 2 -c 1 -f -o gpurun_out/r2_rot_mma_final python tools/rotbench.py --only 0 > /dev/null 2>&1
timeout 80 ncu --set full --clock-control none --import-source on -k regex:spmv_csr_vector -s 5 -c 1 -f -o gpurun_out/r2_spmv_final python tools/spmvbench.py 1e6 16 > /dev/null 2>&1
ls -la gpurun_out/*final*.ncu-rep 2>/dev/null | tail -3
