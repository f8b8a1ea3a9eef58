#!/bin/bash
# ncu evidence: launch list of the bench's timed region, full capture of the eval kernel, full capture of the non-product kernels
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-configs > gpurun_out/r2r_bench_under_ncu.log 2>&1
echo "launch list rc=$?"; tail -3 gpurun_out/r2r_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_blocked -s 40 -c 3 -o gpurun_out/r2r_eval python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-configs > gpurun_out/r2r_eval_ncu.log 2>&1
echo "eval full rc=$?"
timeout 900 ncu --set full --metrics sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --import-source on -k regex:'k_eval_geomean|k_eval_pair' -s 40 -c 14 -o gpurun_out/r2r_mixed python scripts/profile_mixed.py > gpurun_out/r2r_mixed_ncu.log 2>&1
echo "mixed full rc=$?"; tail -3 gpurun_out/r2r_mixed_ncu.log
ls -la gpurun_out/*.ncu-rep
