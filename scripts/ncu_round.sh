#!/bin/bash
# ncu captures of the hot kernels (run under gpurun): full sets for the H/4 MLP GEMMs and dwln, launch list of bench.py
set -x
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/ncu_fc1_s4 -f python scripts/bench_gemm.py 4 "s4 enc fc1" > gpurun_out/ncu_fc1_s4.log 2>&1
$NCU -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/ncu_fc2_s4 -f python scripts/bench_gemm.py 4 "s4 enc fc2" > gpurun_out/ncu_fc2_s4.log 2>&1
$NCU -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/ncu_fc2_s8 -f python scripts/bench_gemm.py 4 "s8 enc fc2" > gpurun_out/ncu_fc2_s8.log 2>&1
$NCU -k regex:dwln -s 40 -c 1 -o gpurun_out/ncu_dwln_s4 -f python scripts/profile_plan.py f16x3 > gpurun_out/ncu_dwln.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_f16x3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
