#!/bin/bash
python scripts/bench_dwln.py > gpurun_out/dwln_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dwln -s 2 -c 1 -o gpurun_out/ncu_dwln_s4 -f python scripts/bench_dwln.py "s4 enc" > gpurun_out/ncu_dwln.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dwln -s 2 -c 1 -o gpurun_out/ncu_dwln_s16 -f python scripts/bench_dwln.py "s16 k5" >> gpurun_out/ncu_dwln.log 2>&1
cat gpurun_out/dwln_bench.log
