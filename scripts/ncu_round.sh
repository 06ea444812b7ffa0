#!/bin/bash
# Round-1 evidence (run under gpurun): per-op profile, ncu launch list of the bench command, full captures of the top kernels
python scripts/profile_plan.py f16x3 > gpurun_out/r1_profile_plan_f16x3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r1_launches_f16x3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_launches_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:dwln -s 2 -c 1 -o gpurun_out/r1_dwln_s4 -f python scripts/bench_dwln.py "s4 enc" > /dev/null 2>&1
$NCU -k regex:latent_kernel -s 22 -c 1 -o gpurun_out/r1_latent_L3b64 -f python scripts/bench_latent.py > /dev/null 2>&1
python bench.py > gpurun_out/r1_bench_f16x3.json 2> gpurun_out/r1_bench_f16x3.err
python bench.py --precision bf16x6 --no-cpu-baseline > gpurun_out/r1_bench_bf16x6.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2>/dev/null
python bench.py --workload rd --no-cpu-baseline > gpurun_out/r1_bench_rd.json 2>/dev/null
tail -c 600 gpurun_out/r1_bench_f16x3.json
