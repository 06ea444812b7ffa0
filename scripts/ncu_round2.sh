#!/bin/bash
# Round-2 evidence (run under gpurun from the repo root): GPU tests with parity records, per-shape microbenches, the ncu
# launch list of the bench command, `ncu --set full` captures of the top kernels, and the bench lines of every workload.
cd ${GRAFT_REPO_ROOT:-.}
O=gpurun_out
rm -f $O/r2_parity_records.jsonl
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 300 --timeout-method thread 2>&1 | tail -4 > $O/r2_gputest.log; cat $O/r2_gputest.log
python scripts/bench_latent.py > $O/r2_bench_latent.log 2>&1; cat $O/r2_bench_latent.log
python scripts/bench_dwln.py > $O/r2_bench_dwln.log 2>&1
python scripts/bench_gemm.py 4 > $O/r2_bench_gemm.log 2>&1
python scripts/profile_plan.py f16x3+tail1 8 > $O/r2_profile_plan.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file $O/r2_launches_f16x3_tail1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-record > $O/r2_launches_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/r2_ncu_fc1_s8 -f python scripts/bench_gemm.py 4 "s8 enc fc1" > /dev/null 2>&1
$NCU -k regex:gemm_tc_kernel -s 3 -c 1 -o $O/r2_ncu_fc2_s8 -f python scripts/bench_gemm.py 4 "s8 enc fc2" > /dev/null 2>&1
$NCU -k regex:mlp_tc_kernel -s 8 -c 1 -o $O/r2_ncu_mlp_s4 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-train-record > /dev/null 2>&1
$NCU -k regex:dwln -s 2 -c 1 -o $O/r2_ncu_dwln_s4 -f python scripts/bench_dwln.py "s4 enc" > /dev/null 2>&1
$NCU -k regex:latent_kernel -s 3 -c 1 -o $O/r2_ncu_latent_L3b64 -f python scripts/bench_latent.py "L3 b64" > /dev/null 2>&1
python bench.py --steps 20 --warmup 5 > $O/r2_final_bench.json 2> $O/r2_final_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_final_bench_reference.json 2>/dev/null
python bench.py --precision f16x3 --no-cpu-baseline --no-train-record > $O/r2_final_bench_f16x3.json 2>/dev/null
python bench.py --workload rd --no-cpu-baseline > $O/r2_final_bench_rd.json 2>/dev/null
python bench.py --workload qres --no-cpu-baseline > $O/r2_final_bench_qres.json 2>/dev/null
python bench.py --workload codec --steps 10 > $O/r2_final_bench_codec.json 2>/dev/null
python bench.py --workload codec --codec-batch 8 --steps 10 > $O/r2_final_bench_codec_batch8.json 2>/dev/null
python bench.py --workload train --steps 10 > $O/r2_train_qarv.json 2>/dev/null
python bench.py --workload train-qres --steps 5 > $O/r2_train_qres.json 2>/dev/null
tail -c 400 $O/r2_final_bench.json; tail -3 $O/r2_final_bench.err
