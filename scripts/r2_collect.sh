#!/bin/bash
# After `gpurun -- bash scripts/ncu_round2.sh`: turn gpurun_out/ into the committed evidence under profiles/ (run in the build
# container; the .ncu-rep files are read with the local ncu).
cd "$(dirname "$0")/.."
O=gpurun_out; P=profiles
python scripts/ncu_traffic.py \
  "gemm|fc1 H/8 M=49152 K=384 N=768 (planes in, GELU, planes out)|227700000|$O/r2_ncu_fc1_s8.ncu-rep" \
  "gemm|fc2 H/8 M=49152 K=768 N=384 (+layer scale, residual)|303200000|$O/r2_ncu_fc2_s8.ncu-rep" \
  "gemm|fused MLP H/4 M=196608 C=192 hidden=384|453000000|$O/r2_ncu_mlp_s4.ncu-rep" \
  "dwln|H/4 C=192 k=7 launch (8 x 128 x 192 positions), 302 MB algorithmic|302000000|$O/r2_ncu_dwln_s4.ncu-rep" \
  "latent|latent_kernel<eval> B=64 L3 shape (9.4 M elements, 16 B each)|150994944|$O/r2_ncu_latent_L3b64.ncu-rep" \
  > $P/r2_ncu_traffic.json
python scripts/ncu_extract.py \
  "gemm_tc_kernel<2,EK_GELU> fc1 H/8: M=49152 K=384 N=768 (f16x3)::$O/r2_ncu_fc1_s8.ncu-rep" \
  "gemm_tc_kernel<2,EK_ROWS> fc2 H/8: M=49152 K=768 N=384 (+layer scale, residual)::$O/r2_ncu_fc2_s8.ncu-rep" \
  "mlp_tc_kernel<3> fused MLP H/4: M=196608 C=192 hidden=384::$O/r2_ncu_mlp_s4.ncu-rep" \
  "dwln_kernel<1,7,3> H/4 C=192 k=7::$O/r2_ncu_dwln_s4.ncu-rep" \
  "latent_kernel<eval, vec4> B=64 L3 shape::$O/r2_ncu_latent_L3b64.ncu-rep" \
  > $P/r2_ncu_sections.md   # appended by hand below the analysis header of r2_ncu_summary.md
python scripts/summarize_launches.py $O/r2_launches_f16x3_tail1.csv > $P/r2_launches_f16x3_tail1.md
cp $O/r2_launches_f16x3_tail1.csv $P/
python scripts/parity_report.py $O/r2_parity_records.jsonl $P/r2_parity.md
for f in r2_final_bench.json r2_final_bench_reference.json r2_final_bench_f16x3.json r2_final_bench_rd.json r2_final_bench_qres.json \
         r2_final_bench_codec.json r2_final_bench_codec_batch8.json r2_train_qarv.json r2_train_qres.json \
         r2_bench_latent.log r2_bench_dwln.log r2_bench_gemm.log r2_gputest.log r2_profile_plan.log; do
  [ -s $O/$f ] && cp $O/$f $P/$f
done
ls -la $P | grep r2_ | wc -l
