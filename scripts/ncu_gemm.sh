#!/bin/bash
NCU="ncu --set full --clock-control none --import-source on"
for sh in "s4 enc fc1" "s8 enc fc1" "s4 enc fc2" "s8 enc fc2"; do
  tag=$(echo $sh | tr ' ' '_')
  $NCU -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/ncu2_$tag -f python scripts/bench_gemm.py 4 "$sh" > gpurun_out/ncu2_$tag.log 2>&1
done
python scripts/bench_gemm.py 4 > gpurun_out/bench_gemm_f16x3.log 2>&1
