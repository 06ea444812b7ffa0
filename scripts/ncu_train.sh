#!/bin/bash
# ncu --set full of the training-step kernels at the qres34m H/4 shape (one launch each, after the warm-up launches)
python scripts/bench_train_kernels.py > gpurun_out/train_kernels.log 2>&1
for kn in dwconv_wgrad_kernel ln_mod_bwd_kernel split_planes_t_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -o gpurun_out/ncu_train_$kn -f python scripts/bench_train_kernels.py > gpurun_out/ncu_train_$kn.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:dwconv_kernel -s 2 -c 1 -o gpurun_out/ncu_train_dwconv_kernel -f python scripts/bench_train_kernels.py > gpurun_out/ncu_train_dwconv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/ncu_train_wgrad_gemm -f python scripts/bench_train_kernels.py > gpurun_out/ncu_train_wgrad_gemm.log 2>&1
cat gpurun_out/train_kernels.log
