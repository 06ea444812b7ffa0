cd $GRAFT_REPO_ROOT
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm2" --timeout 120 --timeout-method thread 2>&1 | grep -E "AssertionError|passed|failed|FAILED" > gpurun_out/r2_g2_test.log
cat gpurun_out/r2_g2_test.log
LVAE_G2_BN=128 timeout -s KILL 300 python scripts/bench_gemm.py 4 "s8 enc,s16 enc,s8 dec2" > gpurun_out/r2_g2_bench_bn128.log 2>&1; cat gpurun_out/r2_g2_bench_bn128.log
