cd $GRAFT_REPO_ROOT
timeout -s KILL 1500 python -m pytest tests/test_gpu_qres.py tests/test_gpu_train.py -q -m gpu -k "lossless or crop" -s --timeout 600 --timeout-method thread 2>&1 | grep -E "^FAILED|^ERROR|passed|failed|Error|assert|^E |lossless residual" | head -40
