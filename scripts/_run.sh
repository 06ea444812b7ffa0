cd $GRAFT_REPO_ROOT
timeout -s KILL 1500 python -m pytest tests/test_gpu_train.py -q -m gpu --timeout 600 --timeout-method thread 2>&1 | grep -E "^FAILED|^ERROR|passed|failed|Error|assert|worst" | head -30
timeout -s KILL 600 python bench.py --workload train-qres --steps 5 --warmup 3 > gpurun_out/r2_train_qres_b.json 2> gpurun_out/r2_train_qres_b.err; python -c "
import json; d=json.load(open('gpurun_out/r2_train_qres_b.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['peak_mem_gb'])"; tail -3 gpurun_out/r2_train_qres_b.err
