cd $GRAFT_REPO_ROOT
timeout -s KILL 600 python -m pytest tests/test_gpu_train.py -q -m gpu -k "fused_clip or clipping" --timeout 300 --timeout-method thread 2>&1 | tail -3
time (timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err)
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['n_gpus']); print(d['train'])"; tail -5 gpurun_out/r2_bench_n2.err
