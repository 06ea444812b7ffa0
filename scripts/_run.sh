cd $GRAFT_REPO_ROOT
rm -f gpurun_out/r2_parity_records.jsonl
timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 300 --timeout-method thread 2>&1 | grep -E "^FAILED|^ERROR|passed|failed|AssertionError" | head -40 > gpurun_out/r2_gputest.log; cat gpurun_out/r2_gputest.log
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 --no-train-record > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_b.json')); print(d['value'], d['e2e']['value'], d['parity'])"; tail -3 gpurun_out/r2_bench_b.err
