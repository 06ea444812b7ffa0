"""profiles/r2_ncu_traffic.json from `ncu --set full` reports: per captured kernel launch the MEASURED DRAM traffic
(dram__bytes_read.sum + dram__bytes_write.sum), duration and tensor / FMA pipe activity, next to the algorithmic bytes.
bench.py reads this file for the `traffic` fields of its roofline records (never typed in by hand).

    python scripts/ncu_traffic.py 'gemm|fc2 H/8 M=49152 K=768 N=384|603979776|gpurun_out/x.ncu-rep'  'dwln|...'  > profiles/r2_ncu_traffic.json
    (class | label | algorithmic bytes | report)"""
import csv
import io
import json
import subprocess
import sys

UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}
out = {}
for arg in sys.argv[1:]:
    cls, label, alg, path = arg.split('|', 3)
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    get = lambda k: float(vals[hdr.index(k)].replace(',', '')) * UNIT.get(units[hdr.index(k)], 1)
    rec = dict(what=label, kernel=vals[hdr.index('Kernel Name')][:80], dram_bytes=get('dram__bytes_read.sum') + get('dram__bytes_write.sum'),
               dram_read=get('dram__bytes_read.sum'), dram_write=get('dram__bytes_write.sum'), algorithmic_bytes=float(alg),
               us=get('gpu__time_duration.sum'), tensor_active_pct=get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
               fma_active_pct=get('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
               issue_active_pct=get('smsp__issue_active.avg.pct_of_peak_sustained_active'), report=path.split('/')[-1])
    if cls == 'gemm':
        out.setdefault('gemm', {})[label] = rec
    else:
        out[cls] = rec
print(json.dumps(out, indent=1))
