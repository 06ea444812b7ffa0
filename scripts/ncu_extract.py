"""Markdown summary of ncu --set full reports: python scripts/ncu_extract.py title::path.ncu-rep ... > profiles/x.md"""
import csv, subprocess, sys, io, collections
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__cluster_size', 'smsp__inst_executed.sum']
for arg in sys.argv[1:]:
    title, path = arg.rsplit('::', 1)
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
    print(f'## {title}\n\n`{name[:110]}`\n')
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            print(f'- `{h}` = {v} {u}')
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    tot = sum(int(r[idx['# Samples']]) for r in body) or 1
    cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = sorted(((sum(int(r[idx[h]] or 0) for r in body), h) for h in cols), reverse=True)[:7]
    print('- stall samples: ' + ', '.join(f'{h[6:]} {v / tot * 100:.0f}%' for v, h in agg))
    ops = collections.Counter()
    for r in body:
        op = r[idx['Source']].strip().split()
        o = op[1] if op and op[0].startswith('@') and len(op) > 1 else (op[0] if op else '?')
        ops[o.split('.')[0]] += int(r[idx['Instructions Executed']])
    n = sum(ops.values()) or 1
    print('- warp instructions: ' + ', '.join(f'{k} {v / n * 100:.0f}%' for k, v in ops.most_common(8)) + '\n')
