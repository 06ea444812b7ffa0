"""Top stall sites of an `ncu --page source --csv` dump: python scripts/ncu_top_stalls.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[idx['# Samples']]) for r in body)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[idx[h]] or 0) for r in body) for h in stall_cols}
print('total samples', tot, {k: f'{v / tot * 100:.0f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for i, r in sorted(enumerate(body), key=lambda ir: -int(ir[1][idx['# Samples']]))[:n]:
    s = int(r[idx['# Samples']])
    top = sorted(((int(r[idx[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
    print(f'{i:5d} {s / tot * 100:5.1f}%  {r[idx["Source"]].strip()[:90]:90s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}')
