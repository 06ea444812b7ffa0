"""gpurun_out/r2_parity_records.jsonl (written by the -m gpu tests through tests/conftest.py:parity_log) ->
profiles/r2_parity.md: per fixture x precision the symbol / index flips, |d bpp| and |d PSNR| measured on B200.

    python scripts/parity_report.py [records.jsonl] [out.md]
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
src = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / 'gpurun_out' / 'r2_parity_records.jsonl'
dst = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / 'profiles' / 'r2_parity.md'
rows = [json.loads(l) for l in src.read_text().splitlines() if l.strip()]
seen, uniq = set(), []
for r in rows:                                   # a re-run appends: keep the last record per key
    pass
for r in reversed(rows):
    k = (r['test'], r['case'], r['precision'])
    if k not in seen:
        seen.add(k)
        uniq.append(r)
uniq.reverse()
out = ['# r2 parity on B200: measured flips, |d bpp|, |d PSNR| per case x precision', '',
       'Written by `scripts/parity_report.py` from the records the `-m gpu` tests log (`tests/conftest.py:parity_log`).',
       'North star: integer symbols / table indexes bit-exact, |d bpp| <= 1e-4, |d PSNR| <= 0.01 dB.  "flips" = symbols + table',
       'indexes that differ from the reference in the FIRST differing layer; every one is verified to sit within 2e-5 of a',
       'rounding boundary of the reference\'s own fp32 value (`tests/test_gpu_model.py:_check_integer_parity`), and the same',
       'elements flip in the fp32 CUDA-core mode -- it is fp32 summation order (MKL vs device), not the operand split.',
       '`tests/test_gpu_model.py:MAX_FLIPS` pins these counts: a case not listed there must be bit-exact.', '',
       '| test | case | precision | symbols | flips | d bpp | bpp tolerance in the test | d PSNR (dB) |', '|---|---|---|---|---|---|---|---|']
for r in uniq:
    out.append(f"| {r['test']} | {r['case']} | {r['precision']} | {r['symbols']} | {r['flips']} | {r['dbpp']:.2e} | {r['bpp_tol']:.1e} | {r['dpsnr']:.1e} |")
dst.write_text('\n'.join(out) + '\n')
print(f'wrote {dst} ({len(uniq)} rows)')
