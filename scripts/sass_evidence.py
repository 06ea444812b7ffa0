"""Per-kernel counts of the SASS mnemonics that show tcgen05 / TMEM / TMA / mbarrier / cluster use (B200_PROFILING.md:
the PTX names never appear in SASS), from `cuobjdump -sass` of the built library.  CPU only.
usage: python scripts/sass_evidence.py > profiles/r2_sass_evidence.md"""
import collections
import re
import subprocess
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
lib = ROOT / 'lossy-vae_b200' / 'lib' / 'liblvae_b200.so'
out = subprocess.run(['cuobjdump', '-sass', str(lib)], capture_output=True, text=True).stdout
WANT = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMAPF', 'SYNCS', 'UCGABAR', 'FFMA2']
KEY = WANT[:-1]       # a kernel is listed when it uses any of the Blackwell-specific ones
counts, cur = {}, None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1).split('.')[0]
        if op.startswith('UCGABAR'):
            op = 'UCGABAR'
        counts[cur]['_total'] += 1
        if op in WANT:
            counts[cur][op] += 1
demangle = subprocess.run(['c++filt'], input='\n'.join(counts), capture_output=True, text=True).stdout.splitlines()
print('# SASS evidence: instruction counts per kernel of liblvae_b200.so (cuobjdump -sass, sm_100a)\n')
print('UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG / UTMAPF = TMA tensor loads / prefetch, '
      'SYNCS = mbarrier ops, UCGABAR = cluster barrier, FFMA2 = packed fp32x2 FMA.\n')
print('| kernel | SASS instr | ' + ' | '.join(WANT) + ' |')
print('|---|---|' + '---|' * len(WANT))
for (name, c), dm in sorted(zip(counts.items(), demangle), key=lambda t: t[1]):
    if not any(c[w] for w in KEY):
        continue
    short = re.sub(r'\(.*', '', dm).replace('void ', '').replace('lvae::', '')
    print(f'| `{short[:70]}` | {c["_total"]} | ' + ' | '.join(str(c[w]) if c[w] else '' for w in WANT) + ' |')
