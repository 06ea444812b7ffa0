"""Per-op device times of the batch-8 512x768 eval plan (CUDA events around eager launches)."""
import sys
from pathlib import Path
from collections import defaultdict
import torch
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / 'lossy-vae_b200', ROOT / 'oracle'):
    sys.path.insert(0, str(p))
import lvae, lvae_oracle as O
from oracle_inputs import make_input
prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x6'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
name = sys.argv[3] if len(sys.argv) > 3 else 'qarv_base'
if name == 'qres34m':
    import qres_oracle as Q
    m = lvae.get_model('qres34m', lmb=2048); m.load_state_dict(O.sensitised_state_dict(Q.qres_param_shapes(), seed=0), strict=False)
else:
    m = lvae.get_model('qarv_base'); m.load_state_dict(O.sensitised_state_dict(O.qarv_param_shapes(), seed=0), strict=False)
m.precision = prec
m = m.cuda().eval()
eng = m.engine
P = eng.forward_plan(B, 512, 768, 'eval')
P.im.copy_(make_input('rand', B, 512, 768, 0)); P.lmb.fill_(2048.0)
for _ in range(3):
    eng.replay(P)
torch.cuda.synchronize()
prof = eng.profile_ops(P, reps=5)
tot = sum(ms for _, _, ms in prof)
print(f'precision {prec}: {len(prof)} ops, {tot:.2f} ms per step (eager, serialised)')
groups = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for name, meta, ms in prof:
    if meta.get('kind') == 'gemm':
        key = f"{name:10s} M={meta['M']:6d} K={meta['K']:4d} N={meta['N']:4d}"
    elif meta.get('kind') == 'dwln':
        key = f"dwln bytes={meta['bytes'] / 1e6:6.1f}MB"
    else:
        key = name
    g = groups[key]; g[0] += 1; g[1] += ms; g[2] += meta.get('flops', 0); g[3] += meta.get('bytes', 0)
for key, (n, ms, fl, by) in sorted(groups.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f'{key:44s} x{n:3d} {ms:7.3f} ms ({ms / tot * 100:4.1f}%)  {ms / n * 1e3:7.1f} us each  {fl / ms / 1e9 if ms else 0:6.1f} TFLOP/s  {by / ms / 1e6 if ms else 0:6.0f} GB/s')
