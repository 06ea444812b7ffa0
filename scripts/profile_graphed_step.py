"""Kernel-time table of one replay of GraphedTrainStep (qarv 16 x 256^2): where the captured step's GPU time goes."""
import sys
from pathlib import Path
from collections import defaultdict
import torch
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / 'lossy-vae_b200', ROOT / 'oracle'):
    sys.path.insert(0, str(p))
import lvae
from lvae.training import GraphedTrainStep
from oracle_inputs import make_input
torch.manual_seed(0)
m = lvae.get_model('qarv_base').cuda().train()
opt = torch.optim.Adam(m.parameters(), lr=1e-4)
im = make_input('rand', 16, 256, 256, 1).cuda()
step = GraphedTrainStep(m, opt, tuple(im.shape))
for _ in range(4):
    step(im)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(im)
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name[:70]
        agg[k][0] += 1; agg[k][1] += e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total
tot = sum(v[1] for v in agg.values())
print(f'{sum(v[0] for v in agg.values())} kernels / memcpys, {tot / 1e3:.2f} ms of kernel time in one replay')
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:32]:
    print(f'{t / 1e3:8.3f} ms {t / tot * 100:5.1f} %  x{n:4d}  {k}')
