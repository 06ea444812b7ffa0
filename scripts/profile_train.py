"""Kernel-time table of one training step (torch.profiler, CUDA activities): which kernels the step spends its time in.
usage: profile_train.py [qarv|qres] [batch] [H] [W]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / 'lossy-vae_b200', ROOT / 'oracle'):
    sys.path.insert(0, str(p))
import lvae
from oracle_inputs import make_input
name = sys.argv[1] if len(sys.argv) > 1 else 'qarv'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
H = int(sys.argv[3]) if len(sys.argv) > 3 else 256
W = int(sys.argv[4]) if len(sys.argv) > 4 else 256
torch.manual_seed(0)
m = (lvae.get_model('qres34m', lmb=2048) if name == 'qres' else lvae.get_model('qarv_base')).cuda().train()
opt = torch.optim.Adam(m.parameters(), lr=1e-4)
im = make_input('rand', B, H, W, 1).cuda()


def step():
    out = m(im)
    opt.zero_grad(set_to_none=True)
    out['loss'].backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=90))
