"""Throughput of the training-step kernels (csrc/dwln_bwd.cu, csrc/wgrad.cu) at the qres34m H/4 shape of BASELINE
configs[2] (batch 16 x 512x768 -> [16,128,192,192], hidden 384), CUDA events, buffers cycled so every launch reads HBM.
usage: bench_train_kernels.py [B]"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'lossy-vae_b200'))
from lvae import _native as N
lib = N.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, W, C, hid, k = 128, 192, 192, 384, 7
M = B * H * W
nbuf = 3
g = torch.Generator().manual_seed(0)
dw = (torch.randn(k * k, C, generator=g) / k).cuda(); db = torch.randn(C, generator=g).cuda()
ada = (torch.randn(B, 2 * C, generator=g) * 0.3).cuda()
xs = [torch.randn(M, C, device='cuda') for _ in range(nbuf)]
ys = [torch.randn(M, C, device='cuda') for _ in range(nbuf)]
out = [torch.empty(M, C, device='cuda') for _ in range(nbuf)]
hs = [torch.randn(M, hid, device='cuda') for _ in range(2)]
dmod, dwp, dbp = torch.empty(B, 2 * C, device='cuda'), torch.empty(k * k, C, device='cuda'), torch.empty(C, device='cuda')
ta = [torch.empty(C * M, dtype=torch.bfloat16, device='cuda') for _ in range(2)]
tb = [torch.empty(hid * M, dtype=torch.bfloat16, device='cuda') for _ in range(2)]
dW = torch.empty(hid, C, device='cuda'); cs = torch.zeros(hid, device='cuda')
p = lambda t: t.data_ptr()


def timed(name, fn, bytes_, flops=0, reps=3):
    fn(0); fn(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        for i in range(nbuf):
            fn(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * nbuf)
    print(f'{name:34s} {us:8.1f} us  {bytes_ / us / 1e3:7.1f} GB/s algorithmic' + (f'  {flops / us / 1e6:7.1f} TFLOP/s' if flops else ''))


E4 = M * C * 4
timed('dwconv (recompute c)', lambda i: N.check(lib.lvae_dwconv(p(xs[i % nbuf]), p(dw), p(db), 0, p(out[i % nbuf]), B, H, W, C, k, 0, 0)), 2 * E4, 2 * M * C * k * k)
timed('ln_mod_bwd', lambda i: N.check(lib.lvae_ln_mod_bwd(p(xs[i % nbuf]), p(ys[i % nbuf]), p(ada), 2 * C, 0, 0, p(out[i % nbuf]), p(dmod), B, H * W, C, 0)), 3 * E4)
timed('dwconv_wgrad', lambda i: N.check(lib.lvae_dwconv_wgrad(p(ys[i % nbuf]), p(xs[i % nbuf]), p(dwp), p(dbp), B, H, W, C, k, 0)), 2 * E4, 2 * M * C * k * k)
timed('dwconv flipped + add (dgrad)', lambda i: N.check(lib.lvae_dwconv(p(ys[i % nbuf]), p(dw), 0, p(xs[i % nbuf]), p(out[i % nbuf]), B, H, W, C, k, 1, 0)), 3 * E4, 2 * M * C * k * k)
timed('split_planes_t [M,192]', lambda i: N.check(lib.lvae_split_planes_t_ex(p(xs[i % nbuf]), p(ta[0]), p(ta[1]), M, C, 0, 0, 0)), 2 * E4)
timed('split_planes_t gelu+colsum [M,384]', lambda i: N.check(lib.lvae_split_planes_t_ex(p(hs[i % 2]), p(tb[0]), p(tb[1]), M, hid, 1, p(cs), 0)), 2 * M * hid * 4)
timed('gemm_wgrad 384x192 over M', lambda i: N.check(lib.lvae_gemm_wgrad(p(tb[0]), p(tb[1]), p(ta[0]), p(ta[1]), p(dW), hid, C, M, 0)), 4 * M * (C + hid), 2 * M * C * hid)
