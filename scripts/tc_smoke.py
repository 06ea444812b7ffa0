"""Minimal tensor-core GEMM check (run under `timeout`: a pipeline bug shows up as a hang)."""
import sys, ctypes as C
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'lossy-vae_b200'))
from lvae import _native as N
lib = N.lib()
def run(M, K, Nn, prec):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, K, generator=g).cuda(); w = (torch.randn(Nn, K, generator=g) / K ** 0.5).cuda()
    pl = [torch.empty_like(w, dtype=torch.bfloat16) for _ in range(3)]
    N.check(lib.lvae_split_bf16(w.data_ptr(), pl[0].data_ptr(), pl[1].data_ptr(), pl[2].data_ptr(), w.numel(), 0))
    ws = torch.empty(M * K * 3, dtype=torch.bfloat16, device='cuda')
    out = torch.full((M, Nn), float('nan'), device='cuda')
    d = N.GemmDesc()
    d.a0 = x.data_ptr(); d.B, d.H, d.W, d.C0 = 1, 1, M, K; d.ksize, d.stride, d.pad = 1, 1, 0
    d.w = w.data_ptr(); d.N = Nn; d.out = out.data_ptr(); d.precision = prec
    N.set_planes(d, 'w', pl); d.workspace = ws.data_ptr(); d.workspace_bytes = ws.numel() * 2
    N.check(lib.lvae_gemm(C.byref(d), 0)); torch.cuda.synchronize()
    ref = x.double() @ w.double().t()
    err = (out.double() - ref).abs().max().item()
    print(f'M={M} K={K} N={Nn} prec={prec} max err {err:.3e} finite={torch.isfinite(out).all().item()}', flush=True)
for args in [(128, 64, 16, 2), (128, 64, 16, 1), (128, 64, 16, 3), (1000, 192, 384, 3), (5000, 448, 256, 3), (20000, 2048, 512, 3), (3000, 1024, 2048, 3), (256, 128, 128, 1), (1000, 192, 384, 1), (5000, 448, 256, 1), (20000, 2048, 512, 1), (777, 48, 192, 1)]:
    run(*args)
