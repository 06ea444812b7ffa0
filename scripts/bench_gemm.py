"""Per-shape throughput of lvae_gemm on the main qarv_base contractions at batch 8 x 512x768 (CUDA events, L2 flushed
between repetitions by cycling through enough distinct operand buffers)."""
import sys, ctypes as C
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'lossy-vae_b200'))
from lvae import _native as N
import os
if os.environ.get('LVAE_LIB_PATH'):      # tuning builds (scripts only)
    N._LIB_PATH = Path(os.environ['LVAE_LIB_PATH'])
lib = N.lib()
NPL = {1: 2, 2: 1, 3: 3, 4: 2}
TERMS = {0: 1, 1: 3, 2: 1, 3: 6, 4: 3}
precs = [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else [1, 2]
only = sys.argv[2].split(',') if len(sys.argv) > 2 else None
SHAPES = [  # (name, M, K, N, epi)
    ('s4 enc fc1', 196608, 192, 384, 1), ('s4 enc fc2', 196608, 384, 192, 2),
    ('s8 enc fc1', 49152, 384, 768, 1), ('s8 enc fc2', 49152, 768, 384, 2),
    ('s16 enc fc1', 12288, 512, 1024, 1), ('s16 enc fc2', 12288, 1024, 512, 2),
    ('s32 fc1', 3072, 512, 1024, 1), ('s64 fc1', 768, 512, 2048, 1),
    ('s4 dec fc1', 196608, 128, 192, 1), ('s4 dec fc2', 196608, 192, 128, 2),
    ('s8 dec fc1', 49152, 256, 448, 1), ('s8 dec fc2', 49152, 448, 256, 2),
    ('s8 dec2 fc1', 49152, 256, 512, 1), ('s8 dec2 fc2', 49152, 512, 256, 2),
    ('s16 dec fc1', 12288, 384, 768, 1), ('s16 dec fc2', 12288, 768, 384, 2),
    ('s32 fc2', 3072, 1024, 512, 2), ('s32 wide fc1', 3072, 512, 1536, 1), ('s32 wide fc2', 3072, 1536, 512, 2),
    ('s64 fc2', 768, 1024, 512, 2), ('s64 wide fc2', 768, 2048, 512, 2),
    ('n32 probe', 196608, 192, 32, 0), ('n64 probe', 196608, 192, 64, 0), ('n128 probe', 196608, 192, 128, 0),
]
def planes(x, n, prec=3, weight=False):
    ps = [torch.empty(x.shape, dtype=torch.bfloat16, device='cuda') for _ in range(n)]
    args = [p.data_ptr() for p in ps] + [0] * (3 - n)
    N.check(lib.lvae_split_planes(x.data_ptr(), args[0], args[1], args[2], x.numel(), 1 if prec == 4 else 0,
                                  256.0 if (prec == 4 and weight) else 1.0, 0))
    return ps
for prec in precs:
    for name, M, K, Nn, epi in SHAPES:
        if only and not any(o in name for o in only):
            continue
        nbuf = max(2, int(300e6 // (M * (K + Nn) * 4)) + 1)
        g = torch.Generator().manual_seed(0)
        w = (torch.randn(Nn, K, generator=g) / K ** 0.5).cuda(); b = torch.randn(Nn, generator=g).cuda()
        gamma = torch.rand(Nn, generator=g).cuda()
        wp = planes(w, NPL.get(prec, 1), prec, True) if prec else []
        bufs = []
        for i in range(nbuf):
            x = torch.randn(M, K, device='cuda')
            ap = planes(x, NPL[prec], prec) if prec else []
            out = torch.randn(M, Nn, device='cuda')
            op = [torch.empty(M, Nn, dtype=torch.bfloat16, device='cuda') for _ in range(NPL[prec])] if (prec and epi == 1) else []
            d = N.GemmDesc()
            d.a0 = x.data_ptr(); d.B, d.H, d.W, d.C0 = 1, 1, M, K; d.ksize, d.stride, d.pad = 1, 1, 0
            d.w = w.data_ptr(); d.bias = b.data_ptr(); d.N = Nn; d.epilogue = epi; d.precision = prec
            d.gamma = gamma.data_ptr(); d.res = out.data_ptr(); d.out = 0 if op else out.data_ptr()
            N.set_planes(d, 'w', wp); N.set_planes(d, 'a', ap); N.set_planes(d, 'out', op)
            if not prec:
                del x
            bufs.append((d, x if not prec else None, ap, out, op))
        for d, *_ in bufs[:2]:
            N.check(lib.lvae_gemm(C.byref(d), 0))
        torch.cuda.synchronize()
        reps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for d, *_ in bufs:
                N.check(lib.lvae_gemm(C.byref(d), 0))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * nbuf)
        fl = 2.0 * M * K * Nn
        byts = M * K * 2 * NPL.get(prec, 2) + M * Nn * (2 * NPL.get(prec, 2) if epi == 1 else 8) if prec else M * (K + Nn) * 4
        pair = lib.lvae_gemm2_launch_count()
        print(f'prec={prec} pair={int(pair > globals().get("_pair0", 0))} {name:12s} M={M:6d} K={K:4d} N={Nn:4d}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s algorithmic '
              f'({fl * TERMS[prec] / ms / 1e9:7.1f} issued)  {byts / ms / 1e6:7.0f} GB/s', flush=True)
        import numpy as _np
        pr = _np.zeros(16, dtype=_np.uint64)
        lib.lvae_debug_prof(2 if pair > globals().get("_pair0", 0) else 1, pr.ctypes.data)
        nt = max(1, int(pr[8]))
        print(f'    per tile (clk): mma thread {int(pr[0]) // nt} = wait operands {int(pr[1]) // nt} + wait accumulator {int(pr[2]) // nt} + issue '
              f'{(int(pr[0]) - int(pr[1]) - int(pr[2])) // nt}; producer waits for a free stage {int(pr[4]) // nt} of {int(pr[3]) // nt}; '
              f'epilogue warp: wait acc {int(pr[6]) // nt}, drain {int(pr[7]) // nt}, total {int(pr[5]) // nt}; tiles {nt}', flush=True)
        _pair0 = pair
        del bufs
        torch.cuda.empty_cache()
