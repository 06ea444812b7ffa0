"""Per-shape throughput of lvae_dwconv_ln_adaln_planes on the qarv_base shapes at batch 8 x 512x768 (CUDA events; the
input buffers are cycled so that every launch reads from HBM, not L2)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'lossy-vae_b200'))
from lvae import _native as N
import os
if os.environ.get('LVAE_LIB_PATH'):      # tuning builds (scripts only)
    N._LIB_PATH = Path(os.environ['LVAE_LIB_PATH'])
lib = N.lib()
only = sys.argv[1].split(',') if len(sys.argv) > 1 else None
npl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
SHAPES = [('s4 enc', 8, 128, 192, 192, 7), ('s4 dec', 8, 128, 192, 128, 7), ('s8 enc', 8, 64, 96, 384, 7),
          ('s8 dec', 8, 64, 96, 256, 7), ('s16 k5', 8, 32, 48, 512, 5), ('s16 k7', 8, 32, 48, 384, 7),
          ('s32 k3', 8, 16, 24, 512, 3), ('s64 k1', 8, 8, 12, 512, 1)]
for name, B, H, W, C, k in SHAPES:
    if only and not any(o in name for o in only):
        continue
    M = B * H * W
    nbuf = max(2, int(400e6 // (M * C * 4)) + 1)
    g = torch.Generator().manual_seed(0)
    dw = (torch.randn(k * k, C, generator=g) / k).cuda(); db = torch.randn(C, generator=g).cuda()
    ada = (torch.randn(B, 2 * C, generator=g) * 0.3).cuda()
    xs = [torch.randn(M, C, device='cuda') for _ in range(nbuf)]
    pl = [[torch.empty(M, C, dtype=torch.float16, device='cuda') for _ in range(npl)] for _ in range(nbuf)]
    def run(i):
        p = [t.data_ptr() for t in pl[i]] + [0] * (3 - npl)
        N.check(lib.lvae_dwconv_ln_adaln_planes(xs[i].data_ptr(), dw.data_ptr(), db.data_ptr(), ada.data_ptr(), 2 * C, 0, 0, 0,
                                                p[0], p[1], p[2], 1, B, H, W, C, k, 0))
    run(0); run(1)
    torch.cuda.synchronize()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for i in range(nbuf):
            run(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * nbuf)
    byts = M * C * (4 + 2 * npl)
    print(f'{name:8s} B={B} {H}x{W} C={C} k={k}: {ms * 1e3:7.1f} us  {byts / ms / 1e6:6.0f} GB/s  ({2 * M * C * k * k / ms / 1e9:5.1f} TFLOP/s fp32)', flush=True)
