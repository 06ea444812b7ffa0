"""Throughput of the fused latent kernel (lvae_latent_eval) versus size: the config-2 per-layer shapes (batch 8) and the
batch-64 L3 shape of SURVEY 8(d) (9.4 M elements, 151 MB algorithmic) -- CUDA events, buffers cycled beyond L2."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'lossy-vae_b200'))
from lvae import _native as N
lib = N.lib()
tab = torch.exp(torch.linspace(torch.log(torch.tensor(0.11)), torch.log(torch.tensor(20.0)), 64)).cuda()
SHAPES = [('L0  b8', 8, 8 * 12, 32), ('L1  b8', 8, 16 * 24, 32), ('L3  b8', 8, 32 * 48, 96), ('L6  b8', 8, 64 * 96, 8),
          ('L3 b64', 64, 32 * 48, 96), ('L3 b256', 256, 32 * 48, 96)]
only = sys.argv[1].split(',') if len(sys.argv) > 1 else None
for name, B, hw, zd in SHAPES:
    if only and not any(o in name for o in only):
        continue
    M = B * hw
    elems = M * zd
    nbuf = max(2, int(400e6 // (elems * 16)) + 1)
    bufs = []
    for i in range(nbuf):
        qm = torch.randn(M, zd, device='cuda') * 2; prior = torch.randn(M, 2 * zd, device='cuda')
        z = torch.empty(M, zd, device='cuda')
        np_ = lib.lvae_latent_num_partials(hw, zd)
        klp = torch.zeros(B, np_, device='cuda')
        bufs.append((qm, prior, z, klp, np_))
    def run(i, compress=False):
        qm, prior, z, klp, np_ = bufs[i]
        N.check(lib.lvae_latent_eval(qm.data_ptr(), prior.data_ptr(), tab.data_ptr(), 64, z.data_ptr(), klp.data_ptr(), np_,
                                     0, 0, 0, B, hw, zd, 0, 0))
    run(0); run(1); torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for i in range(nbuf):
            run(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * nbuf)
    print(f'{name:8s} elems={elems:9d} ({elems * 16 / 1e6:7.1f} MB algorithmic): {ms * 1e3:8.1f} us  {elems * 16 / ms / 1e6:7.0f} GB/s', flush=True)
