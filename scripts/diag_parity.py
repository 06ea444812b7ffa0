"""GPU-side parity diagnostics against the golden fixtures (prints per-layer mismatch counts)."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / 'lossy-vae_b200', ROOT / 'oracle'):
    sys.path.insert(0, str(p))
import lvae
import lvae_oracle as O
from oracle_inputs import CASES, make_input

prec = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
sd = O.sensitised_state_dict(O.qarv_param_shapes(), seed=0)
torch.manual_seed(0)
m = lvae.get_model('qarv_base')
m.load_state_dict(sd, strict=False)
m.precision = prec
m = m.to('cuda:0').eval()
m.compress_mode()
for name, (kind, nB, H, W, lmbs, seed) in CASES.items():
    g = np.load(ROOT / 'tests' / 'golden' / f'{name}.npz')
    im = make_input(kind, nB, H, W, seed).cuda()
    lmb = torch.tensor(lmbs).cuda()
    st = m(im, lmb=lmb, return_rec=True)
    print(f'== {name} [{prec}] bppix {st["bppix"]:.6f} vs {float(g["bppix"]):.6f} (d={st["bppix"]-float(g["bppix"]):+.2e}) '
          f'psnr {st["psnr"]:.4f} vs {float(g["psnr"]):.4f}  im_hat maxerr {(st["im_hat"].cpu()-torch.from_numpy(g["im_hat"])).abs().max():.2e}')
    m.forward_end2end(im, lmb, mode='compress')
    P = m.engine._plans[(nB, H, W, 'compress', False)]
    x_hat, lat = m.forward_end2end(im, lmb, get_latent=True)
    for li in range(9):
        sym, idx = P.sym[li].cpu().numpy(), P.idx[li].cpu().numpy()
        ns = int((sym != g[f'sym{li}']).sum()); ni = int((idx != g[f'idx{li}']).sum())
        kl = lat[li]['kl'].sum(dim=(1, 2, 3)).cpu().numpy()
        zerr = (lat[li]['z'].cpu() - torch.from_numpy(g[f'z{li}'])).abs().max().item()
        print(f'  L{li}: n={sym.size:7d} sym_mismatch={ns} idx_mismatch={ni} zmaxerr={zerr:.2e} kl {kl} ref {g["kl_per_image"][li]} rel {(kl-g["kl_per_image"][li])/g["kl_per_image"][li]}')
