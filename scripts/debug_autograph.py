"""Debug helper: at which call level does the capture after an eager step break."""
import sys, traceback
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / 'lossy-vae_b200', ROOT / 'oracle'):
    sys.path.insert(0, str(p))
import lvae
from oracle_inputs import make_input
variant = sys.argv[1] if len(sys.argv) > 1 else 'capture'
m = lvae.get_model('qarv_base').cuda().train()
im = make_input('synth', 2, 64, 64, 1).cuda()
lmb = torch.tensor([64.0, 1024.0], device='cuda')
T = m.train_path
out = m(im, lmb=lmb); out['loss'].backward(); torch.cuda.synchronize()
T.autograph_enabled = True
try:
    if variant == 'capture':
        T.autograph._capture(im, lmb)
    elif variant == 'call':
        r = T.autograph(im, lmb)
    elif variant == 'ftrain':
        r = m._forward_train(im, lmb)
    else:
        r = m(im, lmb=lmb)
    print(variant, 'OK')
except Exception as e:
    print(variant, 'FAILED', type(e).__name__, str(e)[:60])
