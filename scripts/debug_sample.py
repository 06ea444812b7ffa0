import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / 'lossy-vae_b200', ROOT / 'oracle'):
    sys.path.insert(0, str(p))
import lvae, lvae_oracle as O
from lvae import _native as N
sd = O.sensitised_state_dict(O.qarv_param_shapes(), seed=0)
m = lvae.get_model('qarv_base'); m.load_state_dict(sd, strict=False); m = m.cuda().eval(); m.compress_mode()
eng = m.engine; eng.use_graphs = False
eng.refresh_weights()
B, nH, nW = 2, 1, 2
P = eng._build_decode_plan(B, nH, nW, sampling=True)
P.lmb.copy_(torch.full((B,), 256.0))
st = torch.cuda.current_stream().cuda_stream
for li in range(len(P.z)):
    for o in P.segments[li]:
        rc = o.fn(*o.args, st); assert rc == 0, o.name
        torch.cuda.synchronize()
    hw, zd, _, _, Hs, Ws = P.layout[li]
    print(li, 'prior finite', torch.isfinite(P.prior[li]).all().item(), P.prior[li].abs().max().item(), 'segment ops', [o.name for o in P.segments[li]])
    rn = torch.randn(B * hw, zd, device='cuda'); un = torch.empty(B * hw, zd, device='cuda').uniform_(-0.5, 0.5)
    N.check(eng.lib.lvae_latent_sample(P.prior[li].data_ptr(), rn.data_ptr(), un.data_ptr(), 1.0, P.z[li].data_ptr(), B, hw, zd, st))
    torch.cuda.synchronize()
    print('   z finite', torch.isfinite(P.z[li]).all().item(), P.z[li].abs().max().item())
for o in P.segments[len(P.z)]:
    rc = o.fn(*o.args, st); assert rc == 0
torch.cuda.synchronize()
print('x_hat finite', torch.isfinite(P.x_hat).all().item())
