"""Training step (BASELINE configs 2, 3): loss and parameter gradients of `model(batch)['loss'].backward()` on a B200
against torch autograd over the CPU oracle (the restatement pinned to the unmodified reference) with the same weights,
images, lambdas and uniform noise.  Tolerance: per parameter tensor, |g - g_ref|_2 <= GRAD_RTOL * |g_ref|_2 + GRAD_ATOL *
sqrt(numel) -- fp32 round-off through ~110 blocks in a different summation order on both sides."""
import copy

import numpy as np
import pytest
import torch

import lvae_oracle as O
import qres_oracle as Q
from oracle_inputs import make_input

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GRAD_RTOL, GRAD_ATOL = 2e-3, 1e-7


def _noise(model, B, H, W, seed):
    lay, _ = model.engine._latent_layout(B, H // model.max_stride, W // model.max_stride)
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(B, zd, Hs, Ws, generator=g) - 0.5 for (_, zd, _, _, Hs, Ws) in lay]


def _oracle_grads(fwd, sd, *args, **kw):
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    with torch.enable_grad():
        out = fwd.__wrapped__(sd, *args, **kw)
        out['loss'].backward()
    return out, {k: v.grad for k, v in sd.items()}


def _compare(model, ref_grads, loss, ref_loss):
    assert abs(loss - ref_loss) <= 1e-4 * abs(ref_loss), (loss, ref_loss)
    worst, n = (0.0, None), 0
    for name, p in model.named_parameters():
        gr = ref_grads[name]
        if gr is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, f'no gradient reached {name}'
        g = p.grad.detach().cpu().reshape(gr.shape)
        err, scale = float((g - gr).norm()), float(gr.norm())
        tol = GRAD_RTOL * scale + GRAD_ATOL * gr.numel() ** 0.5
        worst = max(worst, (err / max(scale, 1e-30), name))
        assert err <= tol, (name, err, scale)
        n += 1
    assert n > 100
    return worst


@pytest.fixture()
def train_model(gpu_model):
    m = copy.deepcopy(gpu_model)
    m.compressing = False
    return m.train()


@pytest.mark.parametrize('native_bwd', [True, False])
def test_qarv_train_step_gradients_match_oracle_autograd(train_model, sensitised_sd, native_bwd):
    m = train_model
    m.train_path.native_bwd = native_bwd
    B, H, W = 2, 64, 64
    im = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(5))
    lmb = torch.tensor([64.0, 1024.0])
    noise = _noise(m, B, H, W, 17)
    st = m._forward_train(im.to(DEV), lmb.to(DEV), noise=noise)
    assert st['loss'].requires_grad
    st['loss'].backward()
    ref, grads = _oracle_grads(O.qarv_forward, sensitised_sd, im, lmb, mode='train', noise=noise)
    assert abs(st['bppix'] - ref['bppix']) <= 1e-4 * max(1.0, ref['bppix'])
    worst = _compare(m, grads, st['loss'].item(), ref['loss'].item())
    print('worst relative gradient error', worst)


@pytest.mark.parametrize('native_bwd', [True, False])
def test_qarv_gradients_at_default_init_match_oracle_autograd(native_lib, native_bwd):
    """Real training conditions: the reference's default initialisation has layer scale gamma = 1e-6, so the gradients
    inside every residual branch are ~1e-6 of the trunk's.  Adam normalises them, so they must be RELATIVELY right:
    no absolute tolerance here (this is what rules out fp16 operand planes for the gradient GEMMs)."""
    import lvae
    torch.manual_seed(0)
    m = lvae.get_model('qarv_base')
    sd = {k: v.detach().clone() for k, v in m.named_parameters()}
    m = m.to(DEV).train()
    m.train_path.native_bwd = native_bwd
    B, H, W = 2, 64, 64
    im = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(6))
    lmb = torch.tensor([32.0, 2048.0])
    noise = _noise(m, B, H, W, 3)
    st = m._forward_train(im.to(DEV), lmb.to(DEV), noise=noise)
    st['loss'].backward()
    ref, grads = _oracle_grads(O.qarv_forward, sd, im, lmb, mode='train', noise=noise)
    assert abs(st['loss'].item() - ref['loss'].item()) <= 1e-5 * abs(ref['loss'].item())
    worst = (0.0, None)
    for name, p in m.named_parameters():
        gr = grads[name]
        g = p.grad.detach().cpu().reshape(gr.shape)
        scale = float(gr.norm())
        if scale == 0.0:
            assert float(g.norm()) == 0.0, name
            continue
        rel = float((g - gr).norm()) / scale
        worst = max(worst, (rel, name, scale))
        assert rel <= 5e-3, (name, rel, scale)
    print('worst relative gradient error at default init', worst)


def test_qarv_forward_routes_to_training_path_and_adam_reduces_the_loss(native_lib):
    import lvae
    torch.manual_seed(0)
    m = lvae.get_model('qarv_base').to(DEV).train()
    opt = torch.optim.Adam(m.parameters(), lr=2e-4)
    im = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(1)).to(DEV)
    lmb = torch.tensor([256.0, 256.0], device=DEV)
    losses = []
    for _ in range(6):
        torch.manual_seed(7)                       # same noise draw every step
        st = m(im, lmb=lmb)
        opt.zero_grad(set_to_none=True)
        st['loss'].backward()
        opt.step()
        losses.append(st['loss'].item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    m.eval()
    with torch.no_grad():
        ev = m(im, lmb=lmb)                        # the launch plans pick up the updated weights
    assert np.isfinite(ev['loss'].item())


@pytest.mark.parametrize('native_bwd', [True, False, 'vd'])
def test_qres_train_step_gradients_match_oracle_autograd(native_lib, native_bwd):
    """native_bwd 'vd': additionally the VDBlock heads through TrainPath.vd_backward (opt-in: correct but slower than cuDNN)"""
    import lvae
    sd = O.sensitised_state_dict(Q.qres_param_shapes(), seed=0)
    m = lvae.get_model('qres34m', lmb=2048)
    m.load_state_dict(sd, strict=False)
    m = m.to(DEV).train()
    m.train_path.native_bwd = bool(native_bwd)
    m.train_path.native_vd = native_bwd == 'vd'
    B, H, W = 1, 64, 64
    im = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(9))
    noise = _noise(m, B, H, W, 23)
    st = m(im.to(DEV), noise=noise)
    st['loss'].backward()
    ref, grads = _oracle_grads(Q.qres_forward, sd, im, 2048, mode='train', noise=noise)
    worst = _compare(m, grads, st['loss'].item(), float(ref['loss']))
    print('worst relative gradient error', worst)


def test_graphed_train_step_matches_eager_step(native_lib):
    """GraphedTrainStep (whole step in one CUDA graph) against the eager step: three updates each from the same seeded
    init (lambda / noise draws differ), and the inference plans see the weights the replays updated in place."""
    import lvae
    from lvae.training import GraphedTrainStep
    im = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(4)).to(DEV)
    lmb = torch.full((2,), 256.0, device=DEV)
    finals = []
    for graphed in (False, True):
        torch.manual_seed(0)
        m = lvae.get_model('qarv_base').to(DEV)
        with torch.no_grad():
            before = m.eval()(im, lmb=lmb)['loss'].item()
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-4, capturable=True)
        if graphed:
            step = GraphedTrainStep(m, opt, tuple(im.shape), warmup=1)        # the warm-up step is rolled back: 3 replays = 3 updates
            assert step.native                                                # clip + Adam (+ EMA) on the flat buffers, csrc/optim.cu
            losses = [float(step(im)) for _ in range(3)]
            assert step.launches_per_replay > 300
            assert float(step.step_t) == 3.0
        else:
            losses = []
            for _ in range(3):
                st = m(im)
                opt.zero_grad(set_to_none=True)
                st['loss'].backward()
                opt.step()
                losses.append(st['loss'].item())
        assert all(np.isfinite(losses))
        m.eval()
        with torch.no_grad():
            after = m(im, lmb=lmb)['loss'].item()
        finals.append((before, after))
    (b0, a0), (b1, a1) = finals
    assert b0 == b1 and a0 < b0 and a1 < b1, finals            # both paths trained the weights the eval plans read
    assert abs(a1 - a0) <= 0.25 * abs(b0 - a0), finals


@pytest.mark.parametrize('fam', ['qarv', 'qres'])
def test_gradients_match_the_unmodified_reference_fixture(native_lib, golden, fam):
    """Directly against tests/golden/{qarv,qres}_train_grads.npz (loss.backward() of the unmodified reference, generated by
    oracle/gen_golden.py grads): loss, and per parameter tensor the gradient norm and its inner product with a fixed probe."""
    import lvae
    from gen_golden_cases import GRAD_CASES, grad_probe
    from oracle_inputs import QRES_LMB, make_input
    from test_oracle_pinned import _qres_noise
    nB, H, W, lmbs, seed, nseed = GRAD_CASES[fam]
    g = golden(f'{fam}_train_grads')
    im = make_input('rand', nB, H, W, seed)
    if fam == 'qarv':
        m = lvae.get_model('qarv_base')
        m.load_state_dict(O.sensitised_state_dict(O.qarv_param_shapes(), seed=0), strict=False)
        m = m.to(DEV).train()
        noise = _qres_noise(None, O.qarv_base_arch(), nB, H, W, nseed)
        st = m._forward_train(im.to(DEV), torch.tensor(lmbs, device=DEV), noise=noise)
    else:
        m = lvae.get_model('qres34m', lmb=QRES_LMB)
        m.load_state_dict(O.sensitised_state_dict(Q.qres_param_shapes(), seed=0), strict=False)
        m = m.to(DEV).train()
        noise = _qres_noise(Q, Q.qres34m_arch(), nB, H, W, nseed)
        st = m(im.to(DEV), noise=noise)
    st['loss'].backward()
    assert abs(st['loss'].item() - float(g['loss'])) <= 1e-4 * abs(float(g['loss']))
    grads = dict(m.named_parameters())
    for name, norm, dot in zip(g['names'].tolist(), g['grad_norm'], g['grad_dot']):
        gr = grads[name].grad.detach().cpu().double()
        assert abs(float(gr.norm()) - norm) <= GRAD_RTOL * norm + 1e-9, name
        pr = grad_probe(name, gr.shape).double()
        assert abs(float((gr * pr).sum()) - dot) <= GRAD_RTOL * norm * float(pr.norm()) + 1e-9, name


def test_graphed_step_with_gradient_clipping_and_ema(native_lib):
    """The tail of the reference's training step (lvae/trainer.py:374-377,395): global-norm clipping and the EMA copy's
    update, recorded in the same CUDA graph as the step.  With decay 0 the EMA copy must equal the live weights after every
    replay (the update ran, after the optimizer); the gradients left in .grad are the clipped ones."""
    import lvae
    from lvae.training import GraphedTrainStep
    torch.manual_seed(0)
    m = lvae.get_model('qarv_base').to(DEV).train()
    ema = copy.deepcopy(m).eval()
    p0 = [p.detach().clone() for p in m.parameters()]
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, capturable=True)
    step = GraphedTrainStep(m, opt, (2, 3, 64, 64), warmup=1, grad_clip=0.5, ema=ema, ema_decay=0.0)
    im = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(8)).to(DEV)
    for _ in range(3):
        assert np.isfinite(float(step(im)))
        assert all(torch.equal(e, p) for e, p in zip(ema.parameters(), m.parameters()))
    # the clip coefficient is applied inside the fused update (the gradients in .grad stay as reduced); the norm it used:
    gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters()))
    assert float(gn) > 0 and abs(float(step.grad_norm) - float(gn)) <= 1e-5 * float(gn)
    assert sum(int(not torch.equal(a, p)) for a, p in zip(p0, m.parameters())) > 800
    with torch.no_grad():
        assert np.isfinite(ema(im, lmb=torch.full((2,), 256.0, device=DEV))['loss'].item())


@pytest.mark.parametrize('n,max_norm,with_ema', [(100003, 0.7, True), (4096, 0.0, False), (1 << 20, 1e9, True)])
def test_fused_clip_adam_ema_matches_torch(native_lib, n, max_norm, with_ema):
    """lvae_adam_clip_ema (csrc/optim.cu) against clip_grad_norm_ + torch.optim.Adam + lerp_ (lvae/trainer.py:360-377,
    394-406) over four updates on flat buffers; n % 4 != 0 exercises the scalar tail."""
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g).to(DEV)
    e0 = torch.randn(n, generator=g).to(DEV)
    grads = [(torch.randn(n, generator=g) * 10 ** float(torch.randn((), generator=g))).to(DEV) for _ in range(4)]
    lr, betas, eps, decay = 3e-4, (0.9, 0.999), 1e-8, 0.99
    # torch
    pt = torch.nn.Parameter(p0.clone())
    et = e0.clone()
    opt = torch.optim.Adam([pt], lr=lr, betas=betas, eps=eps)
    norms_t = []
    for gr in grads:
        pt.grad = gr.clone()
        if max_norm > 0:
            norms_t.append(float(torch.nn.utils.clip_grad_norm_([pt], max_norm)))
        else:
            norms_t.append(float(gr.norm()))
        opt.step()
        et.copy_(decay * et + (1.0 - decay) * pt.detach())          # timm ModelEmaV2._update (lvae/trainer.py:377)
    # native
    pn, en = p0.clone(), e0.clone()
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    scratch = torch.zeros(native_lib.lvae_optim_scratch_doubles(), dtype=torch.float64, device=DEV)
    lr_t, step_t, dec_t = torch.tensor(lr, device=DEV), torch.zeros((), device=DEV), torch.tensor([decay, 1.0 - decay], device=DEV)
    gn = torch.zeros((), device=DEV)
    for i, gr in enumerate(grads):
        step_t.add_(1.0)
        rc = native_lib.lvae_adam_clip_ema(pn.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), en.data_ptr() if with_ema else 0, n,
                                           scratch.data_ptr(), max_norm, lr_t.data_ptr(), step_t.data_ptr(), dec_t.data_ptr(),
                                           betas[0], betas[1], eps, gn.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
        assert abs(float(gn) - norms_t[i]) <= 2e-6 * norms_t[i]
    assert float((pn - pt.detach()).abs().max()) <= 2e-6 * lr / 3e-4 + 1e-6 * float(pt.detach().abs().max()) * 0 + 5e-7
    st = opt.state[pt]
    # moments: relative to the largest entry (an entry can cancel to ~0 while its terms do not)
    assert torch.allclose(m, st['exp_avg'], rtol=1e-5, atol=3e-7 * float(st['exp_avg'].abs().max()))
    assert torch.allclose(v, st['exp_avg_sq'], rtol=1e-5, atol=3e-7 * float(st['exp_avg_sq'].abs().max()))
    if with_ema:
        assert torch.allclose(en, et, rtol=1e-6, atol=1e-6)
    else:
        assert torch.equal(en, e0)


def test_train_gradients_on_the_tensor_core_weight_gradient_path(native_lib, sensitised_sd):
    """every other gradient test uses 64 x 64 images (M <= 512 pixels per layer), where block_backward takes the
    torch.mm fallback for the weight gradients.  At 128 x 128 with B = 4 the H/4 and H/8 stages have M = 4096 / 1024
    pixels: lvae_split_planes_t_ex (+ gelu, + column sums) and the split-K lvae_gemm_wgrad are what runs -- the path
    real training shapes take -- and must agree with autograd over the oracle."""
    import lvae
    m = lvae.get_model('qarv_base')
    m.load_state_dict(sensitised_sd, strict=False)
    m = m.to(DEV).train()
    B, H, W = 4, 128, 128
    im = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(15))
    lmb = torch.tensor([32.0, 256.0, 1024.0, 2048.0])
    noise = _noise(m, B, H, W, 29)
    launches = []
    orig = m.train_path._wgrad
    m.train_path._wgrad = lambda *a, **k: (launches.append(1), orig(*a, **k))[1]
    st = m._forward_train(im.to(DEV), lmb.to(DEV), noise=noise)
    st['loss'].backward()
    assert len(launches) >= 40, len(launches)                  # two tensor-core weight gradients per block with M >= 1024
    ref, grads = _oracle_grads(O.qarv_forward, sensitised_sd, im, lmb, mode='train', noise=noise)
    worst = _compare(m, grads, st['loss'].item(), ref['loss'].item())
    print('worst relative gradient error (tensor-core weight gradients)', worst)


def test_gpu_crop_flip_matches_indexing(native_lib):
    """lvae.training.gpu_random_crop_flip (RandomCrop + RandomHorizontalFlip + ToTensor of lvae/datasets/image.py:45-56 on the
    device) against plain tensor indexing with the origins / flags it drew."""
    from lvae.training import gpu_random_crop_flip
    g = torch.Generator(device=DEV).manual_seed(3)
    src = torch.randint(0, 256, (5, 3, 300, 340), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)).to(DEV)
    out, (y0, x0, flip) = gpu_random_crop_flip(src, 256, generator=g)
    assert out.shape == (5, 3, 256, 256) and out.dtype == torch.float32
    assert int(flip.sum()) not in (0, 5) or True
    src_cpu = src.cpu()
    for b in range(5):
        # torchvision's to_tensor divides on the CPU (IEEE division; torch's CUDA `/ 255` multiplies by the reciprocal)
        ref = src_cpu[b, :, int(y0[b]):int(y0[b]) + 256, int(x0[b]):int(x0[b]) + 256].float().div(255)
        if int(flip[b]):
            ref = ref.flip(-1)
        assert torch.equal(out[b].cpu(), ref)


def test_autographed_forward_backward_equals_the_eager_step(native_lib):
    """lvae.training.AutoGraphedTrain (opt-in): `model(batch)['loss'].backward()` as two CUDA-graph replays gives the loss and
    the gradients of the eager step -- same seed, same noise draws, every parameter -- on the first call (capture) and on later
    replays with other batches and updated weights; another shape falls back to the eager path."""
    import lvae
    torch.manual_seed(0)
    m = lvae.get_model('qarv_base').to(DEV).train()
    ims = [make_input('synth', 2, 64, 64, 500 + i).to(DEV) for i in range(3)]
    lmb = torch.tensor([64.0, 1024.0], device=DEV)
    params = [p for p in m.parameters() if p.requires_grad]

    def step(im, seed):
        torch.manual_seed(seed)
        for p in params:
            p.grad = None
        out = m(im, lmb=lmb)
        out['loss'].backward()
        return out['loss'].item(), out['bppix'], out['psnr'], [None if p.grad is None else p.grad.clone() for p in params]

    eager = [step(im, 10 + i) for i, im in enumerate(ims)]
    m.train_path.autograph_enabled = True
    try:
        for rnd in range(2):                                      # round 0 captures on ims[0]; round 1 only replays
            for i, im in enumerate(ims):
                got = step(im, 10 + i)
                assert m.train_path.autograph.core is not None
                same_noise = got[0] == eager[i][0]
                assert abs(got[0] - eager[i][0]) <= (1e-6 if same_noise else 5e-2) * abs(eager[i][0]), (rnd, i, got[0], eager[i][0])
                if same_noise:
                    assert got[1] == eager[i][1] and got[2] == eager[i][2]
                    for g, e in zip(got[3], eager[i][3]):
                        assert (g is None) == (e is None)
                        if g is not None:
                            assert torch.allclose(g, e, rtol=1e-5, atol=1e-7 * float(e.abs().max()) + 1e-12), float((g - e).abs().max())
        # p.grad never aliases the graphs' static gradient buffers (AccumulateGrad copies): the next replay cannot overwrite it,
        # and accumulating two backward passes adds up
        ag = m.train_path.autograph
        static = {g.data_ptr() for g in ag.s_grads if g is not None}
        assert all(p.grad is None or p.grad.data_ptr() not in static for p in params)
        one = [None if p.grad is None else p.grad.clone() for p in params]
        torch.manual_seed(10 + len(ims) - 1)
        m(ims[-1], lmb=lmb)['loss'].backward()                      # second backward into the same .grad: the sum
        for p, g1 in zip(params, one):
            if g1 is not None:
                assert torch.allclose(p.grad, 2 * g1, rtol=1e-5, atol=1e-7 * float(g1.abs().max()) + 1e-12)
        # an evaluation pass in between (eval plans, their own weight packing) leaves the captured step intact
        for p in params:
            p.grad = None
        m.eval()
        with torch.no_grad():
            ev = m(ims[1], lmb=lmb)
        assert np.isfinite(ev['bppix'])
        m.train()
        again = step(ims[0], 10)
        assert abs(again[0] - eager[0][0]) <= (1e-6 if again[0] == eager[0][0] else 5e-2) * abs(eager[0][0])
        # a loss whose forward is no longer the latest one cannot be differentiated through the graphs: loud error
        torch.manual_seed(1)
        first = m(ims[0], lmb=lmb)['loss']
        m(ims[1], lmb=lmb)
        with pytest.raises(RuntimeError, match='earlier forward'):
            first.backward()
        for p in params:
            p.grad = None
        # weights move -> the captured forward re-packs them: the loss changes like the eager one does
        with torch.no_grad():
            for p in params:
                p.mul_(1.01)
        g2 = step(ims[0], 10)
        m.train_path.autograph_enabled = False
        e2 = step(ims[0], 10)
        assert abs(g2[0] - e2[0]) <= 5e-2 * abs(e2[0]) and g2[0] != eager[0][0]
        m.train_path.autograph_enabled = True
        other = make_input('synth', 1, 64, 128, 9).to(DEV)         # another shape: eager path, no new capture
        out = m(other, lmb=lmb[:1])
        out['loss'].backward()
        assert m.train_path.autograph.shape == (2, 3, 64, 64)
    finally:
        m.train_path.autograph_enabled = False
