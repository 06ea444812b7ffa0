"""Pins oracle/lvae_oracle.py: (1) against the committed golden fixtures (made by the UNMODIFIED
reference through oracle/shims, see oracle/gen_golden.py) -- runs everywhere; (2) against the
reference itself when /root/reference is present (build container only); (3) against the
known answers of SURVEY.md Appendix A.  CPU only."""
import math

import numpy as np
import pytest
import torch

import lvae_oracle as O
import ref_loader
from oracle_inputs import CASES, make_input


@pytest.fixture(scope='module')
def sd(sensitised_sd):
    return sensitised_sd


def test_param_shapes_count_matches_readme():
    # lvae/models/qarv/README.md:18 -- 93.4 M parameters (93.433 M, 907 tensors)
    shapes = O.qarv_param_shapes()
    n = sum(int(np.prod(s)) for _, s in shapes)
    assert len(shapes) == 907
    assert round(n / 1e6, 3) == 93.433


def test_entropy_kat_appendix_a(golden):
    k = golden('entropy_kat')
    qm, pm, pv = (torch.from_numpy(k[n]) for n in ('qm', 'pm', 'pv'))
    z, P = O.eval_quantize_likelihood(qm, pm, pv)
    assert torch.equal(z, torch.from_numpy(k['z']))
    assert torch.equal(P, torch.from_numpy(k['P']))
    assert torch.equal(O.symbols(qm, pm), torch.from_numpy(k['sym']))
    assert torch.equal(O.build_indexes(pv), torch.from_numpy(k['idx']))
    # hand-checked rows of SURVEY Appendix A (half-to-even, floors, index mapping)
    assert O.symbols(qm, pm).tolist() == [0, 0, 2, 2, 0, -2, 3, -8, 0, 12, 0, 40]
    assert O.build_indexes(pv).tolist() == [27, 27, 27, 8, 0, 19, 36, 0, 13, 4, 63, 63]
    kl = -torch.log(P)
    assert abs(kl[0].item() - 0.9599164) < 1e-6 and abs(kl[3].item() - 20.7232666) < 1e-5
    lp = O.gaussian_log_prob_mass(torch.zeros(5), torch.from_numpy(k['train_scale']), torch.from_numpy(k['train_x']))
    assert torch.equal(lp, torch.from_numpy(k['train_logp']))
    assert abs(lp[1].item() + 5.1198306) < 1e-6 and abs(lp[3].item() + 41.4189377) < 1e-5


def test_prior_floor():
    pm, pv = O.prior_transform(torch.tensor([0.0, -1e9, 0.0, 0.0]).view(1, 4, 1, 1))
    assert abs(pv.flatten()[0].item() - 1.10025895) < 1e-6      # plogv_raw = 0
    pm, pv = O.prior_transform(torch.tensor([0.0, -1e9]).view(1, 2, 1, 1))
    assert abs(pv.item() - 0.100258850) < 1e-7                  # floor e^-2.3


def test_cdf_tables_match_reference_fixture(golden):
    k = golden('entropy_kat')
    cdf, length, offset = O.build_cdf_tables()
    assert torch.equal(cdf, torch.from_numpy(k['cdf']))
    assert torch.equal(length, torch.from_numpy(k['cdf_length']))
    assert torch.equal(offset, torch.from_numpy(k['offset']))
    assert cdf[0, :5].tolist() == [0, 1, 65534, 65535, 65536] and offset[0].item() == -1
    assert torch.equal(torch.from_numpy(k['scale_table']), O.default_scale_table())


@pytest.mark.parametrize('name', ['qarv_rand_1x64x64', 'qarv_synth_3x64x128', 'qarv_rand_2x128x192'])
def test_oracle_forward_matches_golden(name, golden, sd):
    g = golden(name)
    kind, nB, H, W, lmbs, seed = CASES[name]
    im = make_input(kind, nB, H, W, seed)
    out = O.qarv_forward(sd, im, torch.tensor(lmbs))
    assert np.float32(out['loss'].item()) == g['loss']
    assert out['bppix'] == float(g['bppix']) and out['mse'] == float(g['mse']) and out['psnr'] == float(g['psnr'])
    assert torch.equal(out['im_hat'], torch.from_numpy(g['im_hat']))
    for li, r in enumerate(out['records']):
        assert torch.equal(r['z'], torch.from_numpy(g[f'z{li}']))
        assert torch.equal(r['sym'], torch.from_numpy(g[f'sym{li}'].astype(np.int32)))
        assert torch.equal(r['idx'], torch.from_numpy(g[f'idx{li}'].astype(np.int32)))
        assert torch.equal(r['kl'].sum(dim=(1, 2, 3)), torch.from_numpy(g['kl_per_image'][li]))


def test_oracle_compress_decompress_matches_golden(golden, sd):
    name = 'qarv_rand_1x64x64'
    g = golden(name)
    kind, nB, H, W, lmbs, seed = CASES[name]
    im = make_input(kind, nB, H, W, seed)
    blob = O.qarv_compress(sd, im, lmbs[0])
    assert blob == g['bytes0'].tobytes()
    rec = O.qarv_decompress(sd, blob)
    assert torch.equal(rec, torch.from_numpy(g['dec_im_hat']))


def test_rans_python_roundtrip_with_bypass():
    tables = O.build_cdf_tables()
    g = torch.Generator().manual_seed(5)
    idx = torch.randint(0, 64, (4000,), generator=g).tolist()
    sym = torch.round(torch.randn(4000, generator=g) * 3).int().tolist()
    sym[7], sym[100], sym[3999] = 400, -300, 70000
    data = O.rans_encode(sym, idx, *tables)
    assert O.rans_decode(data, idx, *tables) == sym


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not on this machine')
def test_oracle_equals_live_reference(sd):
    """Bit-exact agreement with the reference's own Python run in this process."""
    ref = ref_loader.load_reference()
    try:
        torch.manual_seed(0)
        model = ref.get_model('qarv_base').eval()
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected
        ref_keys = [k for k, _ in model.named_parameters()]
        assert sorted(ref_keys) == sorted(k for k, _ in O.qarv_param_shapes())
        for k, s in O.qarv_param_shapes():
            assert tuple(model.state_dict()[k].shape) == tuple(s), k
        im = torch.rand(2, 3, 64, 128, generator=torch.Generator().manual_seed(11))
        lmb = torch.tensor([100.0, 1500.0])
        with torch.no_grad():
            st = model(im, lmb=lmb, return_rec=True)
        out = O.qarv_forward(sd, im, lmb)
        assert st['loss'].item() == out['loss'].item()
        assert st['bppix'] == out['bppix'] and st['psnr'] == out['psnr'] and st['mse'] == out['mse']
        assert torch.equal(st['im_hat'], out['im_hat'])
    finally:
        ref_loader.unload_reference()


# ----------------------------------------------------------------------------- rd model (continuous posterior)
def test_rd_param_count_matches_readme():
    import rd_oracle as R
    shapes = R.rd_param_shapes()       # lvae/models/rd/README.md:23 -- 186.7 M parameters
    assert round(sum(int(np.prod(s)) for _, s in shapes) / 1e6, 3) == 186.670


@pytest.mark.parametrize('name', ['rd_rand_1x64x64', 'rd_synth_2x128x128'])
def test_rd_oracle_matches_golden(name, golden):
    import rd_oracle as R
    from oracle_inputs import RD_CASES
    g = golden(name)
    kind, nB, H, W, lmbs, seed, nseed = RD_CASES[name]
    sd = O.sensitised_state_dict(R.rd_param_shapes(), seed=0, wide_heads=False)
    im = make_input(kind, nB, H, W, seed)
    arch = R.rd_base_arch()
    out = R.rd_forward(sd, im, torch.tensor(lmbs), R.draw_noise(R.latent_shapes(arch, nB, H, W), nseed))
    assert np.float32(out['loss'].item()) == g['loss']
    assert out['bppix'] == float(g['bppix']) and out['psnr'] == float(g['psnr']) and out['mse'] == float(g['mse'])
    assert torch.equal(out['im_hat'], torch.from_numpy(g['im_hat']))
    assert torch.equal(out['records'][0]['z'], torch.from_numpy(g['z0']))
    assert torch.equal(out['records'][14]['z'], torch.from_numpy(g['z14']))
    for li, r in enumerate(out['records']):
        assert torch.equal(r['kl'].sum(dim=(1, 2, 3)), torch.from_numpy(g['kl_per_image'][li]))


def test_rd_scalar_functions_known_answers():
    import rd_oracle as R
    x = torch.tensor([0.0, -0.0, 1.0, -2.0, 6.0, 6.5, -9.0])
    y = R.linear_sqrt(x)
    assert y[0] == 0 and y[2] == 1.0                                  # |x|^(1 - tanh/2) at 1 is 1
    assert abs(y[3].item() + 2.0 ** (1 - 0.5 * math.tanh(2.0))) < 1e-6
    assert abs(y[5].item() - math.sqrt(6.5 + 1e-8)) < 1e-6 and abs(y[6].item() + 3.0) < 1e-6
    assert abs(R.std_smooth(torch.tensor([0.0])).item() - 1.0) < 1e-6  # softplus_beta=ln2 (0) = ln2 / ln2
    kl = R.gaussian_kl(torch.tensor([0.3]), torch.tensor([0.5]), torch.tensor([0.1]), torch.tensor([2.0]))
    assert abs(kl.item() - (-0.5 + math.log(2.0) - math.log(0.5) + 0.5 * (0.25 + 0.04) / 4.0)) < 1e-6


# ----------------------------------------------------------------------------- qres34m (fixed rate, VDBlock heads)
def _qres_noise(Q, arch, nB, H, W, nseed):
    """What `torch.empty_like(qm).uniform_(-0.5, 0.5)` yields layer by layer after torch.manual_seed(nseed)."""
    g = torch.Generator().manual_seed(nseed)
    out, h, w = [], H // 64, W // 64
    for ent in arch['dec']:
        if ent[0] == 'lat':
            out.append(torch.empty(nB, ent[2], h, w).uniform_(-0.5, 0.5, generator=g))
        elif ent[0] == 'up':
            h, w = h * ent[3], w * ent[3]
    return out


def test_qres_param_count_matches_readme():
    import qres_oracle as Q
    shapes = Q.qres_param_shapes()     # lvae/models/qresvae/README.md: qres34m, 34.0 M parameters
    assert round(sum(int(np.prod(s)) for _, s in shapes) / 1e6, 2) == 34.04


def test_qres_cdf_tables_match_reference_fixture(golden):
    import qres_oracle as Q
    t = golden('qres_tables')
    assert np.array_equal(Q.qres_scale_table().numpy(), t['scale_table'])
    cdf, length, offset = Q.qres_tables()
    assert np.array_equal(cdf.numpy(), t['cdf']) and np.array_equal(length.numpy(), t['cdf_length'])
    assert np.array_equal(offset.numpy(), t['offset'])


@pytest.mark.parametrize('name', ['qres_rand_2x64x128', 'qres_synth_1x192x256'])
def test_qres_oracle_matches_golden(name, golden):
    import qres_oracle as Q
    from oracle_inputs import QRES_CASES, QRES_LMB
    g = golden(name)
    kind, nB, H, W, seed, nseed = QRES_CASES[name]
    sd = O.sensitised_state_dict(Q.qres_param_shapes(), seed=0)
    im = make_input(kind, nB, H, W, seed)
    out = Q.qres_forward(sd, im, QRES_LMB)
    assert np.float32(out['loss'].item()) == g['loss']
    assert out['bppix'] == float(g['bppix']) and out['psnr'] == float(g['psnr']) and out['mse'] == float(g['mse'])
    assert torch.equal(out['im_hat'], torch.from_numpy(g['im_hat']))
    for li, r in enumerate(out['records']):
        assert torch.equal(r['z'], torch.from_numpy(g[f'z{li}']))
        assert torch.equal(r['kl'].sum(dim=(1, 2, 3)), torch.from_numpy(g['kl_per_image'][li]))
        assert np.array_equal(r['sym'].numpy(), g[f'sym{li}']) and np.array_equal(r['idx'].numpy(), g[f'idx{li}'])
    noise = _qres_noise(Q, Q.qres34m_arch(), nB, H, W, nseed)
    tr = Q.qres_forward(sd, im, QRES_LMB, mode='train', noise=noise)
    assert np.float32(tr['loss'].item()) == g['train_loss'] and tr['bppix'] == float(g['train_bppix'])
    for li, r in enumerate(tr['records']):
        assert torch.equal(r['kl'].sum(dim=(1, 2, 3)), torch.from_numpy(g['train_kl_per_image'][li]))
    if H * W <= 64 * 128:          # the pure-python coder is slow: small case only
        obj = Q.qres_compress(sd, im)
        assert tuple(obj[-1]) == tuple(g['shape'])
        for li in range(12):
            for b in range(nB):
                assert obj[li][b] == g[f'bytes{li}_{b}'].tobytes()
        assert torch.equal(Q.qres_decompress(sd, obj), torch.from_numpy(g['dec_im_hat']))


# ----------------------------------------------------------------------------- training-step gradients
def _grad_case(fam):
    from gen_golden_cases import GRAD_CASES, grad_probe
    return GRAD_CASES[fam], grad_probe


@pytest.mark.parametrize('fam', ['qarv', 'qres'])
def test_oracle_autograd_matches_reference_gradients(fam, golden):
    """tests/golden/{qarv,qres}_train_grads.npz hold, from the UNMODIFIED reference's loss.backward() in train mode, every
    parameter gradient's norm and its inner product with a fixed probe: autograd over the oracle reproduces them, which
    makes the oracle the gradient reference of tests/test_gpu_train.py."""
    import qres_oracle as Q
    (nB, H, W, lmbs, seed, nseed), grad_probe = _grad_case(fam)
    g = golden(f'{fam}_train_grads')
    im = make_input('rand', nB, H, W, seed)
    if fam == 'qarv':
        sd = O.sensitised_state_dict(O.qarv_param_shapes(), seed=0)
        noise = _qres_noise(None, O.qarv_base_arch(), nB, H, W, nseed)
        fwd, args = O.qarv_forward, (im, torch.tensor(lmbs))
    else:
        sd = O.sensitised_state_dict(Q.qres_param_shapes(), seed=0)
        noise = _qres_noise(Q, Q.qres34m_arch(), nB, H, W, nseed)
        from oracle_inputs import QRES_LMB
        fwd, args = Q.qres_forward, (im, QRES_LMB)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    with torch.enable_grad():
        out = fwd.__wrapped__(sd, *args, mode='train', noise=noise)
        out['loss'].backward()
    assert abs(float(out['loss']) - float(g['loss'])) <= 1e-6 * abs(float(g['loss']))
    for name, norm, dot in zip(g['names'].tolist(), g['grad_norm'], g['grad_dot']):
        gr = sd[name].grad.double()
        assert abs(float(gr.norm()) - norm) <= 1e-4 * norm + 1e-12, name
        pr = grad_probe(name, gr.shape).double()
        assert abs(float((gr * pr).sum()) - dot) <= 1e-4 * norm * float(pr.norm()) + 1e-12, name


# ----------------------------------------------------------------------------- qres34m_lossless (GaussianNLLOutputNet)
def test_lossless_oracle_matches_golden(golden):
    """oracle/qres_oracle.py's lossless variant (qresvae/model.py:16-94, zoo.py:63-114) against the fixture
    the unmodified reference produced: loss = kl + nll, train loss with the reference's noise, every bit stream incl. the
    image's own residual stream, the decoded image, and the round trip being lossless on 8-bit input."""
    import qres_oracle as Q
    from oracle_inputs import LOSSLESS_CASES, make_input_8bit
    name = 'qresll_synth_2x64x128'
    g = golden(name)
    kind, nB, H, W, seed, nseed = LOSSLESS_CASES[name]
    arch = Q.qres34m_lossless_arch()
    sd = O.sensitised_state_dict(Q.qres_param_shapes(arch), seed=0)
    im = make_input_8bit(kind, nB, H, W, seed)
    out = Q.qres_forward(sd, im, 0.0, arch)
    assert np.float32(out['loss'].item()) == g['loss'] and out['mse'] == float(g['nll'])
    assert out['bppix'] == float(g['bppix']) and out['psnr'] == float(g['psnr'])
    assert torch.equal(out['im_hat'], torch.from_numpy(g['im_hat']))
    tr = Q.qres_forward(sd, im, 0.0, arch, mode='train', noise=_qres_noise(Q, arch, nB, H, W, nseed))
    assert np.float32(tr['loss'].item()) == g['train_loss'] and tr['mse'] == float(g['train_nll'])
    t = golden('qresll_tables')
    assert np.array_equal(Q.lossless_scale_table().numpy(), t['scale_table'])
    cdf, length, offset = Q.lossless_tables()
    assert np.array_equal(cdf.numpy(), t['cdf']) and np.array_equal(length.numpy(), t['cdf_length'])
    obj, _ = Q.lossless_compress(sd, im[:1], arch)          # one image: the pure-python coder is slow
    for li in range(12):
        assert obj[li][0] == g[f'bytes{li}_0'].tobytes()
    assert obj[-1][0] == g['final_bytes_0'].tobytes()
    dec = Q.lossless_decompress(sd, obj, arch)
    assert torch.equal(dec, torch.from_numpy(g['dec_im_hat'][:1]))
    assert torch.equal((dec * 255).round(), (im[:1] * 255).round())          # lossless


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_lossless_oracle_equals_live_reference():
    import qres_oracle as Q
    from oracle_inputs import make_input_8bit
    arch = Q.qres34m_lossless_arch()
    shapes = Q.qres_param_shapes(arch)
    sd = O.sensitised_state_dict(shapes, seed=0)
    ref = ref_loader.load_reference()
    try:
        torch.manual_seed(0)
        model = ref.get_model('qres34m_lossless').eval()
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected and all('discrete_gaussian' in k for k in missing)
        assert sorted(k for k, _ in model.named_parameters()) == sorted(k for k, _ in shapes)
        im = make_input_8bit('rand', 1, 64, 64, 5)
        with torch.no_grad():
            st = model(im)
        out = Q.qres_forward(sd, im, 0.0, arch)
        assert st['loss'].item() == out['loss'].item() and st['nll'] == out['mse'] and st['bppix'] == out['bppix']
    finally:
        ref_loader.unload_reference()
