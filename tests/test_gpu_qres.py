"""End-to-end parity of the qres34m path on a B200 against the fixtures produced by the
unmodified reference (tests/golden/qres_*.npz, oracle/gen_golden.py): eval forward, train forward with the
reference's noise, integer symbols / table indexes, bit streams and decompression."""
import numpy as np
import pytest
import torch

import lvae_oracle as O
import qres_oracle as Q
from oracle_inputs import QRES_CASES, QRES_LMB, make_input
from conftest import parity_log
from test_gpu_model import PSNR_TOL, bpp_tol, _check_integer_parity
from test_oracle_pinned import _qres_noise

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def qres_sd():
    return O.sensitised_state_dict(Q.qres_param_shapes(), seed=0)


@pytest.fixture(scope='module')
def qres_model(native_lib, qres_sd):
    import lvae
    torch.manual_seed(0)
    model = lvae.get_model('qres34m', lmb=QRES_LMB)
    missing, unexpected = model.load_state_dict(qres_sd, strict=False)
    assert not unexpected and all('discrete_gaussian' in k for k in missing), (missing, unexpected)
    model = model.to(DEV).eval()
    model.compress_mode()
    return model


@pytest.mark.parametrize('precision', ['f16x3', 'bf16x6', 'fp32'])
@pytest.mark.parametrize('name', list(QRES_CASES))
def test_qres_forward_matches_reference_fixture(name, precision, qres_model, qres_sd, golden):
    g = golden(name)
    kind, nB, H, W, seed, nseed = QRES_CASES[name]
    im_cpu = make_input(kind, nB, H, W, seed)
    im = im_cpu.to(DEV)
    qres_model.precision = precision
    try:
        st = qres_model(im, return_rec=True)
        lat = qres_model.forward_get_latents(im)
        obj = qres_model.compress(im)
        P = qres_model.engine._plans[(nB, H, W, 'compress', False)]
        torch.cuda.synchronize()
        syms, idxs = [s.cpu() for s in P.sym], [i.cpu() for i in P.idx]
        rec = qres_model.decompress(obj)
        qres_model.train()
        noise = _qres_noise(Q, Q.qres34m_arch(), nB, H, W, nseed)
        with torch.no_grad():       # the launch-plan train forward; with autograd recording see tests/test_gpu_train.py
            tr = qres_model(im, noise=noise)
    finally:
        qres_model.eval()
        qres_model.precision = 'f16x3'
    assert abs(st['bppix'] - float(g['bppix'])) <= bpp_tol(H, W), (st['bppix'], float(g['bppix']))
    assert abs(st['psnr'] - float(g['psnr'])) <= PSNR_TOL
    assert abs(st['loss'].item() - float(g['loss'])) <= 1e-4 * abs(float(g['loss']))
    assert abs(st['mse'] - float(g['mse'])) <= 1e-4 * float(g['mse']) and abs(st['kl'] - float(g['kl'])) <= 1e-4
    flips = _check_integer_parity(
        syms, idxs, [torch.from_numpy(g[f'sym{li}'].astype(np.int32)) for li in range(12)],
        [torch.from_numpy(g[f'idx{li}'].astype(np.int32)) for li in range(12)],
        lambda: Q.qres_forward(qres_sd, im_cpu, QRES_LMB)['records'], scale_table=Q.qres_scale_table())
    parity_log(test='qres34m fixture (unmodified reference)', case=name, precision=precision, symbols=sum(s_.numel() for s_ in syms),
               flips=flips, dbpp=abs(st['bppix'] - float(g['bppix'])), dpsnr=abs(st['psnr'] - float(g['psnr'])), bpp_tol=bpp_tol(H, W))
    tol_nats = bpp_tol(H, W) * H * W / 1.4427
    for li, stl in enumerate(lat):
        kl = stl['kl'].sum(dim=(1, 2, 3)).cpu().numpy()
        assert np.all(np.abs(kl - g['kl_per_image'][li]) <= tol_nats + 2e-5 * g['kl_per_image'][li]), li
    assert tuple(obj[-1]) == tuple(g['shape'])
    assert (rec - st['im_hat']).abs().max().item() < 1e-5          # decoder reproduces the encoder's reconstruction
    if flips == 0:
        assert (st['im_hat'].cpu() - torch.from_numpy(g['im_hat'])).abs().max().item() < 1e-5
        for li, stl in enumerate(lat):
            assert (stl['z'].cpu() - torch.from_numpy(g[f'z{li}'])).abs().max().item() < 2e-5, f'layer {li} latents differ'
            for b in range(nB):
                assert obj[li][b] == g[f'bytes{li}_{b}'].tobytes(), f'bit stream of layer {li} image {b} differs'
        assert (rec.cpu() - torch.from_numpy(g['dec_im_hat'])).abs().max().item() < 1e-5
    # training branch (uniform noise + gaussian_log_prob_mass): smooth in the activations -> fp32 round-off only
    assert abs(tr['loss'].item() - float(g['train_loss'])) <= 2e-5 * abs(float(g['train_loss']))
    assert abs(tr['bppix'] - float(g['train_bppix'])) <= 1e-4 and abs(tr['psnr'] - float(g['train_psnr'])) <= PSNR_TOL


def test_qres_batched_container_roundtrip_and_errors(qres_model, tmp_path):
    from PIL import Image
    im = make_input('synth', 3, 64, 128, 30).to(DEV)
    obj = qres_model.compress(im)
    assert len(obj) == 13 and all(len(layer) == 3 for layer in obj[:-1]) and obj[-1] == (3, 384, 1, 2)
    rec = qres_model.decompress(obj)
    ref = qres_model(im, return_rec=True)['im_hat']
    assert (rec - ref).abs().max().item() < 1e-5
    one = qres_model.compress(im[1:2])
    assert all(one[li][0] == obj[li][1] for li in range(12))      # batch invariance of the bit stream
    # file container: pickle of the list + (h, w), cropped on decode (qresvae/model.py:689-725)
    arr = (make_input('synth', 1, 100, 150, 8)[0].permute(1, 2, 0).numpy() * 255).round().astype(np.uint8)
    src, bits = tmp_path / 'x.png', tmp_path / 'x.bits'
    Image.fromarray(arr).save(src)
    qres_model.compress_file(src, bits)
    out = qres_model.decompress_file(bits)
    assert tuple(out.shape) == (1, 3, 100, 150)
    mse = ((out.cpu()[0].permute(1, 2, 0).numpy() - arr / 255.0) ** 2).mean()
    assert np.isfinite(mse)
    with pytest.raises(AssertionError):
        qres_model(torch.rand(1, 3, 60, 64, device=DEV))
    with pytest.raises(AssertionError):
        qres_model(torch.rand(1, 3, 64, 64, device=DEV) * 2)


def test_qres_forward_stream_equals_forward(qres_model):
    """HierarchicalVAE.forward_stream: same numbers as forward(), per-layer rate log included."""
    ims = [make_input('synth', 2, 64, 128, 400 + i) for i in range(3)]
    want, logs = [], []
    for im in ims:
        want.append(qres_model(im.to(DEV)))
        logs.append(list(qres_model._stats_log['eval_bppix']))
    got = []
    for i, g in enumerate(qres_model.forward_stream([im.pin_memory() for im in ims])):
        got.append(g)
        np.testing.assert_allclose(qres_model._stats_log['eval_bppix'], logs[i], rtol=1e-6)
    for g, w_ in zip(got, want):
        assert g['loss'] == w_['loss'].item() and g['bppix'] == w_['bppix'] and g['psnr'] == w_['psnr'] and g['kl'] == w_['kl']
        assert g['mse'] == w_['mse']


def test_qres_sampling_with_given_latents_reproduces_decoder(qres_model):
    im = make_input('rand', 1, 64, 64, 31).to(DEV)
    lat = qres_model.forward_get_latents(im)
    rec = qres_model.cond_sample([s['z'] for s in lat])
    ref = qres_model(im, return_rec=True)['im_hat']
    assert (rec - ref).abs().max().item() < 1e-5
    smp = qres_model.uncond_sample((2, 1, 1), temprature=0.0)
    assert tuple(smp.shape) == (2, 3, 64, 64) and bool(torch.isfinite(smp).all())


def test_qres_against_live_oracle_at_config_shape_and_batch16_invariance(qres_model, qres_sd):
    """BASELINE configs[2] shape: qres34m at 512 x 768 against the oracle run live on the host CPU
    (north-star tolerances unwidened: bpp 1e-4, PSNR 0.01 dB, symbols / indexes one by one), and the same image inside a
    batch of 16 -- the benched batch size -- must give bit-identical symbols, indexes and per-image statistics."""
    H, W = 512, 768
    im1 = make_input('synth', 1, H, W, 44)
    ref = Q.qres_forward(qres_sd, im1, QRES_LMB)
    st = qres_model(im1.to(DEV), return_rec=True)
    assert bpp_tol(H, W) == 1e-4
    assert abs(st['bppix'] - ref['bppix']) <= 1e-4, (st['bppix'], ref['bppix'])
    assert abs(st['psnr'] - ref['psnr']) <= PSNR_TOL
    qres_model.compress(im1.to(DEV))
    P = qres_model.engine._plans[(1, H, W, 'compress', False)]
    torch.cuda.synchronize()
    syms, idxs = [s.cpu() for s in P.sym], [i.cpu() for i in P.idx]
    flips = _check_integer_parity(syms, idxs, [r['sym'] for r in ref['records']], [r['idx'] for r in ref['records']],
                                  lambda: ref['records'], scale_table=Q.qres_scale_table())
    parity_log(test='qres34m live oracle, BASELINE configs[2] image size', case='synth 1x512x768', precision=qres_model.precision,
               symbols=sum(s_.numel() for s_ in syms), flips=flips, dbpp=abs(st['bppix'] - ref['bppix']),
               dpsnr=abs(st['psnr'] - ref['psnr']), bpp_tol=1e-4)
    if flips == 0:
        assert (st['im_hat'].cpu() - ref['im_hat']).abs().max().item() < 1e-5
    # batch 16: image 5 of the batch is the image above
    im16 = make_input('rand', 16, H, W, 45)
    im16[5] = im1[0]
    qres_model.compress(im16.to(DEV))
    P16 = qres_model.engine._plans[(16, H, W, 'compress', False)]
    torch.cuda.synchronize()
    for li in range(len(syms)):
        assert torch.equal(P16.sym[li][5].cpu(), syms[li][0]) and torch.equal(P16.idx[li][5].cpu(), idxs[li][0]), li
    lat16 = qres_model.forward_get_latents(im16.to(DEV))
    lat1 = qres_model.forward_get_latents(im1.to(DEV))
    for a, b in zip(lat16, lat1):
        assert torch.equal(a['z'][5], b['z'][0]) and torch.equal(a['kl'][5], b['kl'][0])


# ----------------------------------------------------------------------------- qres34m_lossless
@pytest.fixture(scope='module')
def lossless_model(native_lib):
    import lvae
    torch.manual_seed(0)
    model = lvae.get_model('qres34m_lossless')
    sd = O.sensitised_state_dict(Q.qres_param_shapes(Q.qres34m_lossless_arch()), seed=0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all('discrete_gaussian' in k for k in missing), (missing, unexpected)
    model = model.to(DEV).eval()
    model.compress_mode()
    return model, sd


def test_lossless_forward_and_codec_match_reference_fixture(lossless_model, golden):
    """GaussianNLLOutputNet path (reference qresvae/model.py:16-94): loss = kl + nll against the fixture of the unmodified
    reference, the compressed object (12 latent layers + feature shape + the image's own residual stream) byte for byte,
    and decompress(compress(x)) == x exactly on 8-bit images."""
    from oracle_inputs import LOSSLESS_CASES, make_input_8bit
    model, sd = lossless_model
    name = 'qresll_synth_2x64x128'
    g = golden(name)
    kind, nB, H, W, seed, nseed = LOSSLESS_CASES[name]
    im_cpu = make_input_8bit(kind, nB, H, W, seed)
    im = im_cpu.to(DEV)
    st = model(im, return_rec=True)
    assert abs(st['loss'].item() - float(g['loss'])) <= 1e-5 * abs(float(g['loss']))
    assert abs(st['nll'] - float(g['nll'])) <= 1e-5 * float(g['nll'])
    assert abs(st['bppix'] - float(g['bppix'])) <= bpp_tol(H, W) and abs(st['psnr'] - float(g['psnr'])) <= PSNR_TOL
    assert (st['im_hat'].cpu() - torch.from_numpy(g['im_hat'])).abs().max().item() < 1e-5
    obj = model.compress(im)
    assert len(obj) == 14 and tuple(obj[-2]) == tuple(g['shape']) and len(obj[-1]) == nB
    same = all(obj[li][b] == g[f'bytes{li}_{b}'].tobytes() for li in range(12) for b in range(nB))
    rec = model.decompress(obj)
    assert torch.equal((rec.cpu() * 255).round(), (im_cpu * 255).round())           # lossless, whatever the latents did
    # the residual stream: its means round(p_mean * 127.5 + 127.5) and scale-table indexes come out of two GEMMs, so -- like
    # the latents' symbols -- they can differ from the CPU reference at rounding boundaries.  Compare them element by
    # element with the oracle's; when none differs the stream must be the reference's byte for byte.
    fw = Q.qres_forward(sd, im_cpu, 0.0, Q.qres34m_lossless_arch())
    pm_o, plogv_o, x_o = Q._prepare_codec(sd, fw['feature'], (im_cpu - 0.5) * 2.0)
    idx_o = O.build_indexes(torch.exp(plogv_o), scale_table=Q.lossless_scale_table())
    P = model.engine._plans[(nB, H, W, 'compress', False)]
    torch.cuda.synchronize()
    d_pm = (P.on_pm.cpu().view_as(pm_o) != pm_o)
    d_idx = (P.on_idx.cpu().view_as(idx_o) != idx_o)
    n_diff = int(d_pm.sum()) + int(d_idx.sum())
    assert n_diff <= 1e-3 * pm_o.numel(), (int(d_pm.sum()), int(d_idx.sum()))
    if d_pm.any():      # only where p_mean * 127.5 + 127.5 sits on a .5 boundary (the two sides differ by exactly one bin)
        assert float((P.on_pm.cpu().view_as(pm_o) - pm_o)[d_pm].abs().max()) == 1.0
    if same and n_diff == 0:
        for b in range(nB):
            assert obj[-1][b] == g[f'final_bytes_{b}'].tobytes(), b
        assert (rec.cpu() - torch.from_numpy(g['dec_im_hat'])).abs().max().item() < 1e-6
    print(f'lossless residual stream: {int(d_pm.sum())} mean and {int(d_idx.sum())} index differences of {pm_o.numel()}')
    parity_log(test='qres34m_lossless fixture (unmodified reference)', case=name, precision=model.precision,
               symbols=nB * 3 * H * W, flips=n_diff, dbpp=abs(st['bppix'] - float(g['bppix'])),
               dpsnr=abs(st['psnr'] - float(g['psnr'])), bpp_tol=bpp_tol(H, W))
    # training branch with the reference's noise (launch-plan forward, no autograd) and with autograd recording
    model.train()
    try:
        noise = _qres_noise(Q, Q.qres34m_lossless_arch(), nB, H, W, nseed)
        with torch.no_grad():
            tr = model(im, noise=noise)
        assert abs(tr['loss'].item() - float(g['train_loss'])) <= 2e-5 * abs(float(g['train_loss']))
        tg = model(im, noise=noise)
        assert tg['loss'].requires_grad and abs(tg['loss'].item() - float(g['train_loss'])) <= 2e-5 * abs(float(g['train_loss']))
        tg['loss'].backward()
        gmean = model.out_net.conv_mean[0].weight.grad
        assert gmean is not None and bool(torch.isfinite(gmean).all()) and float(gmean.abs().max()) > 0
    finally:
        model.eval()
        model.zero_grad(set_to_none=True)


def test_lossless_file_roundtrip_and_live_oracle_on_a_photo_shaped_image(lossless_model, tmp_path):
    """compress_file / decompress_file on a non-aligned 8-bit PNG (pad + crop, pickle container) reproduce the file exactly;
    a 192 x 256 forward agrees with the oracle run live on the host."""
    from PIL import Image
    from oracle_inputs import make_input_8bit
    model, sd = lossless_model
    arr = (make_input('synth', 1, 100, 150, 8)[0].permute(1, 2, 0).numpy() * 255).round().astype(np.uint8)
    src, bits = tmp_path / 'x.png', tmp_path / 'x.bits'
    Image.fromarray(arr).save(src)
    model.compress_file(src, bits)
    out = model.decompress_file(bits)
    assert tuple(out.shape) == (1, 3, 100, 150)
    assert np.array_equal((out.cpu()[0].permute(1, 2, 0).numpy() * 255).round().astype(np.uint8), arr)
    im = make_input_8bit('synth', 1, 192, 256, 13)
    ref = Q.qres_forward(sd, im, 0.0, Q.qres34m_lossless_arch())
    st = model(im.to(DEV))
    assert abs(st['loss'].item() - ref['loss'].item()) <= 1e-5 * abs(ref['loss'].item())
    assert abs(st['nll'] - ref['mse']) <= 1e-5 * ref['mse'] and abs(st['bppix'] - ref['bppix']) <= bpp_tol(192, 256)
    smp = model.uncond_sample((1, 1, 1), temprature=0.5)
    assert tuple(smp.shape) == (1, 3, 64, 64) and bool(torch.isfinite(smp).all())
