"""rd model (continuous-posterior R(D) model, BASELINE configs[4]) on a B200 against the fixtures of the unmodified
reference and the live CPU oracle.  The reference samples z = qm + qv * randn also in eval; the same noise values
are injected here, so the comparison is deterministic.  Tolerances: bpp 1e-4 relative to ... see asserts."""
import numpy as np
import pytest
import torch

import lvae_oracle as O
import rd_oracle as R
from oracle_inputs import RD_CASES, make_input

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def rd_sd():
    return O.sensitised_state_dict(R.rd_param_shapes(), seed=0, wide_heads=False)


@pytest.fixture(scope='module')
def rd_model(native_lib, rd_sd):
    import lvae
    torch.manual_seed(0)
    model = lvae.get_model('rd_model_base')
    model.load_state_dict(rd_sd, strict=True)
    return model.to(DEV).eval()


@pytest.mark.parametrize('precision', ['f16x3', 'bf16x6', 'fp32'])
@pytest.mark.parametrize('name', list(RD_CASES))
def test_rd_forward_matches_reference_fixture(name, precision, rd_model, golden):
    g = golden(name)
    kind, nB, H, W, lmbs, seed, nseed = RD_CASES[name]
    im = make_input(kind, nB, H, W, seed)
    lmb = torch.tensor(lmbs, device=DEV)
    noise = R.draw_noise(R.latent_shapes(R.rd_base_arch(), nB, H, W), nseed)
    rd_model.precision = precision
    try:
        st = rd_model(im.to(DEV), lmb=lmb, return_rec=True, noise=noise)
        x_hat, lat = rd_model.forward_end2end(im.to(DEV), lmb, get_latents=True, noise=noise)
    finally:
        rd_model.precision = 'f16x3'
    # the KL is a smooth function of fp32 activations (no quantisation): everything agrees to fp32 round-off
    # accumulated over ~110 blocks.  bpp ~ 19 here, so 1e-4 absolute is 5e-6 relative.
    assert abs(st['bppix'] - float(g['bppix'])) <= 1e-4, (st['bppix'], float(g['bppix']))
    assert abs(st['psnr'] - float(g['psnr'])) <= 0.01
    assert abs(st['loss'].item() - float(g['loss'])) <= 2e-5 * abs(float(g['loss']))
    assert (st['im_hat'].cpu() - torch.from_numpy(g['im_hat'])).abs().max().item() < 2e-5
    for li, s in enumerate(lat):
        kl = s['kl'].sum(dim=(1, 2, 3)).cpu().numpy()
        assert np.allclose(kl, g['kl_per_image'][li], rtol=2e-5, atol=1e-3), li
    assert (lat[0]['z'].cpu() - torch.from_numpy(g['z0'])).abs().max().item() < 2e-5
    assert (lat[14]['z'].cpu() - torch.from_numpy(g['z14'])).abs().max().item() < 5e-5


def test_rd_against_live_oracle_and_batch_invariance(rd_model, rd_sd):
    """BASELINE configs[4] shape (256x256): a batch of 4 equals four single-image runs bit for bit (what makes the
    8-GPU batch shard exact), and matches the oracle run on this host."""
    H = W = 256
    im = torch.rand(4, 3, H, W, generator=torch.Generator().manual_seed(31))
    lmb = torch.tensor([8.0, 64.0, 512.0, 2048.0])
    noise = R.draw_noise(R.latent_shapes(R.rd_base_arch(), 4, H, W), 9)
    full = rd_model.engine.run(im.to(DEV), lmb.to(DEV), mode='eval', noise=noise)['stats_host']
    for b in (0, 3):
        one = rd_model.engine.run(im[b:b + 1].to(DEV), lmb[b:b + 1].to(DEV), mode='eval',
                                  noise=[n[b:b + 1] for n in noise])['stats_host']
        assert one[4] == full[4 + b] and one[5] == full[4 + 4 + b]
    ref = R.rd_forward(rd_sd, im[:2], lmb[:2], [n[:2] for n in noise])
    kl_dim = full[4:6]                                  # nats per dimension, images 0 and 1
    assert np.allclose(kl_dim, ref['kl_per_image'].numpy(), rtol=1e-5)
    assert np.allclose(full[8:10], ref['mse_per_image'].numpy(), rtol=2e-5)


def test_rd_default_noise_and_sampling(rd_model):
    im = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(2)).to(DEV)
    torch.manual_seed(1)
    a = rd_model(im, lmb=torch.tensor([100.0], device=DEV))
    torch.manual_seed(1)
    b = rd_model(im, lmb=torch.tensor([100.0], device=DEV))
    assert a['bppix'] == b['bppix'] and a['psnr'] == b['psnr']          # device generator, reproducible under a seed
    x_hat, lat = rd_model.forward_end2end(im, 100.0, get_latents=True)
    rec = rd_model.conditional_sample(100.0, [s['z'] for s in lat])
    assert (rec - rd_model.process_output(x_hat)).abs().max().item() < 1e-5
    out = rd_model.unconditional_sample(100.0, bhw_repeat=(2, 1, 1), t=0.0)
    assert out.shape == (2, 3, 64, 64) and bool(torch.isfinite(out).all())
