"""Generate tests/golden/*.npz by running the UNMODIFIED reference (through oracle/shims).

TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference):
    python oracle/gen_golden.py
Weights are the seeded "sensitised" recipe of lvae_oracle.sensitised_state_dict (regenerated
identically on any machine from key names), inputs are seeded, so fixtures hold only outputs.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_loader          # noqa: E402
import lvae_oracle as O    # noqa: E402

OUT = HERE.parent / 'tests' / 'golden'


from oracle_inputs import CASES, RD_CASES, QRES_CASES, QRES_LMB, make_input, synth_image   # noqa: E402,F401
import rd_oracle as R      # noqa: E402
import qres_oracle as Q    # noqa: E402


def main():
    ref = ref_loader.load_reference()
    torch.manual_seed(0)
    model = ref.get_model('qarv_base').eval()
    sd = O.sensitised_state_dict(O.qarv_param_shapes(), seed=0)
    model.load_state_dict(sd, strict=False)
    model.compress_mode()
    OUT.mkdir(parents=True, exist_ok=True)
    for name, (kind, nB, H, W, lmbs, seed) in CASES.items():
        im = make_input(kind, nB, H, W, seed)
        lmb = torch.tensor(lmbs)
        with torch.no_grad():
            stats = model(im, lmb=lmb, return_rec=True)
            x_hat, lat = model.forward_end2end(im, lmb, get_latent=True)
            blobs, dec = [], []
            for b in range(nB):
                s = model.compress(im[b:b + 1], lmb=float(lmbs[b]))
                blobs.append(np.frombuffer(s, dtype=np.uint8))
                dec.append(model.decompress(s))
        rec = dict(
            lmb=np.array(lmbs, dtype=np.float32),
            loss=np.float32(stats['loss'].item()), bppix=np.float64(stats['bppix']),
            mse=np.float64(stats['mse']), psnr=np.float64(stats['psnr']),
            im_hat=stats['im_hat'].numpy(),
            dec_im_hat=torch.cat(dec, 0).numpy(),
            kl_per_image=np.stack([st['kl'].sum(dim=(1, 2, 3)).numpy() for st in lat]),  # [L, B]
        )
        for li, (st, blk) in enumerate(zip(lat, [b for b in model.dec_blocks if getattr(b, 'is_latent_block', False)])):
            z = st['z']
            # recover pm from z and qm is not exposed; store symbols via the block arithmetic
            rec[f'z{li}'] = z.numpy()
        # symbols / indexes from the reference's own compress path pieces
        with torch.no_grad():
            emb = model._get_lmb_embedding(lmb, n=nB)
            x = model.preprocess_input(im)
            _, feats = model.encoder(x, emb)
            feature = model.get_bias(bhw_repeat=(nB, H // 64, W // 64))
            li = 0
            for blk in model.dec_blocks:
                if getattr(blk, 'is_latent_block', False):
                    f2, pm, pv = blk.transform_prior(feature, emb)
                    qm = blk.transform_posterior(f2, feats[blk.enc_key], emb)
                    rec[f'sym{li}'] = blk.discrete_gaussian.quantize(qm, 'symbols', pm).numpy().astype(np.int16)
                    rec[f'idx{li}'] = blk.discrete_gaussian.build_indexes(pv).numpy().astype(np.uint8)
                    feature, _ = blk(feature, emb, enc_feature=feats[blk.enc_key])
                    li += 1
                elif getattr(blk, 'requires_embedding', False):
                    feature = blk(feature, emb)
                else:
                    feature = blk(feature)
        for b, blob in enumerate(blobs):
            rec[f'bytes{b}'] = blob
        np.savez_compressed(OUT / f'{name}.npz', **rec)
        print(name, 'loss', float(rec['loss']), 'bppix', float(rec['bppix']), 'psnr', float(rec['psnr']),
              'bytes', [len(b) for b in blobs])
    # CDF tables + entropy KATs straight from the reference class
    dg = [b for b in model.dec_blocks if getattr(b, 'is_latent_block', False)][0].discrete_gaussian
    qm = torch.tensor([0.3, 0.5, 1.5, 2.5, -0.5, -1.5, 3.7, -7.2, 0.49999997, 12.0, 0.0, 40.0])
    pm = torch.tensor([0, 0, 0, 0, 0, 0, 0.25, 0.4, 0, 0.1, 0, 0.0])
    pv = torch.tensor([1.0, 1.0, 1.0, 0.2, 0.1003, 0.5, 2.0, 0.11, 0.3, 0.15, 20.0, 25.0])
    z, P = dg(qm, scales=pv, means=pm)
    ent = ref_loader  # noqa
    import lvae.models.entropy_coding as ec
    x_t = torch.tensor([0.2, 3.0, 6.0, -9.0, 0.0])
    s_t = torch.tensor([1.0, 1.0, 1.0, 1.0, 0.1003])
    np.savez_compressed(
        OUT / 'entropy_kat.npz',
        qm=qm.numpy(), pm=pm.numpy(), pv=pv.numpy(), z=z.numpy(), P=P.numpy(),
        kl=(-torch.log(P)).numpy(), sym=dg.quantize(qm, 'symbols', pm).numpy(),
        idx=dg.build_indexes(pv).numpy(), scale_table=dg.scale_table.numpy(),
        cdf=dg._quantized_cdf.numpy(), cdf_length=dg._cdf_length.numpy(), offset=dg._offset.numpy(),
        train_x=x_t.numpy(), train_scale=s_t.numpy(),
        train_logp=ec.gaussian_log_prob_mass(torch.zeros(5), s_t, x_t).numpy())
    print('entropy_kat written')


def main_rd():
    """rd_model_base fixtures: the unmodified reference with torch.manual_seed(noise_seed) before the forward, so
    its randn_like draws are reproducible as rd_oracle.draw_noise(shapes, noise_seed)."""
    ref = ref_loader.load_reference()
    torch.manual_seed(0)
    model = ref.get_model('rd_model_base').eval()
    sd = O.sensitised_state_dict(R.rd_param_shapes(), seed=0, wide_heads=False)
    model.load_state_dict(sd, strict=True)
    for name, (kind, nB, H, W, lmbs, seed, nseed) in RD_CASES.items():
        im = make_input(kind, nB, H, W, seed)
        lmb = torch.tensor(lmbs)
        with torch.no_grad():
            torch.manual_seed(nseed)
            stats = model(im, lmb=lmb, return_rec=True)
            torch.manual_seed(nseed)
            x_hat, lat = model.forward_end2end(im, lmb, get_latents=True)
        rec = dict(lmb=np.array(lmbs, dtype=np.float32), loss=np.float32(stats['loss'].item()),
                   bppix=np.float64(stats['bppix']), mse=np.float64(stats['mse']), psnr=np.float64(stats['psnr']),
                   im_hat=stats['im_hat'].numpy(),
                   kl_per_image=np.stack([st['kl'].sum(dim=(1, 2, 3)).numpy() for st in lat]),
                   z0=lat[0]['z'].numpy(), z14=lat[14]['z'].numpy())
        np.savez_compressed(OUT / f'{name}.npz', **rec)
        print(name, 'loss', float(rec['loss']), 'bppix', float(rec['bppix']), 'psnr', float(rec['psnr']))


def main_qres():
    """qres34m fixtures: the unmodified reference in eval mode (forward, latents, compress / decompress) and in train
    mode with torch.manual_seed(noise_seed) before the forward (its uniform_ draws are then reproducible)."""
    ref = ref_loader.load_reference()
    torch.manual_seed(0)
    model = ref.get_model('qres34m', lmb=QRES_LMB).eval()
    sd = O.sensitised_state_dict(Q.qres_param_shapes(), seed=0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all('discrete_gaussian' in k for k in missing)
    model.compress_mode()
    for name, (kind, nB, H, W, seed, nseed) in QRES_CASES.items():
        im = make_input(kind, nB, H, W, seed)
        with torch.no_grad():
            model.eval()
            stats = model(im, return_rec=True)
            lat = model.forward_get_latents(im)
            obj = model.compress(im)
            dec = model.decompress(obj)
            # symbols / indexes from the reference's own block pieces
            feats = model.encoder(model.preprocess_input(im))
            feature = model.decoder.bias.expand(feats[min(feats.keys())].shape)
            syms, idxs = [], []
            for blk in model.decoder.dec_blocks:
                if hasattr(blk, 'forward_train'):
                    f_enc = feats[int(feature.shape[2])]
                    f2, pm, plogv = blk.transform_prior(feature)
                    qm = blk.posterior(torch.cat([f2, f_enc], dim=1))
                    syms.append(blk.discrete_gaussian.quantize(qm, 'symbols', pm).numpy().astype(np.int16))
                    idxs.append(blk.discrete_gaussian.build_indexes(torch.exp(plogv)).numpy().astype(np.uint8))
                    feature, _ = blk.forward_train(feature, f_enc)
                else:
                    feature = blk(feature)
            model.train()
            torch.manual_seed(nseed)
            tstats = model(im)
            torch.manual_seed(nseed)
            tlat = model.forward_get_latents(im)
            model.eval()
        rec = dict(loss=np.float32(stats['loss'].item()), kl=np.float64(stats['kl']), mse=np.float64(stats['mse']),
                   bppix=np.float64(stats['bppix']), psnr=np.float64(stats['psnr']), im_hat=stats['im_hat'].numpy(),
                   dec_im_hat=dec.numpy(), shape=np.array(obj[-1]),
                   kl_per_image=np.stack([st['kl'].sum(dim=(1, 2, 3)).numpy() for st in lat]),
                   train_loss=np.float32(tstats['loss'].item()), train_bppix=np.float64(tstats['bppix']),
                   train_psnr=np.float64(tstats['psnr']),
                   train_kl_per_image=np.stack([st['kl'].sum(dim=(1, 2, 3)).numpy() for st in tlat]))
        for li in range(len(lat)):
            rec[f'z{li}'] = lat[li]['z'].numpy()
            rec[f'sym{li}'], rec[f'idx{li}'] = syms[li], idxs[li]
            for b in range(nB):
                rec[f'bytes{li}_{b}'] = np.frombuffer(obj[li][b], dtype=np.uint8)
        np.savez_compressed(OUT / f'{name}.npz', **rec)
        print(name, 'loss', float(rec['loss']), 'bppix', float(rec['bppix']), 'psnr', float(rec['psnr']),
              'train loss', float(rec['train_loss']), 'bytes', sum(len(s) for l in obj[:-1] for s in l))
    dg = [b for b in model.decoder.dec_blocks if hasattr(b, 'forward_train')][0].discrete_gaussian
    np.savez_compressed(OUT / 'qres_tables.npz', scale_table=dg.scale_table.numpy(), cdf=dg._quantized_cdf.numpy(),
                        cdf_length=dg._cdf_length.numpy(), offset=dg._offset.numpy())


def main_lossless():
    """qres34m_lossless fixtures (GaussianNLLOutputNet, qresvae/model.py:16-94, zoo.py:63-114): the unmodified reference's
    eval forward (loss = kl + nll), train forward with torch.manual_seed(noise_seed), compress (latent layers + the image's
    own residual stream) and decompress on 8-bit images."""
    from oracle_inputs import LOSSLESS_CASES, make_input_8bit
    ref = ref_loader.load_reference()
    torch.manual_seed(0)
    model = ref.get_model('qres34m_lossless').eval()
    sd = O.sensitised_state_dict(Q.qres_param_shapes(Q.qres34m_lossless_arch()), seed=0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all('discrete_gaussian' in k for k in missing)
    model.compress_mode()
    for name, (kind, nB, H, W, seed, nseed) in LOSSLESS_CASES.items():
        im = make_input_8bit(kind, nB, H, W, seed)
        with torch.no_grad():
            stats = model(im, return_rec=True)
            lat = model.forward_get_latents(im)
            obj = model.compress(im)
            dec = model.decompress(obj)
            model.train()
            torch.manual_seed(nseed)
            tstats = model(im)
            model.eval()
        rec = dict(loss=np.float32(stats['loss'].item()), kl=np.float64(stats['kl']), nll=np.float64(stats['nll']),
                   bppix=np.float64(stats['bppix']), psnr=np.float64(stats['psnr']), im_hat=stats['im_hat'].numpy(),
                   dec_im_hat=dec.numpy(), shape=np.array(obj[-2]),
                   kl_per_image=np.stack([st['kl'].sum(dim=(1, 2, 3)).numpy() for st in lat]),
                   train_loss=np.float32(tstats['loss'].item()), train_nll=np.float64(tstats['nll']))
        for li in range(len(obj) - 2):
            for b in range(nB):
                rec[f'bytes{li}_{b}'] = np.frombuffer(obj[li][b], dtype=np.uint8)
        for b in range(nB):
            rec[f'final_bytes_{b}'] = np.frombuffer(obj[-1][b], dtype=np.uint8)
        np.savez_compressed(OUT / f'{name}.npz', **rec)
        print(name, 'loss', float(rec['loss']), 'nll', float(rec['nll']), 'bppix', float(rec['bppix']),
              'lossless', bool(torch.equal((dec * 255).round(), (im * 255).round())), 'final bytes', [len(s) for s in obj[-1]])
    dg = model.out_net.discrete_gaussian
    np.savez_compressed(OUT / 'qresll_tables.npz', scale_table=dg.scale_table.numpy(), cdf=dg._quantized_cdf.numpy(),
                        cdf_length=dg._cdf_length.numpy(), offset=dg._offset.numpy())


from gen_golden_cases import GRAD_CASES, grad_probe   # noqa: E402


def main_grads():
    """Training-step gradients of the UNMODIFIED reference (train mode, torch.manual_seed(noise_seed) right before the
    forward so that its uniform_ draws are reproducible; sensitised weights): loss, and per parameter tensor the gradient's
    L2 norm and its inner product with grad_probe().  tests/test_oracle_pinned.py checks the oracles' autograd against it."""
    ref = ref_loader.load_reference()
    for fam, (nB, H, W, lmbs, seed, nseed) in GRAD_CASES.items():
        torch.manual_seed(0)
        if fam == 'qarv':
            model = ref.get_model('qarv_base')
            sd = O.sensitised_state_dict(O.qarv_param_shapes(), seed=0)
        else:
            model = ref.get_model('qres34m', lmb=QRES_LMB)
            sd = O.sensitised_state_dict(Q.qres_param_shapes(), seed=0)
        model.load_state_dict(sd, strict=False)
        model.train()
        im = make_input('rand', nB, H, W, seed)
        torch.manual_seed(nseed)
        stats = model(im, lmb=torch.tensor(lmbs)) if fam == 'qarv' else model(im)
        stats['loss'].backward()
        names = [k for k, _ in model.named_parameters()]
        norms = np.array([float(p.grad.double().norm()) for _, p in model.named_parameters()])
        dots = np.array([float((p.grad.double() * grad_probe(k, p.shape).double()).sum()) for k, p in model.named_parameters()])
        np.savez_compressed(OUT / f'{fam}_train_grads.npz', loss=np.float64(stats['loss'].item()), names=np.array(names),
                            grad_norm=norms, grad_dot=dots)
        print(fam, 'train loss', stats['loss'].item(), 'grad norms', norms.min(), norms.max())


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which in ('qarv', 'all'):
        main()
    if which in ('rd', 'all'):
        main_rd()
    if which in ('qres', 'all'):
        main_qres()
    if which in ('grads', 'all'):
        main_grads()
    if which in ('lossless', 'all'):
        main_lossless()
