"""CPU oracle for the fixed-rate qres34m model of the reference.

TEST INFRASTRUCTURE ONLY (same rules as lvae_oracle.py).  Restates as plain functions over a state dict:
  * `qres34m` architecture table                         /root/reference/lvae/models/qresvae/zoo.py:9-58
  * `MyConvNeXtBlock` / `MyConvNeXtPatchDown` (timm ConvNeXtBlock, affine LayerNorm)  qresvae/model.py:163-192
  * `VDBlock` (GELU before every conv)                   qresvae/model.py:120-149
  * `QLatentBlockX.transform_prior / forward_train / compress / decompress / update`  qresvae/model.py:246-360
  * `BottomUpEncoder` (features keyed by height, last writer wins), `TopDownDecoder`  qresvae/model.py:195-207,363-454
  * `HierarchicalVAE.forward / compress / decompress`, `MSEOutputNet.forward_loss`    qresvae/model.py:95-113,517-687
  * CompressAI's stock `GaussianConditional` (erfc CDF, scale bound 0.11, likelihood bound 1e-9; table 0.1 ... 20)
The training branch draws `uniform_(-0.5, 0.5)` per latent layer in layer order (qresvae/model.py:274); here the noise
is an explicit list.  Pinned bit-exactly against the unmodified reference by tests/test_oracle_pinned.py (live, build
container) and tests/golden/qres_*.npz (everywhere).
"""
import math

import torch
import torch.nn.functional as F

import lvae_oracle as O


def qres34m_arch():
    ch = 96
    w = [ch * 2, ch * 4, ch * 4, ch * 4, ch * 4]
    ks = [7, 7, 5, 3, 1]
    enc_nums, dec_nums, zd = [6, 6, 6, 4, 2], [1, 2, 3, 3, 3], [16, 14, 12, 10, 8]
    enc = [('down', 3, w[0], 4)]
    for s in range(5):
        enc += [('blk', w[s], ks[s], 2)] * enc_nums[s]
        if s < 4:
            enc.append(('blkdown', w[s], w[s + 1], 7, 2))
    dec = []
    for s in range(5):
        width = w[4 - s]
        dec += [('lat', width, zd[s], width, ks[4 - s])] * dec_nums[s]
        dec.append(('up', width, w[3 - s], 2) if s < 4 else ('up', width, 3, 4))
    return dict(enc=enc, dec=dec, im_shift=-0.4546259594901961, im_scale=3.67572653978347, max_stride=64)


def qres34m_lossless_arch():
    """`qres34m_lossless` (qresvae/zoo.py:63-114): qres34m without the final patch up-sampler; the decoder's H/4 feature
    feeds GaussianNLLOutputNet's two patch_upsample(192, 3, rate=4) heads (qresvae/model.py:16-94)."""
    a = qres34m_arch()
    assert a['dec'][-1] == ('up', 192, 3, 4)
    a['dec'] = a['dec'][:-1]
    a['out'] = 'nll'
    return a


def lossless_scale_table():
    """GaussianNLLOutputNet.update (qresvae/model.py:58-66): 128 scales from the 0.11 scale bound to 20."""
    t = torch.exp(torch.linspace(math.log(0.11), math.log(20), steps=128))
    return torch.Tensor(tuple(float(s) for s in t))


BIN = 1 / 127.5


def qres_param_shapes(arch=None):
    arch = arch or qres34m_arch()
    out = []

    def blk(prefix, C, k, ratio):
        hid = int(ratio * C)
        out.extend([
            (prefix + 'gamma', (C,)),
            (prefix + 'conv_dw.weight', (C, 1, k, k)), (prefix + 'conv_dw.bias', (C,)),
            (prefix + 'norm.weight', (C,)), (prefix + 'norm.bias', (C,)),
            (prefix + 'mlp.fc1.weight', (hid, C)), (prefix + 'mlp.fc1.bias', (hid,)),
            (prefix + 'mlp.fc2.weight', (C, hid)), (prefix + 'mlp.fc2.bias', (C,)),
        ])

    def vd(prefix, cin, hid, cout, k3):
        k = 3 if k3 else 1
        out.extend([(prefix + 'c1.weight', (hid, cin, 1, 1)), (prefix + 'c1.bias', (hid,)),
                    (prefix + 'c2.weight', (hid, hid, k, k)), (prefix + 'c2.bias', (hid,)),
                    (prefix + 'c3.weight', (hid, hid, k, k)), (prefix + 'c3.bias', (hid,)),
                    (prefix + 'c4.weight', (cout, hid, 1, 1)), (prefix + 'c4.bias', (cout,))])

    for i, ent in enumerate(arch['enc']):
        p = f'encoder.enc_blocks.{i}.'
        if ent[0] == 'down':
            out += [(p + 'weight', (ent[2], ent[1], ent[3], ent[3])), (p + 'bias', (ent[2],))]
        elif ent[0] == 'blk':
            blk(p, ent[1], ent[2], ent[3])
        elif ent[0] == 'blkdown':
            blk(p, ent[1], ent[3], ent[4])
            out += [(p + 'downsapmle.weight', (ent[2], ent[1], 2, 2)), (p + 'downsapmle.bias', (ent[2],))]   # sic
    out.append(('decoder.bias', (1, arch['dec'][0][1], 1, 1)))
    for i, ent in enumerate(arch['dec']):
        p = f'decoder.dec_blocks.{i}.'
        if ent[0] == 'up':
            _, cin, cout, r = ent
            out += [(p + '0.weight', (cout * r * r, cin, 1, 1)), (p + '0.bias', (cout * r * r,))]
        elif ent[0] == 'lat':
            _, W, zd, We, k = ent
            hid = int(max(W, We) * 0.25)
            k3 = k >= 3
            blk(p + 'resnet_front.', W, k, 2)
            blk(p + 'resnet_end.', W, k, 2)
            vd(p + 'posterior.', W + We, hid, zd, k3)
            vd(p + 'prior.', W, hid, 2 * zd, k3)
            kz = 3 if k3 else 1
            out += [(p + 'z_proj.0.weight', (hid // 2, zd, kz, kz)), (p + 'z_proj.0.bias', (hid // 2,)),
                    (p + 'z_proj.2.weight', (W, hid // 2, 1, 1)), (p + 'z_proj.2.bias', (W,))]
    if arch.get('out') == 'nll':
        W = arch['dec'][-1][1]
        for head in ('conv_mean', 'conv_scale'):
            out += [(f'out_net.{head}.0.weight', (3 * 16, W, 1, 1)), (f'out_net.{head}.0.bias', (3 * 16,))]
    return out


def qres_scale_table():
    """QLatentBlockX.update (qresvae/model.py:317-325) -> update_scale_table (values pass through python floats)."""
    t = torch.exp(torch.linspace(math.log(0.1), math.log(20), steps=64))
    return torch.Tensor(tuple(float(s) for s in t))


def convnext_block(sd, p, x):
    """MyConvNeXtBlock.forward (qresvae/model.py:168-182); x NCHW."""
    w = sd[p + 'conv_dw.weight']
    C, k = w.shape[0], w.shape[-1]
    y = F.conv2d(x, w, sd[p + 'conv_dw.bias'], padding=(k - 1) // 2, groups=C)
    y = y.permute(0, 2, 3, 1).contiguous()
    y = F.layer_norm(y, (C,), sd[p + 'norm.weight'], sd[p + 'norm.bias'], eps=1e-6)
    y = F.linear(y, sd[p + 'mlp.fc1.weight'], sd[p + 'mlp.fc1.bias'])
    y = F.gelu(y)
    y = F.linear(y, sd[p + 'mlp.fc2.weight'], sd[p + 'mlp.fc2.bias'])
    y = y.permute(0, 3, 1, 2).contiguous()
    y = y.mul(sd[p + 'gamma'].reshape(1, -1, 1, 1))
    return y + x


def vdblock(sd, p, x):
    """VDBlock.forward, residual=False (qresvae/model.py:143-149)."""
    pad = (sd[p + 'c2.weight'].shape[-1] - 1) // 2
    h = F.conv2d(F.gelu(x), sd[p + 'c1.weight'], sd[p + 'c1.bias'])
    h = F.conv2d(F.gelu(h), sd[p + 'c2.weight'], sd[p + 'c2.bias'], padding=pad)
    h = F.conv2d(F.gelu(h), sd[p + 'c3.weight'], sd[p + 'c3.bias'], padding=pad)
    return F.conv2d(F.gelu(h), sd[p + 'c4.weight'], sd[p + 'c4.bias'])


def z_proj(sd, p, z):
    pad = (sd[p + 'z_proj.0.weight'].shape[-1] - 1) // 2
    t = F.gelu(F.conv2d(z, sd[p + 'z_proj.0.weight'], sd[p + 'z_proj.0.bias'], padding=pad))
    return F.conv2d(t, sd[p + 'z_proj.2.weight'], sd[p + 'z_proj.2.bias'])


def encoder(sd, arch, x):
    feats = {}
    for i, ent in enumerate(arch['enc']):
        p = f'encoder.enc_blocks.{i}.'
        if ent[0] == 'down':
            x = F.conv2d(x, sd[p + 'weight'], sd[p + 'bias'], stride=ent[3])
        elif ent[0] == 'blk':
            x = convnext_block(sd, p, x)
        elif ent[0] == 'blkdown':
            x = convnext_block(sd, p, x)
            x = F.conv2d(x, sd[p + 'downsapmle.weight'], sd[p + 'downsapmle.bias'], stride=2)
        feats[int(x.shape[2])] = x
    return feats


def prior_transform(sd, p, feature):
    pm, plogv = vdblock(sd, p + 'prior.', feature).chunk(2, dim=1)
    plogv = F.softplus(plogv + 2.3) - 2.3
    return pm, plogv


@torch.no_grad()
def qres_forward(sd, im, lmb, arch=None, mode='eval', noise=None):
    """HierarchicalVAE.forward (qresvae/model.py:517-569).  mode 'eval' | 'train' (explicit noise list)."""
    arch = arch or qres34m_arch()
    nB, imC, imH, imW = im.shape
    x = (im + arch['im_shift']) * arch['im_scale']
    x_target = (im - 0.5) * 2.0
    feats = encoder(sd, arch, x)
    min_res = min(feats.keys())
    feature = sd['decoder.bias'].expand(feats[min_res].shape)
    table = qres_scale_table()
    records, li = [], 0
    for i, ent in enumerate(arch['dec']):
        p = f'decoder.dec_blocks.{i}.'
        if ent[0] == 'up':
            feature = F.pixel_shuffle(F.conv2d(feature, sd[p + '0.weight'], sd[p + '0.bias']), ent[3])
        elif ent[0] == 'lat':
            f_enc = feats[int(feature.shape[2])]
            feature = convnext_block(sd, p + 'resnet_front.', feature)
            pm, plogv = prior_transform(sd, p, feature)
            pv = torch.exp(plogv)
            qm = vdblock(sd, p + 'posterior.', torch.cat([feature, f_enc], dim=1))
            if mode == 'train':
                z = qm + noise[li]
                kl = -1.0 * O.gaussian_log_prob_mass(pm, pv, x=z, bin_size=1.0, prob_clamp=1e-6)
            else:
                z, probs = O.eval_quantize_likelihood(qm, pm, pv, cdf=O.std_normal_cdf_erfc)
                kl = -1.0 * torch.log(probs)
            records.append(dict(kl=kl, z=z, qm=qm, pm=pm, pv=pv, sym=O.symbols(qm, pm),
                                idx=O.build_indexes(pv, scale_table=table)))
            li += 1
            feature = feature + z_proj(sd, p, z)
            feature = convnext_block(sd, p + 'resnet_end.', feature)
    if arch.get('out') == 'nll':
        # GaussianNLLOutputNet.forward_loss (qresvae/model.py:24-40)
        p_mean = F.pixel_shuffle(F.conv2d(feature, sd['out_net.conv_mean.0.weight'], sd['out_net.conv_mean.0.bias']), 4)
        p_ls = F.pixel_shuffle(F.conv2d(feature, sd['out_net.conv_scale.0.weight'], sd['out_net.conv_scale.0.bias']), 4)
        p_ls = F.softplus(p_ls + 16) - 16
        log_prob = O.gaussian_log_prob_mass(p_mean, torch.exp(p_ls), x=x_target, bin_size=BIN, prob_clamp=1e-6)
        out_loss = -log_prob.mean(dim=(1, 2, 3))
        x_hat = p_mean
    else:
        x_hat = feature
        mse = F.mse_loss(x_hat, x_target, reduction='none').mean(dim=(1, 2, 3))
        out_loss = mse * float(lmb)
    kls = [r['kl'].sum(dim=(1, 2, 3)) for r in records]
    ndims = imC * imH * imW
    kl = sum(kls) / ndims
    loss = (kl + out_loss).mean(0)
    nats_per_dim = kl.mean(0).item()
    im_hat = x_hat.clone().clamp_(min=-1.0, max=1.0).mul_(0.5).add_(0.5)
    im_mse = F.mse_loss(im_hat, im, reduction='mean')
    return dict(loss=loss, kl=nats_per_dim, mse=out_loss.mean(0).item(), bppix=nats_per_dim * O.LOG2_E * imC,
                psnr=-10 * math.log10(im_mse.item()), kl_per_image=kl, x_hat=x_hat, im_hat=im_hat, records=records,
                feature=feature, out_loss=out_loss)


def lossless_tables():
    return O.build_cdf_tables(lossless_scale_table(), cdf=O.std_normal_cdf_erfc)


def _prepare_codec(sd, feature, x=None):
    """GaussianNLLOutputNet._preapre_codec (qresvae/model.py:68-79)."""
    pm = F.pixel_shuffle(F.conv2d(feature, sd['out_net.conv_mean.0.weight'], sd['out_net.conv_mean.0.bias']), 4)
    pm = torch.round(pm * 127.5 + 127.5) / 127.5 - 1
    plogv = F.pixel_shuffle(F.conv2d(feature, sd['out_net.conv_scale.0.weight'], sd['out_net.conv_scale.0.bias']), 4)
    pm = pm / BIN
    plogv = plogv - math.log(BIN)
    if x is not None:
        x = x / BIN
    return pm, plogv, x


@torch.no_grad()
def lossless_compress(sd, im, arch=None, tables=None, out_tables=None):
    """HierarchicalVAE.compress with a GaussianNLLOutputNet (qresvae/model.py:649-668,81-86): latent strings, feature
    shape, then the image's own residual symbols round(x / bin - pm) coded against the out-net's 128-scale tables."""
    arch = arch or qres34m_lossless_arch()
    out_tables = out_tables or lossless_tables()
    fw = qres_forward(sd, im, 0.0, arch)
    obj = qres_compress(sd, im, arch, tables)
    pm, plogv, x = _prepare_codec(sd, fw['feature'], (im - 0.5) * 2.0)
    idx = O.build_indexes(torch.exp(plogv), scale_table=lossless_scale_table())
    sym = torch.round(x - pm).to(torch.int32)
    obj.append([O.rans_encode(sym[b].reshape(-1).tolist(), idx[b].reshape(-1).tolist(), *out_tables) for b in range(im.shape[0])])
    return obj, dict(sym=sym, idx=idx, pm=pm)


@torch.no_grad()
def lossless_decompress(sd, obj, arch=None, tables=None, out_tables=None):
    """HierarchicalVAE.decompress, lossless branch (qresvae/model.py:670-687,88-94) -> image in [0, 1]."""
    arch = arch or qres34m_lossless_arch()
    out_tables = out_tables or lossless_tables()
    feature = qres_decompress(sd, obj[:-1], arch, tables, return_feature=True)
    pm, plogv, _ = _prepare_codec(sd, feature)
    idx = O.build_indexes(torch.exp(plogv), scale_table=lossless_scale_table())
    nB = pm.shape[0]
    vals = [O.rans_decode(obj[-1][b], idx[b].reshape(-1).tolist(), *out_tables) for b in range(nB)]
    x_hat = (torch.tensor(vals, dtype=torch.int32).reshape(pm.shape).type_as(pm) + pm) * BIN
    return x_hat.clone().clamp_(min=-1.0, max=1.0).mul_(0.5).add_(0.5)


def qres_tables():
    return O.build_cdf_tables(qres_scale_table(), cdf=O.std_normal_cdf_erfc)


@torch.no_grad()
def qres_compress(sd, im, arch=None, tables=None):
    """HierarchicalVAE.compress (qresvae/model.py:649-668): [strings per layer (list over the batch)] + [shape]."""
    arch = arch or qres34m_arch()
    tables = tables or qres_tables()
    rec = qres_forward(sd, im, 0.0, arch)['records']
    nB = im.shape[0]
    out = []
    for r in rec:
        out.append([O.rans_encode(r['sym'][b].reshape(-1).tolist(), r['idx'][b].reshape(-1).tolist(), *tables)
                    for b in range(nB)])
    out.append((nB, arch['dec'][0][1], im.shape[2] // 64, im.shape[3] // 64))
    return out


@torch.no_grad()
def qres_decompress(sd, obj, arch=None, tables=None, return_feature=False):
    """HierarchicalVAE.decompress (qresvae/model.py:670-687)."""
    arch = arch or qres34m_arch()
    tables = tables or qres_tables()
    table = qres_scale_table()
    feature = sd['decoder.bias'].expand(obj[-1])
    nB = obj[-1][0]
    si = 0
    for i, ent in enumerate(arch['dec']):
        p = f'decoder.dec_blocks.{i}.'
        if ent[0] == 'up':
            feature = F.pixel_shuffle(F.conv2d(feature, sd[p + '0.weight'], sd[p + '0.bias']), ent[3])
        elif ent[0] == 'lat':
            feature = convnext_block(sd, p + 'resnet_front.', feature)
            pm, plogv = prior_transform(sd, p, feature)
            idx = O.build_indexes(torch.exp(plogv), scale_table=table)
            vals = [O.rans_decode(obj[si][b], idx[b].reshape(-1).tolist(), *tables) for b in range(nB)]
            si += 1
            z = torch.tensor(vals, dtype=torch.int32).reshape(pm.shape).type_as(pm) + pm
            feature = feature + z_proj(sd, p, z)
            feature = convnext_block(sd, p + 'resnet_end.', feature)
    if return_feature:
        return feature
    return feature.clone().clamp_(min=-1.0, max=1.0).mul_(0.5).add_(0.5)
