"""CPU oracle for the rd (continuous-posterior R(D)) model of the reference.

TEST INFRASTRUCTURE ONLY (same rules as lvae_oracle.py).  Restates as plain functions over a state dict:
  * `rd_model_base` architecture table                 /root/reference/lvae/models/rd/zoo.py:10-77
  * `ConvNeXtAdaLNPatchDown`, `linear_sqrt`, `gaussian_kl`  lvae/models/rd/model.py:16-49
  * `LatentVariableBlock.transform_prior/_posterior/forward`  rd/model.py:162-227
  * `FeatureExtractor` (features keyed by spatial height, last writer wins)  rd/model.py:230-244
  * `VariableRateLossyVAE.forward_end2end / forward`   rd/model.py:377-445
The reference draws `torch.randn_like(qm)` per latent layer, in layer order, also in eval mode (rd/model.py:213);
here the noise is an explicit list so that the GPU path can be fed the same values.  Pinned bit-exactly against the
unmodified reference by tests/test_oracle_pinned.py (live, build container) and tests/golden/rd_*.npz (everywhere).
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

import lvae_oracle as O


def rd_base_arch():
    e = [256, 512, 640, 768, 768]
    d = [768, 768, 640, 512, 256]
    enc = [('down', 3, e[0], 4)] + [('blk', e[0], 7, 2)] * 6 + [('blkdown', e[0], e[1], 7, 2)]
    enc += [('blk', e[1], 7, 2)] * 6 + [('blkdown', e[1], e[2], 7, 2)]
    enc += [('blk', e[2], 7, 2)] * 6 + [('blkdown', e[2], e[3], 7, 2)]
    enc += [('blk', e[3], 7, 2)] * 4 + [('blkdown', e[3], e[3], 7, 2)]
    enc += [('blk', e[3], 7, 2)] * 4
    dec = [('lat', d[0], 32, e[4], 7, 2)] + [('up', d[0], d[1], 2)]
    dec += [('lat', d[1], 32, e[3], 7, 2)] * 2 + [('up', d[1], d[2], 2)]
    dec += [('lat', d[2], 32, e[2], 7, 2)] * 3 + [('up', d[2], d[3], 2)]
    dec += [('lat', d[3], 32, e[1], 7, 2)] * 4 + [('up', d[3], d[4], 2)]
    dec += [('lat', d[4], 32, e[0], 7, 2)] * 5 + [('up', d[4], 3, 4)]
    return dict(enc=enc, dec=dec, im_shift=-0.4546259594901961, im_scale=3.67572653978347,
                max_stride=64, sin_period=64, embed_dim=256, lmb_range=(4.0, 2048.0))


def rd_param_shapes(arch=None):
    arch = arch or rd_base_arch()
    E = arch['embed_dim']
    out = []

    def blk(prefix, C, k, ratio):
        hid = int(ratio * C)
        out.extend([
            (prefix + 'gamma', (1, C, 1, 1)),
            (prefix + 'conv_dw.weight', (C, 1, k, k)), (prefix + 'conv_dw.bias', (C,)),
            (prefix + 'embedding_layer.1.weight', (2 * C, E)), (prefix + 'embedding_layer.1.bias', (2 * C,)),
            (prefix + 'mlp.fc1.weight', (hid, C)), (prefix + 'mlp.fc1.bias', (hid,)),
            (prefix + 'mlp.fc2.weight', (C, hid)), (prefix + 'mlp.fc2.bias', (C,)),
        ])

    for i, ent in enumerate(arch['enc']):
        p = f'encoder.enc_blocks.{i}.'
        if ent[0] == 'down':
            out += [(p + 'weight', (ent[2], ent[1], ent[3], ent[3])), (p + 'bias', (ent[2],))]
        elif ent[0] == 'blk':
            blk(p, ent[1], ent[2], ent[3])
        elif ent[0] == 'blkdown':
            blk(p, ent[1], ent[3], ent[4])
            out += [(p + 'downsapmle.weight', (ent[2], ent[1], 2, 2)), (p + 'downsapmle.bias', (ent[2],))]   # sic
    for i, ent in enumerate(arch['dec']):
        p = f'dec_blocks.{i}.'
        if ent[0] == 'up':
            _, cin, cout, r = ent
            out += [(p + '0.weight', (cout * r * r, cin, 1, 1)), (p + '0.bias', (cout * r * r,))]
        elif ent[0] == 'lat':
            _, W, zd, We, k, ratio = ent
            blk(p + 'resnet_front.', W, k, ratio)
            blk(p + 'resnet_end.', W, k, ratio)
            blk(p + 'posterior0.', We, k, 2)
            blk(p + 'posterior1.', W, k, 2)
            blk(p + 'posterior2.', W, k, 2)
            out += [(p + 'post_merge.weight', (W, W + We, 1, 1)), (p + 'post_merge.bias', (W,)),
                    (p + 'posterior.weight', (2 * zd, W, 3, 3)), (p + 'posterior.bias', (2 * zd,)),
                    (p + 'prior.weight', (2 * zd, W, 1, 1)), (p + 'prior.bias', (2 * zd,)),
                    (p + 'z_proj.weight', (W, zd, 1, 1)), (p + 'z_proj.bias', (W,))]
    out.append(('bias', (1, arch['dec'][0][1], 1, 1)))
    for j in (0, 2):
        out += [(f'lmb_embedding.{j}.weight', (E, E)), (f'lmb_embedding.{j}.bias', (E,))]
    return out


def linear_sqrt(x, threshold=6.0):
    """rd/model.py:27-39"""
    x_abs = torch.abs(x)
    soft = torch.sign(x) * torch.pow(x_abs, 1 - 0.5 * torch.tanh(x_abs))
    soft = torch.where(x_abs == 0, input=x, other=soft)
    signed_sqrt = torch.sign(x) * torch.sqrt(x_abs + 1e-8)
    return torch.where(x_abs <= threshold, input=soft, other=signed_sqrt)


def gaussian_kl(mu1, v1, mu2, v2):
    """rd/model.py:41-49"""
    return -0.5 + v2.log() - v1.log() + 0.5 * (v1 ** 2 + (mu1 - mu2) ** 2) / (v2 ** 2)


def std_smooth(v):
    return F.softplus(v, beta=math.log(2), threshold=12)


def latent_shapes(arch, nB, H, W):
    """[(nB, zdim, h, w)] per latent layer, in layer order (the order the reference draws its noise in)."""
    s = arch['max_stride']
    h, w = H // s, W // s
    out = []
    for ent in arch['dec']:
        if ent[0] == 'lat':
            out.append((nB, ent[2], h, w))
        elif ent[0] == 'up':
            h, w = h * ent[3], w * ent[3]
    return out


def draw_noise(shapes, seed):
    """What `torch.randn_like` yields layer by layer after torch.manual_seed(seed) on the CPU generator."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) for s in shapes]


@torch.no_grad()
def rd_forward(sd, im, lmb, noise, arch=None):
    """VariableRateLossyVAE.forward (rd/model.py:399-445) with explicit posterior-sampling noise."""
    arch = arch or rd_base_arch()
    nB, imC, imH, imW = im.shape
    x = im.clone().add_(arch['im_shift']).mul_(arch['im_scale'])
    emb = O.lmb_embedding(sd, lmb, arch)
    feats = OrderedDict()
    for i, ent in enumerate(arch['enc']):
        p = f'encoder.enc_blocks.{i}.'
        if ent[0] == 'down':
            x = F.conv2d(x, sd[p + 'weight'], sd[p + 'bias'], stride=ent[3])
        elif ent[0] == 'blk':
            x = O.convnext_block(sd, p, x, emb)
        elif ent[0] == 'blkdown':
            x = O.convnext_block(sd, p, x, emb)
            x = F.conv2d(x, sd[p + 'downsapmle.weight'], sd[p + 'downsapmle.bias'], stride=2)
        feats[int(x.shape[2])] = x
    nH, nW = feats[min(feats.keys())].shape[2:4]
    feature = sd['bias'].expand(nB, -1, nH, nW)
    records, li = [], 0
    for i, ent in enumerate(arch['dec']):
        p = f'dec_blocks.{i}.'
        if ent[0] == 'up':
            feature = F.pixel_shuffle(F.conv2d(feature, sd[p + '0.weight'], sd[p + '0.bias']), ent[3])
        elif ent[0] == 'lat':
            f_enc = feats[int(feature.shape[2])]
            feature = O.convnext_block(sd, p + 'resnet_front.', feature, emb)
            pm, pv = F.conv2d(feature, sd[p + 'prior.weight'], sd[p + 'prior.bias']).chunk(2, dim=1)
            pm, pv = linear_sqrt(pm), std_smooth(pv)
            e = O.convnext_block(sd, p + 'posterior0.', f_enc, emb)
            f = O.convnext_block(sd, p + 'posterior1.', feature, emb)
            m = F.conv2d(torch.cat([f, e], dim=1), sd[p + 'post_merge.weight'], sd[p + 'post_merge.bias'])
            m = O.convnext_block(sd, p + 'posterior2.', m, emb)
            qm, qv = F.conv2d(m, sd[p + 'posterior.weight'], sd[p + 'posterior.bias'], padding=1).chunk(2, dim=1)
            qm, qv = linear_sqrt(qm), std_smooth(qv)
            kl = gaussian_kl(qm, qv, pm, pv)
            z = qm + qv * noise[li]
            records.append(dict(kl=kl, z=z, qm=qm, qv=qv, pm=pm, pv=pv))
            li += 1
            feature = feature + F.conv2d(z, sd[p + 'z_proj.weight'], sd[p + 'z_proj.bias'])
            feature = O.convnext_block(sd, p + 'resnet_end.', feature, emb)
    x_hat = feature
    kls = [r['kl'].sum(dim=(1, 2, 3)) for r in records]
    ndims = float(imC * imH * imW)
    kl = sum(kls) / ndims
    x_target = im.clone().add_(-0.5).mul_(2.0)
    distortion = F.mse_loss(x_hat, x_target, reduction='none').mean(dim=(1, 2, 3))
    loss = (kl + lmb * distortion).mean(0)
    im_hat = x_hat.clone().clamp_(min=-1.0, max=1.0).mul_(0.5).add_(0.5)
    im_mse = F.mse_loss(im_hat, im, reduction='mean')
    return dict(loss=loss, bppix=kl.mean(0).item() * O.LOG2_E * imC, mse=distortion.mean(0).item(),
                psnr=-10 * math.log10(im_mse.item()), kl_per_image=kl, mse_per_image=distortion,
                x_hat=x_hat, im_hat=im_hat, records=records)
