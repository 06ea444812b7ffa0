"""CPU oracle for the qarv hierarchical-VAE rate-distortion path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package (`lossy-vae_b200/`).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it, and only as the checker / the reported CPU baseline.

It restates, as plain functions over a state dict (torch CPU fp32, same ATen op order as the
reference so results are bit-identical to it on the same machine), the algorithm of:

  * `ConvNeXtBlockAdaLN.forward`            /root/reference/lvae/models/common.py:142-161
  * `patch_downsample` / `patch_upsample`   common.py:29-38
  * `sinusoidal_embedding`                  common.py:101-107
  * `VRLVBlockBase.transform_prior/_posterior/forward`  lvae/models/qarv/model.py:44-121
  * `VariableRateLossyVAE.forward_end2end/forward/compress/decompress`  qarv/model.py:294-363,516-557
  * `qarv_base` architecture table          lvae/models/qarv/zoo.py:35-88
  * the CompressAI (un-vendored, un-pinned) `GaussianConditional` arithmetic reached through
    `DiscretizedGaussian` (lvae/models/entropy_coding.py:52-82): quantize / _likelihood /
    likelihood lower bound / build_indexes / update() CDF tables / rANS coding
    (CompressAI entropy_models.py, cpp_exts/ops/ops.cpp, cpp_exts/rans/rans_interface.cpp).

Pinning: `tests/test_oracle_pinned.py` checks this file bit-exactly against the unmodified
reference imported through `oracle/shims` (build container only) and against the committed
fixtures in `tests/golden/` (everywhere), plus SURVEY Appendix-A known answers. The rANS byte
stream and CDF tables are "parity unpinned" w.r.t. a real CompressAI build (absent offline).
"""
import math
import struct
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

LOG2_E = math.log2(math.e)
MAX_LMB = 8192

# --------------------------------------------------------------------------------------
# architecture table (qarv/zoo.py:35-88).  Entries:
#   ('down', cin, cout, rate) | ('blk', C, k, ratio) | ('key', name) | ('up', cin, cout, rate)
#   ('lat', width, zdim, enc_key, enc_width, k, ratio) | ('stop',)
# --------------------------------------------------------------------------------------

def qarv_base_arch():
    ch = 128
    e = [192, ch * 3, ch * 4, ch * 4, ch * 4]
    enc = [('down', 3, e[0], 4)]
    enc += [('blk', e[0], 7, 2)] * 7
    enc += [('down', e[0], e[1], 2)] + [('blk', e[1], 7, 2)] * 6 + [('key', 'enc_s8'), ('blk', e[1], 7, 2)]
    enc += [('down', e[1], e[2], 2)] + [('blk', e[2], 5, 2)] * 6 + [('key', 'enc_s16'), ('blk', e[2], 7, 2)]
    enc += [('down', e[2], e[3], 2)] + [('blk', e[3], 3, 2)] * 4 + [('key', 'enc_s32'), ('blk', e[3], 7, 2)]
    enc += [('down', e[3], e[4], 2)] + [('blk', e[4], 1, 2)] * 4 + [('key', 'enc_s64')]
    d = [ch * 4, ch * 4, ch * 3, ch * 2, ch * 1]
    z = [32, 32, 96, 8]
    dec = [('lat', d[0], z[0], 'enc_s64', e[4], 1, 4), ('blk', d[0], 1, 4), ('up', d[0], d[1], 2)]
    dec += [('blk', d[1], 3, 3)] + [('lat', d[1], z[1], 'enc_s32', e[3], 3, 3)] * 2 + [('blk', d[1], 3, 3), ('up', d[1], d[2], 2)]
    dec += [('blk', d[2], 5, 2)] + [('lat', d[2], z[2], 'enc_s16', e[2], 5, 2)] * 3 + [('blk', d[2], 5, 2), ('up', d[2], d[3], 2)]
    dec += [('blk', d[3], 7, 1.75)] + [('lat', d[3], z[3], 'enc_s8', e[1], 7, 1.75)] * 3 + [('stop',)]
    dec += [('blk', d[3], 7, 1.75), ('up', d[3], d[4], 2)]
    dec += [('blk', d[4], 7, 1.5)] * 8 + [('up', d[4], 3, 4)]
    return dict(enc=enc, dec=dec, im_shift=-0.4546259594901961, im_scale=3.67572653978347,
                max_stride=64, sin_period=64, embed_dim=256, lmb_range=(16.0, 2048.0))


# --------------------------------------------------------------------------------------
# seeded, order-independent "sensitised" weights (default init has gamma=1e-6 so a
# broken MLP kernel would go unnoticed).  Every tensor is drawn from its own generator keyed by
# (seed, crc32(name)), so the oracle, the reference and the product all get identical values
# from nothing but the key names and shapes.
# --------------------------------------------------------------------------------------

def _key_generator(seed, key):
    import zlib
    g = torch.Generator()
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def sensitised_tensor(key, shape, seed=0, wide_heads=True):
    """wide_heads: widen the posterior / prior head weights (qarv: symbols beyond {-1,0,1}); the rd fixtures use
    False, which keeps its continuous KL in a sane range."""
    g = _key_generator(seed, key)
    shape = tuple(shape)
    if key.endswith('gamma'):
        return torch.rand(shape, generator=g) * 0.4 + 0.2
    if key == 'bias' or key == 'decoder.bias':
        return torch.randn(shape, generator=g) * 0.5
    if key.endswith('norm.weight'):         # affine LayerNorm gain (qres): around 1
        return 1.0 + (torch.rand(shape, generator=g) * 2 - 1) * 0.3
    if key.endswith('prior.bias') or key.endswith('prior.c4.bias'):          # spread prior means and log-scales
        return torch.randn(shape, generator=g) * 0.7
    if key.endswith('.bias'):
        return torch.randn(shape, generator=g) * 0.05
    assert key.endswith('.weight'), key
    fan_in = int(np.prod(shape[1:]))
    bound = 1.0 / math.sqrt(fan_in)
    if wide_heads and (key.endswith('posterior.weight') or key.endswith('posterior.c4.weight')):
        bound *= 3.0                                       # wider posterior means -> symbols beyond {-1,0,1}
    if wide_heads and (key.endswith('prior.weight') or key.endswith('prior.c4.weight')):
        bound *= 2.0
    return (torch.rand(shape, generator=g) * 2 - 1) * bound


def sensitised_state_dict(named_shapes, seed=0, wide_heads=True):
    """named_shapes: iterable of (key, shape) for the floating-point parameters."""
    return OrderedDict((k, sensitised_tensor(k, s, seed, wide_heads)) for k, s in named_shapes)


def qarv_param_shapes(arch=None):
    """(key, shape) of every parameter of the qarv model, in module order (Appendix B of SURVEY)."""
    arch = arch or qarv_base_arch()
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
            _, cin, cout, r = ent
            out += [(p + 'weight', (cout, cin, r, r)), (p + 'bias', (cout,))]
        elif ent[0] == 'blk':
            blk(p, ent[1], ent[2], ent[3])
    for i, ent in enumerate(arch['dec']):
        p = f'dec_blocks.{i}.'
        if ent[0] == 'blk':
            blk(p, ent[1], ent[2], ent[3])
        elif ent[0] == 'up':
            _, cin, cout, r = ent
            out += [(p + '0.weight', (cout * r * r, cin, 1, 1)), (p + '0.bias', (cout * r * r,))]
        elif ent[0] == 'lat':
            _, W, zd, _key, We, k, ratio = ent
            blk(p + 'resnet_front.', W, k, ratio)
            blk(p + 'resnet_end.', W, k, ratio)
            blk(p + 'posterior0.', We, k, 2)
            blk(p + 'posterior1.', W, k, 2)
            blk(p + 'posterior2.', W, k, 2)
            out += [(p + 'post_merge.weight', (W, W + We, 1, 1)), (p + 'post_merge.bias', (W,)),
                    (p + 'posterior.weight', (zd, W, 3, 3)), (p + 'posterior.bias', (zd,)),
                    (p + 'z_proj.weight', (W, zd, 1, 1)), (p + 'z_proj.bias', (W,)),
                    (p + 'prior.weight', (2 * zd, W, 1, 1)), (p + 'prior.bias', (2 * zd,))]
    width0 = arch['dec'][0][1]
    out.append(('bias', (1, width0, 1, 1)))
    for j in (0, 2):
        out += [(f'lmb_embedding.{j}.weight', (E, E)), (f'lmb_embedding.{j}.bias', (E,))]
    return out


# --------------------------------------------------------------------------------------
# entropy model arithmetic (CompressAI GaussianConditional via DiscretizedGaussian)
# --------------------------------------------------------------------------------------

def default_scale_table():
    # entropy_coding.py:73-75
    return torch.exp(torch.linspace(math.log(0.11), math.log(20.0), steps=64))


_STD_NORMAL = torch.distributions.Normal(loc=0, scale=1)


def std_normal_cdf(t):
    # entropy_coding.py:81-82 -> td.Normal(0,1).cdf
    return _STD_NORMAL.cdf(t)


def prior_transform(prior_out):
    """qarv/model.py:51-53. prior_out: [B, 2*zdim, h, w] -> pm, pv"""
    pm, plogv = prior_out.chunk(2, dim=1)
    plogv = F.softplus(plogv + 2.3) - 2.3
    pv = torch.exp(plogv)
    return pm, pv


def std_normal_cdf_erfc(t):
    # CompressAI GaussianConditional._standardized_cumulative (what qres34m runs, unmodified)
    return float(0.5) * torch.erfc(float(-(2 ** -0.5)) * t)


def eval_quantize_likelihood(qm, pm, pv, scale_bound=0.11, likelihood_bound=1e-9, cdf=None):
    """GaussianConditional.forward(training=False): returns z, P."""
    std_normal_cdf = cdf or globals()['std_normal_cdf']
    z = qm.clone()
    z -= pm
    z = torch.round(z)
    z += pm
    values = torch.abs(z - pm)
    scales = torch.max(pv, torch.tensor([scale_bound]))
    upper = std_normal_cdf((0.5 - values) / scales)
    lower = std_normal_cdf((-0.5 - values) / scales)
    lik = upper - lower
    lik = torch.max(lik, torch.tensor([likelihood_bound]))
    return z, lik


def symbols(qm, pm):
    return torch.round(qm - pm).int()


def build_indexes(pv, scale_table=None, scale_bound=0.11):
    """GaussianConditional.build_indexes (Appendix A.2)."""
    scale_table = default_scale_table() if scale_table is None else scale_table
    s = torch.max(pv, torch.tensor([scale_bound]))
    idx = torch.full(s.shape, len(scale_table) - 1, dtype=torch.int32)
    for t in scale_table[:-1]:
        idx -= (s <= t).int()
    return idx


def gaussian_log_prob_mass(mean, scale, x, bin_size=1.0, prob_clamp=1e-6):
    """entropy_coding.py:17-49 (training likelihood with tail fallback)."""
    dist = torch.distributions.Normal(mean, scale)
    mass = dist.cdf(x + 0.5 * bin_size) - dist.cdf(x - 0.5 * bin_size)
    return torch.where(mass > prob_clamp, torch.log(mass.clamp(min=1e-8)),
                       dist.log_prob(x) + math.log(bin_size))


def pmf_to_quantized_cdf(pmf, precision=16):
    """CompressAI cpp_exts/ops/ops.cpp pmf_to_quantized_cdf. pmf: list of python floats (fp32 values)."""
    n = len(pmf)
    cdf = [0] * (n + 1)
    for i, p in enumerate(pmf):
        v = np.float32(np.float32(p) * np.float32(1 << precision))
        cdf[i + 1] = int(np.floor(float(v) + 0.5))
    total = sum(cdf)
    cdf = [((1 << precision) * c) // total for c in cdf]
    for i in range(1, n + 1):
        cdf[i] += cdf[i - 1]
    cdf[-1] = 1 << precision
    for i in range(n):
        if cdf[i] == cdf[i + 1]:
            best_freq, best_steal = 1 << 62, -1
            for j in range(n):
                freq = cdf[j + 1] - cdf[j]
                if 1 < freq < best_freq:
                    best_freq, best_steal = freq, j
            if best_steal < i:
                for j in range(best_steal + 1, i + 1):
                    cdf[j] -= 1
            else:
                for j in range(i + 1, best_steal + 1):
                    cdf[j] += 1
    return cdf


def build_cdf_tables(scale_table=None, tail_mass=1e-9, precision=16, cdf=None):
    """GaussianConditional.update() (Appendix A.4) -> (cdf [n, L+2] int32, cdf_length [n], offset [n])."""
    import scipy.stats
    std_normal_cdf = cdf or globals()['std_normal_cdf']
    scale_table = default_scale_table() if scale_table is None else scale_table
    multiplier = -scipy.stats.norm.ppf(tail_mass / 2)
    center = torch.ceil(scale_table * multiplier).int()
    length = 2 * center + 1
    max_length = int(length.max())
    samples = torch.abs(torch.arange(max_length).int() - center[:, None]).float()
    sc = scale_table.unsqueeze(1).float()
    upper = std_normal_cdf((0.5 - samples) / sc)
    lower = std_normal_cdf((-0.5 - samples) / sc)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    cdf = torch.zeros((len(length), max_length + 2), dtype=torch.int32)
    for i in range(len(length)):
        prob = torch.cat((pmf[i, :length[i]], tail[i]), dim=0).tolist()
        c = pmf_to_quantized_cdf(prob, precision)
        cdf[i, :len(c)] = torch.tensor(c, dtype=torch.int32)
    return cdf, (length + 2).int(), (-center).int()


# ---- rANS (ryg rans64, CompressAI interface): pure python, for small cases only ----
_RANS_L = 1 << 31
_BYP = 4
_BYP_MAX = (1 << _BYP) - 1


def rans_encode(symbols_, indexes, cdf, cdf_length, offset, precision=16):
    cdfs, sizes, offs = cdf.tolist(), cdf_length.tolist(), offset.tolist()
    syms = []
    for s, ci in zip(symbols_, indexes):
        c = cdfs[ci]
        max_value = sizes[ci] - 2
        value = s - offs[ci]
        raw = 0
        if value < 0:
            raw, value = -2 * value - 1, max_value
        elif value >= max_value:
            raw, value = 2 * (value - max_value), max_value
        syms.append((c[value], c[value + 1] - c[value], False))
        if value == max_value:
            nb = 0
            while (raw >> (nb * _BYP)) != 0:
                nb += 1
            val = nb
            while val >= _BYP_MAX:
                syms.append((_BYP_MAX, 0, True))
                val -= _BYP_MAX
            syms.append((val, 0, True))
            for j in range(nb):
                syms.append(((raw >> (j * _BYP)) & _BYP_MAX, 0, True))
    x = _RANS_L
    out = []
    for start, freq, bypass in reversed(syms):
        if bypass:
            f = 1 << (16 - _BYP)
            if x >= ((_RANS_L >> 16) << 32) * f:
                out.append(x & 0xFFFFFFFF)
                x >>= 32
            x = (x << _BYP) | start
        else:
            if x >= ((_RANS_L >> precision) << 32) * freq:
                out.append(x & 0xFFFFFFFF)
                x >>= 32
            x = ((x // freq) << precision) + (x % freq) + start
    out += [(x >> 32) & 0xFFFFFFFF, x & 0xFFFFFFFF]
    out.reverse()
    return struct.pack(f'<{len(out)}I', *out)


def rans_decode(data, indexes, cdf, cdf_length, offset, precision=16):
    cdfs, sizes, offs = cdf.tolist(), cdf_length.tolist(), offset.tolist()
    words = struct.unpack(f'<{len(data) // 4}I', data)
    x = words[0] | (words[1] << 32)
    pos = 2
    mask = (1 << precision) - 1
    out = []

    def bits():
        nonlocal x, pos
        v = x & _BYP_MAX
        x >>= _BYP
        if x < _RANS_L:
            x = (x << 32) | words[pos]
            pos += 1
        return v

    for ci in indexes:
        c, size = cdfs[ci], sizes[ci]
        max_value = size - 2
        cum = x & mask
        s = 0
        while s + 1 < size and c[s + 1] <= cum:
            s += 1
        x = (c[s + 1] - c[s]) * (x >> precision) + cum - c[s]
        if x < _RANS_L:
            x = (x << 32) | words[pos]
            pos += 1
        value = s
        if value == max_value:
            val = bits()
            nb = val
            while val == _BYP_MAX:
                val = bits()
                nb += val
            raw = 0
            for j in range(nb):
                raw |= bits() << (j * _BYP)
            value = (-(raw >> 1) - 1) if (raw & 1) else ((raw >> 1) + max_value)
        out.append(value + offs[ci])
    return out


# ---- byte-string container (lvae/utils/coding.py:26-70) ----

def pack_byte_strings(strings):
    lengths = [len(s) for s in strings]
    return struct.pack('B', len(lengths)) + struct.pack(f'{len(lengths)}I', *lengths) + b''.join(strings)


def unpack_byte_string(string):
    num = struct.unpack('B', string[:1])[0]
    lengths = struct.unpack(f'{num}I', string[1:1 + 4 * num])
    body = string[1 + 4 * num:]
    assert sum(lengths) == len(body)
    out, pos = [], 0
    for n in lengths:
        out.append(body[pos:pos + n])
        pos += n
    return out


# --------------------------------------------------------------------------------------
# network
# --------------------------------------------------------------------------------------

def sinusoidal_embedding(values, dim=256, max_period=64):
    exponents = torch.linspace(0, 1, steps=dim // 2)
    freqs = torch.pow(max_period, -1.0 * exponents)
    args = values.view(-1, 1) * freqs.view(1, dim // 2)
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def lmb_embedding(sd, lmb, arch):
    scaled = torch.log(lmb) * arch['sin_period'] / math.log(MAX_LMB)
    e = sinusoidal_embedding(scaled, dim=arch['embed_dim'], max_period=arch['sin_period'])
    e = F.linear(e, sd['lmb_embedding.0.weight'], sd['lmb_embedding.0.bias'])
    e = F.gelu(e)
    return F.linear(e, sd['lmb_embedding.2.weight'], sd['lmb_embedding.2.bias'])


def convnext_block(sd, p, x, emb):
    """common.py:142-161; x NCHW."""
    w = sd[p + 'conv_dw.weight']
    C, k = w.shape[0], w.shape[-1]
    y = F.conv2d(x, w, sd[p + 'conv_dw.bias'], padding=(k - 1) // 2, groups=C)
    y = y.permute(0, 2, 3, 1).contiguous()
    y = F.layer_norm(y, (C,), eps=1e-6)
    e = F.linear(F.gelu(emb), sd[p + 'embedding_layer.1.weight'], sd[p + 'embedding_layer.1.bias'])
    e = e.unflatten(1, (1, 1, 2 * C))
    shift, scale = torch.chunk(e, chunks=2, dim=-1)
    y = y * (1 + scale) + shift
    y = F.linear(y, sd[p + 'mlp.fc1.weight'], sd[p + 'mlp.fc1.bias'])
    y = F.gelu(y)
    y = F.linear(y, sd[p + 'mlp.fc2.weight'], sd[p + 'mlp.fc2.bias'])
    y = y.permute(0, 3, 1, 2).contiguous()
    y = y.mul(sd[p + 'gamma'])
    return y + x


def encoder(sd, arch, x, emb):
    feats = OrderedDict()
    for i, ent in enumerate(arch['enc']):
        p = f'encoder.enc_blocks.{i}.'
        if ent[0] == 'down':
            x = F.conv2d(x, sd[p + 'weight'], sd[p + 'bias'], stride=ent[3])
        elif ent[0] == 'blk':
            x = convnext_block(sd, p, x, emb)
        elif ent[0] == 'key':
            feats[ent[1]] = x
    return feats


def posterior_branch(sd, p, feature, enc_feature, emb):
    e = convnext_block(sd, p + 'posterior0.', enc_feature, emb)
    f = convnext_block(sd, p + 'posterior1.', feature, emb)
    m = torch.cat([f, e], dim=1)
    m = F.conv2d(m, sd[p + 'post_merge.weight'], sd[p + 'post_merge.bias'])
    m = convnext_block(sd, p + 'posterior2.', m, emb)
    return F.conv2d(m, sd[p + 'posterior.weight'], sd[p + 'posterior.bias'], padding=1)


class LayerRecord(dict):
    """Per-latent-layer outputs: kl [B,z,h,w], z, sym (int32), idx (int32), qm, pm, pv."""


def top_down(sd, arch, emb, nB, nH, nW, enc_feats=None, mode='eval', latents=None, noise=None,
             stop_at_flag=False):
    """qarv/model.py:294-315 (+ :77-121 per latent block).
    mode: 'eval' (hard quantisation, K12) | 'train' (uniform noise from `noise` list, K13) |
          'given' (z taken from `latents`, the decompress / sampling-with-latents path).
    """
    feature = sd['bias'].expand(nB, -1, nH, nW)
    records = []
    li = 0
    for i, ent in enumerate(arch['dec']):
        p = f'dec_blocks.{i}.'
        if ent[0] == 'blk':
            feature = convnext_block(sd, p, feature, emb)
        elif ent[0] == 'up':
            r = ent[3]
            feature = F.pixel_shuffle(F.conv2d(feature, sd[p + '0.weight'], sd[p + '0.bias']), r)
        elif ent[0] == 'stop':
            if stop_at_flag:
                return None, records
        elif ent[0] == 'lat':
            feature = convnext_block(sd, p + 'resnet_front.', feature, emb)
            pm, pv = prior_transform(F.conv2d(feature, sd[p + 'prior.weight'], sd[p + 'prior.bias']))
            rec = LayerRecord(pm=pm, pv=pv)
            if mode == 'given':
                z = latents[li]
            else:
                qm = posterior_branch(sd, p, feature, enc_feats[ent[3]], emb)
                rec['qm'] = qm
                if mode == 'train':
                    z = qm + noise[li]
                    rec['kl'] = -1.0 * gaussian_log_prob_mass(pm, pv, z)
                else:
                    z, lik = eval_quantize_likelihood(qm, pm, pv)
                    rec['kl'] = -1.0 * torch.log(lik)
                    rec['sym'] = symbols(qm, pm)
            rec['idx'] = build_indexes(pv)
            rec['z'] = z
            records.append(rec)
            li += 1
            feature = feature + F.conv2d(z, sd[p + 'z_proj.weight'], sd[p + 'z_proj.bias'])
            feature = convnext_block(sd, p + 'resnet_end.', feature, emb)
    return feature, records


@torch.no_grad()
def qarv_forward(sd, im, lmb, arch=None, mode='eval', noise=None):
    """VariableRateLossyVAE.forward (qarv/model.py:317-363) in eval mode (or train-mode forward with
    externally supplied uniform noise). Returns dict with loss, bppix, mse, psnr, per-image kl / mse,
    x_hat, im_hat and the per-layer records."""
    arch = arch or qarv_base_arch()
    nB, imC, imH, imW = im.shape
    assert imH % arch['max_stride'] == 0 and imW % arch['max_stride'] == 0
    x = im.clone().add_(arch['im_shift']).mul_(arch['im_scale'])
    emb = lmb_embedding(sd, lmb, arch)
    feats = encoder(sd, arch, x, emb)
    x_hat, records = top_down(sd, arch, emb, nB, imH // arch['max_stride'], imW // arch['max_stride'],
                              enc_feats=feats, mode=mode, noise=noise)
    kls = [r['kl'].sum(dim=(1, 2, 3)) for r in records]
    ndims = float(imC * imH * imW)
    kl = sum(kls) / ndims
    x_target = im.clone().add_(-0.5).mul_(2.0)
    distortion = F.mse_loss(x_hat, x_target, reduction='none').mean(dim=(1, 2, 3))
    loss = (kl + lmb * distortion).mean(0)
    im_hat = x_hat.clone().clamp_(min=-1.0, max=1.0).mul_(0.5).add_(0.5)
    im_mse = F.mse_loss(im_hat, im, reduction='mean')
    return dict(loss=loss, bppix=kl.mean(0).item() * LOG2_E * imC, mse=distortion.mean(0).item(),
                psnr=-10 * math.log10(im_mse.item()), kl_per_image=kl, mse_per_image=distortion,
                x_hat=x_hat, im_hat=im_hat, records=records)


@torch.no_grad()
def qarv_compress(sd, im, lmb, arch=None, tables=None):
    """qarv/model.py:516-529: one image -> bytes (header + 9 rANS streams)."""
    arch = arch or qarv_base_arch()
    assert im.shape[0] == 1
    tables = tables or build_cdf_tables()
    nB, imC, imH, imW = im.shape
    x = im.clone().add_(arch['im_shift']).mul_(arch['im_scale'])
    lmb_t = torch.full((1,), float(lmb))
    emb = lmb_embedding(sd, lmb_t, arch)
    feats = encoder(sd, arch, x, emb)
    _, records = top_down(sd, arch, emb, nB, imH // 64, imW // 64, enc_feats=feats, mode='eval',
                          stop_at_flag=True)
    strings = [rans_encode(r['sym'][0].reshape(-1).tolist(), r['idx'][0].reshape(-1).tolist(), *tables)
               for r in records]
    body = pack_byte_strings(strings)
    return struct.pack('f', lmb) + struct.pack('3H', nB, imH // 64, imW // 64) + body


@torch.no_grad()
def qarv_decompress(sd, string, arch=None, tables=None):
    """qarv/model.py:531-557."""
    arch = arch or qarv_base_arch()
    tables = tables or build_cdf_tables()
    lmb = struct.unpack('f', string[:4])[0]
    nB, nH, nW = struct.unpack('3H', string[4:10])
    strings = unpack_byte_string(string[10:])
    emb = lmb_embedding(sd, torch.full((nB,), float(lmb)), arch)
    # decode layer by layer: the prior of layer i needs z_<i
    feature = sd['bias'].expand(nB, -1, nH, nW)
    si = 0
    for i, ent in enumerate(arch['dec']):
        p = f'dec_blocks.{i}.'
        if ent[0] == 'blk':
            feature = convnext_block(sd, p, feature, emb)
        elif ent[0] == 'up':
            feature = F.pixel_shuffle(F.conv2d(feature, sd[p + '0.weight'], sd[p + '0.bias']), ent[3])
        elif ent[0] == 'lat':
            feature = convnext_block(sd, p + 'resnet_front.', feature, emb)
            pm, pv = prior_transform(F.conv2d(feature, sd[p + 'prior.weight'], sd[p + 'prior.bias']))
            idx = build_indexes(pv)
            vals = rans_decode(strings[si], idx[0].reshape(-1).tolist(), *tables)
            si += 1
            z = torch.tensor(vals, dtype=torch.int32).reshape(pm.shape).type_as(pm)
            z += pm
            feature = feature + F.conv2d(z, sd[p + 'z_proj.weight'], sd[p + 'z_proj.bias'])
            feature = convnext_block(sd, p + 'resnet_end.', feature, emb)
    return feature.clone().clamp_(min=-1.0, max=1.0).mul_(0.5).add_(0.5)
