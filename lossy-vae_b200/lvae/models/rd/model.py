"""R(D) hierarchical VAE with a continuous Gaussian posterior (rd): the reference's model surface over the B200 engine.

Mirrors `VariableRateLossyVAE`, `LatentVariableBlock`, `ConvNeXtAdaLNPatchDown` and `FeatureExtractor` of the
reference (lvae/models/rd/model.py:16-24,140-244,247-590): same constructor config, same module / parameter names
(including the reference's `downsapmle` spelling, so upstream checkpoints load), same `forward` /
`forward_end2end` / sampling methods and return types.  The arithmetic runs in liblvae_b200 (fused
linear_sqrt + softplus + gaussian_kl + reparameterised sample kernel `lvae_rd_latent`, include/lvae_b200.h),
sequenced by `lvae.engine.QarvEngine`; there is no ATen / CPU fallback.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import common


class ConvNeXtAdaLNPatchDown(common.ConvNeXtBlockAdaLN):
    """ConvNeXt-AdaLN block followed by a patch-downsampling conv (rd/model.py:16-24)."""
    def __init__(self, in_ch, out_ch, down_rate=2, **kwargs):
        super().__init__(in_ch, **kwargs)
        self.downsapmle = common.patch_downsample(in_ch, out_ch, rate=down_rate)      # (sic) reference key name


class LatentVariableBlock(nn.Module):
    """Parameter container of one latent layer (rd/model.py:140-227): prior / posterior heads emit
    (mean_raw | std_raw) with mean = linear_sqrt(raw), std = softplus(raw, beta = ln 2)."""
    softplus_beta = math.log(2)

    def __init__(self, width, zdim, embed_dim, enc_width=None, kernel_size=7, mlp_ratio=2):
        super().__init__()
        self.in_channels = width
        self.out_channels = width
        self.zdim = zdim
        block = common.ConvNeXtBlockAdaLN
        enc_width = enc_width or width
        self.enc_width = enc_width
        self.resnet_front = block(width, embed_dim, kernel_size=kernel_size, mlp_ratio=mlp_ratio)
        self.resnet_end = block(width, embed_dim, kernel_size=kernel_size, mlp_ratio=mlp_ratio)
        self.posterior0 = block(enc_width, embed_dim, kernel_size=kernel_size)
        self.posterior1 = block(width, embed_dim, kernel_size=kernel_size)
        self.posterior2 = block(width, embed_dim, kernel_size=kernel_size)
        self.post_merge = common.conv_k1s1(width + enc_width, width)
        self.posterior = common.conv_k3s1(width, zdim * 2)
        self.prior = common.conv_k1s1(width, zdim * 2)
        self.z_proj = common.conv_k1s1(zdim, width)
        self.is_latent_block = True


class VariableRateLossyVAE(nn.Module):
    log2_e = math.log2(math.e)
    MAX_LMB = 8192
    family = 'rd'

    def __init__(self, config: dict):
        super().__init__()
        self.encoder = common.FeatureExtractorWithEmbedding(config.pop('enc_blocks'))
        self.dec_blocks = nn.ModuleList(config.pop('dec_blocks'))
        width = self.dec_blocks[0].in_channels
        self.bias = nn.Parameter(torch.zeros(1, width, 1, 1))
        self.num_latents = len([b for b in self.dec_blocks if getattr(b, 'is_latent_block', False)])
        self.distortion_name = 'mse'

        low, high = config['lmb_range']
        self.lmb_range = (float(low), float(high))
        self.default_lmb = self.lmb_range[1]
        self.lmb_embed_dim = config['lmb_embed_dim']
        self.lmb_embedding = nn.Sequential(
            common.ParamLinear(self.lmb_embed_dim[0], self.lmb_embed_dim[1]),
            nn.GELU(),
            common.ParamLinear(self.lmb_embed_dim[1], self.lmb_embed_dim[1]),
        )
        self._sin_period = config['sin_period']
        self.im_shift = float(config['im_shift'])
        self.im_scale = float(config['im_scale'])
        self.max_stride = config['max_stride']
        self.register_buffer('_dummy', torch.zeros(1), persistent=False)
        self._logging_images = config.get('log_images', [])
        self._flops_mode = False
        self.precision = config.get('precision', 'f16x3')        # see qarv/model.py
        self.__dict__['_engine'] = None

    # ------------------------------------------------------------------ engine plumbing
    @property
    def engine(self):
        if self.__dict__.get('_engine') is None:
            from ...engine import QarvEngine
            self.__dict__['_engine'] = QarvEngine(self)
        return self.__dict__['_engine']

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == '_engine' else copy.deepcopy(v, memo)
        return new

    def _device(self):
        return self._dummy.device

    # ------------------------------------------------------------------ reference surface
    def sample_lmb(self, n):
        """log-uniform in lmb_range (rd/model.py:338-347)"""
        low, high = self.lmb_range
        low, high = math.log(low), math.log(high)
        transformed = low + (high - low) * torch.rand(n, device=self._device())
        return torch.exp(transformed)

    def expand_to_tensor(self, input_, n):
        assert isinstance(input_, (torch.Tensor, float, int)), f'{type(input_)=}'
        if isinstance(input_, torch.Tensor) and (input_.numel() == 1):
            input_ = input_.item()
        if isinstance(input_, (float, int)):
            input_ = torch.full(size=(n,), fill_value=float(input_), device=self._device())
        assert input_.shape == (n,), f'{input_=}, {input_.shape=}'
        return input_

    def _check_image(self, im):
        assert im.dim() == 4 and im.shape[1] == 3, f'expected [B,3,H,W], got {tuple(im.shape)}'
        assert (im.shape[2] % self.max_stride == 0) and (im.shape[3] % self.max_stride == 0)
        assert not im.requires_grad and im.dtype == torch.float32

    def process_output(self, x: torch.Tensor):
        assert not x.requires_grad
        return x.clone().clamp_(min=-1.0, max=1.0).mul_(0.5).add_(0.5)

    def forward_end2end(self, im: torch.Tensor, lmb: torch.Tensor, get_latents=False, noise=None):
        """Returns (x_hat, [stats_i]) with stats_i['kl'] [B,z,h,w] in nats (and 'z' with get_latents)
        (rd/model.py:377-397).  `noise`: optional per-layer N(0,1) tensors (default: drawn on the device)."""
        self._check_image(im)
        lmb = self.expand_to_tensor(lmb, n=im.shape[0])
        res = self.engine.run(im, lmb, mode='eval', want_elem=True, noise=noise)
        stats = []
        for li in range(self.num_latents):
            st = dict(kl=res['kl_elem'][li])
            if get_latents:
                st['z'] = res['z'][li]
            stats.append(st)
        return res['x_hat'], stats

    def forward(self, batch, lmb=None, return_rec=False, noise=None):
        """R-D objective of a batch (rd/model.py:399-445): OrderedDict(loss, bppix, mse, psnr[, im_hat])."""
        im = batch[0] if isinstance(batch, (tuple, list)) else batch
        nB, imC, imH, imW = im.shape
        if self._flops_mode:
            raise NotImplementedError('_flops_mode is a profiling hook of the ATen modules; use bench.py')
        if lmb is None:
            lmb = self.sample_lmb(n=nB)
        assert isinstance(lmb, torch.Tensor) and lmb.shape == (nB,)
        self._check_image(im)
        res = self.engine.run(im, lmb.to(self._device(), torch.float32), mode='eval', want_elem=False,
                              want_im_hat=return_rec, noise=noise)
        host = res['stats_host']
        stats = OrderedDict()
        stats['loss'] = res['stats'][0]
        stats['bppix'] = float(host[1]) * self.log2_e * imC
        stats[self.distortion_name] = float(host[2])
        stats['psnr'] = -10 * math.log10(float(host[3]))
        if return_rec:
            stats['im_hat'] = res['im_hat']
        return stats

    @torch.no_grad()
    def forward_stream(self, batches, lmb=None, depth=2):
        """Throughput form of forward() (an extension; see lvae.models.qarv.model.VariableRateLossyVAE.forward_stream):
        one OrderedDict of Python floats per batch, batch i+1 copied in while batch i runs."""
        shapes = []

        def items():
            for batch in batches:
                im = batch[0] if isinstance(batch, (tuple, list)) else batch
                self._check_image(im)
                nB = im.shape[0]
                l = self.sample_lmb(n=nB) if lmb is None else self.expand_to_tensor(lmb, n=nB)
                shapes.append(im.shape[1])
                yield im, l.to(self._device(), torch.float32)
        for i, res in enumerate(self.engine.run_stream(items(), mode='eval', depth=depth)):
            host, imC = res['stats_host'], shapes[i]
            yield OrderedDict([('loss', float(host[0])), ('bppix', float(host[1]) * self.log2_e * imC),
                               (self.distortion_name, float(host[2])), ('psnr', -10 * math.log10(float(host[3])))])

    def conditional_sample(self, lmb, latents, emb=None, bhw_repeat=None, t=1.0):
        """rd/model.py:447-486"""
        if latents is None:
            latents = [None] * self.num_latents
            assert bhw_repeat is not None, 'bhw_repeat should be provided'
            nB, nH, nW = bhw_repeat
        else:
            assert (bhw_repeat is None) and (len(latents) == self.num_latents)
            nB, _, nH, nW = latents[0].shape
        assert emb is None, 'pre-computed embeddings are not part of the engine interface'
        lmb = self.expand_to_tensor(lmb, n=nB)
        return self.engine.sample(lmb, latents, (nB, nH, nW), float(t))

    def unconditional_sample(self, lmb, bhw_repeat, t=1.0):
        return self.conditional_sample(lmb, latents=None, bhw_repeat=bhw_repeat, t=t)
