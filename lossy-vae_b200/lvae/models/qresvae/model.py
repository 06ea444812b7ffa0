"""Fixed-rate hierarchical VAE (qres34m family): the reference's model surface over the B200 engine.

Mirrors `HierarchicalVAE`, `BottomUpEncoder`, `TopDownDecoder`, `QLatentBlockX`, `VDBlock`, `MyConvNeXtBlock`,
`MyConvNeXtPatchDown` and `MSEOutputNet` of the reference (lvae/models/qresvae/model.py:95-149,163-207,210-391,
456-725): same constructor config, same module / parameter / buffer names (state-dict compatible, including the
`downsapmle` spelling), same methods (`forward`, `forward_eval`, `forward_get_latents`, `uncond_sample`,
`cond_sample`, `compress_mode`, `compress`, `decompress`, `compress_file`, `decompress_file`), return types and
assertion conventions.  The arithmetic runs in liblvae_b200 sequenced by `lvae.engine.QarvEngine` (family 'qres');
there is no ATen / CPU fallback.

Differences from qarv that the engine handles: no lambda embedding (affine LayerNorm blocks), prior / posterior
heads are VDBlocks (1x1 -> 3x3 -> 3x3 -> 1x1 with GELU before every conv), z_proj is conv -> GELU -> 1x1, the
likelihood uses CompressAI's default erfc-based CDF and the 0.1 ... 20 scale table, the loss is kl + lambda * mse
with lambda fixed at construction, and `compress` handles batches (list-of-lists container, pickled by
`compress_file`)."""
import math
import pickle
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import common
from .. import entropy_coding
from ...utils import coding


class MSEOutputNet(nn.Module):
    def __init__(self, mse_lmb):
        super().__init__()
        self.mse_lmb = float(mse_lmb)
        self.loss_name = 'mse'

    def mean(self, x_hat, temprature=None):
        return x_hat
    sample = mean


class GaussianNLLOutputNet(nn.Module):
    """Output net of `qres34m_lossless` (reference model.py:16-94): two patch_upsample heads turn the decoder's last feature
    into the mean and log-scale of a per-pixel discretised Gaussian over the 8-bit image (bin 1/127.5); the training /
    eval loss is its negative log-likelihood, `compress` codes the image's own residual round(x / bin - mean) with rANS
    against a 128-scale table, which makes the codec lossless.  Parameter container (state-dict keys `conv_mean.0.*`,
    `conv_scale.0.*` as in the reference); the arithmetic is lvae_gemm (heads) + csrc/outnet.cu, sequenced by the engine."""
    def __init__(self, conv_mean, conv_scale, bin_size=1 / 127.5):
        super().__init__()
        self.conv_mean = conv_mean
        self.conv_scale = conv_scale
        assert abs(bin_size - 1 / 127.5) < 1e-12, 'the kernels are built for 8-bit images (bin 1/127.5)'
        self.bin_size = bin_size
        self.loss_name = 'nll'
        self.mse_lmb = 1.0          # the engine's loss assembly is kl + lmb * (per-image out-net loss): lmb = 1 here

    def update(self):
        """128 scales from the 0.11 scale bound to 20 and their CDF tables (reference model.py:58-66)."""
        self.discrete_gaussian = entropy_coding.GaussianConditional(None, scale_bound=0.11)
        self.discrete_gaussian = self.discrete_gaussian.to(device=next(self.parameters()).device)
        scale_table = torch.exp(torch.linspace(math.log(0.11), math.log(20), steps=128))
        self.discrete_gaussian.update_scale_table(scale_table)
        self.discrete_gaussian.update()


class VDBlock(nn.Module):
    """c1 1x1 -> c2 -> c3 (3x3 or 1x1) -> c4 1x1, GELU before every conv (reference model.py:120-149): container."""
    def __init__(self, in_ch, hidden_ch=None, out_ch=None, residual=True, use_3x3=True, zero_last=False):
        super().__init__()
        out_ch = out_ch or in_ch
        hidden_ch = hidden_ch or round(in_ch * 0.25)
        self.in_channels, self.out_channels, self.hidden_channels = in_ch, out_ch, hidden_ch
        self.residual = residual
        mid = common.conv_k3s1 if use_3x3 else common.conv_k1s1
        self.c1 = common.conv_k1s1(in_ch, hidden_ch)
        self.c2 = mid(hidden_ch, hidden_ch)
        self.c3 = mid(hidden_ch, hidden_ch)
        self.c4 = common.conv_k1s1(hidden_ch, out_ch, zero_weights=zero_last)

    def residual_scaling(self, N):
        self.c4.weight.data.mul_(math.sqrt(1 / N))


class MyConvNeXtBlock(common.ConvNeXtBlockLN):
    def __init__(self, dim, mlp_ratio=2, **kwargs):
        super().__init__(dim, mlp_ratio=mlp_ratio, **kwargs)


class MyConvNeXtPatchDown(MyConvNeXtBlock):
    def __init__(self, in_ch, out_ch, down_rate=2, mlp_ratio=2, kernel_size=7):
        super().__init__(in_ch, mlp_ratio=mlp_ratio, kernel_size=kernel_size)
        self.downsapmle = common.patch_downsample(in_ch, out_ch, rate=down_rate)      # (sic) reference key name


class BottomUpEncoder(nn.Module):
    def __init__(self, blocks):
        super().__init__()
        self.enc_blocks = nn.ModuleList(blocks)


class QLatentBlockX(nn.Module):
    """Latent block container (reference model.py:210-360)."""
    def __init__(self, width, zdim, enc_width=None, kernel_size=7):
        super().__init__()
        self.in_channels = width
        self.out_channels = width
        concat_ch = (width * 2) if enc_width is None else (width + enc_width)
        enc_width = enc_width or width
        hidden = int(max(width, enc_width) * 0.25)
        use_3x3 = (kernel_size >= 3)
        self.resnet_front = MyConvNeXtBlock(width, kernel_size=kernel_size)
        self.resnet_end = MyConvNeXtBlock(width, kernel_size=kernel_size)
        self.posterior = VDBlock(concat_ch, hidden, zdim, residual=False, use_3x3=use_3x3)
        self.prior = VDBlock(width, hidden, zdim * 2, residual=False, use_3x3=use_3x3, zero_last=True)
        self.z_proj = nn.Sequential(
            common.conv_k3s1(zdim, hidden // 2) if use_3x3 else common.conv_k1s1(zdim, hidden // 2),
            nn.GELU(),
            common.conv_k1s1(hidden // 2, width),
        )
        self.discrete_gaussian = entropy_coding.GaussianConditional(None)
        self.zdim, self.enc_width = zdim, enc_width
        self.is_latent_block = True

    def residual_scaling(self, N):
        self.z_proj[2].weight.data.mul_(math.sqrt(1 / 3 * N))      # (sic) reference model.py:243-244

    def update(self):
        """Scale table 0.1 ... 20 and its CDF tables (reference model.py:317-325)."""
        log_scales = torch.linspace(math.log(0.1), math.log(20), steps=64)
        self.discrete_gaussian.update_scale_table(torch.exp(log_scales))
        self.discrete_gaussian.update()


class TopDownDecoder(nn.Module):
    def __init__(self, blocks):
        super().__init__()
        self.dec_blocks = nn.ModuleList(blocks)
        width = self.dec_blocks[0].in_channels
        self.bias = nn.Parameter(torch.zeros(1, width, 1, 1))
        total_blocks = len([1 for b in self.dec_blocks if hasattr(b, 'residual_scaling')])
        for block in self.dec_blocks:
            if hasattr(block, 'residual_scaling'):
                block.residual_scaling(total_blocks)

    def update(self):
        for block in self.dec_blocks:
            if hasattr(block, 'update'):
                block.update()


class HierarchicalVAE(nn.Module):
    log2_e = math.log2(math.e)
    family = 'qres'

    def __init__(self, config: dict):
        super().__init__()
        self.encoder = BottomUpEncoder(blocks=config.pop('enc_blocks'))
        self.decoder = TopDownDecoder(blocks=config.pop('dec_blocks'))
        self.out_net = config.pop('out_net')
        self.im_shift = float(config['im_shift'])
        self.im_scale = float(config['im_scale'])
        self.max_stride = config['max_stride']
        self.register_buffer('_dummy', torch.zeros(1), persistent=False)
        self._stats_log = dict()
        self._flops_mode = False
        self.compressing = False
        self.num_latents = len([b for b in self.decoder.dec_blocks if getattr(b, 'is_latent_block', False)])
        self.precision = config.get('precision', 'f16x3')        # see qarv/model.py
        self.__dict__['_engine'] = None

    # ------------------------------------------------------------------ engine plumbing (the engine walks
    # model.encoder.enc_blocks / model.dec_blocks / model.bias, the qarv attribute names)
    @property
    def dec_blocks(self):
        return self.decoder.dec_blocks

    @property
    def bias(self):
        return self.decoder.bias

    @property
    def engine(self):
        if self.__dict__.get('_engine') is None:
            from ...engine import QarvEngine
            self.__dict__['_engine'] = QarvEngine(self)
        return self.__dict__['_engine']

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in ('_engine', '_train_path') else copy.deepcopy(v, memo)
        return new

    def _device(self):
        return self._dummy.device

    def _check_image(self, im):
        assert (im.shape[2] % self.max_stride == 0) and (im.shape[3] % self.max_stride == 0)
        assert im.dim() == 4 and im.shape[1] == 3 and not im.requires_grad and im.dtype == torch.float32

    def _lmb(self, n):
        return torch.full((n,), self.out_net.mse_lmb, device=self._device())

    def process_output(self, x: torch.Tensor):
        assert not x.requires_grad
        return x.clone().clamp_(min=-1.0, max=1.0).mul_(0.5).add_(0.5)

    # ------------------------------------------------------------------ forward paths
    def forward(self, im, return_rec=False, noise=None):
        """Rate + lambda * MSE of a batch (reference model.py:517-569): OrderedDict(loss: 0-d tensor, kl, mse
        (= lambda * MSE, as the reference logs it), bppix, psnr[, im_hat]); also fills `_stats_log`.
        `noise` (optional, train mode): per-layer U(-.5, .5) tensors replacing the generator draws."""
        if self._flops_mode:
            raise NotImplementedError('_flops_mode is a profiling hook of the ATen modules; use bench.py')
        im = im.to(self._device())
        self._check_image(im)
        nB, imC, imH, imW = im.shape
        mode = 'train' if self.training else 'eval'
        if self.training and torch.is_grad_enabled():
            return self._forward_train(im, return_rec, noise)
        res = self.engine.run(im, self._lmb(nB), mode=mode, want_elem=False, want_im_hat=return_rec, noise=noise)
        host = res['stats_host']
        ndims = imC * imH * imW
        kls = torch.stack([res['kl_layers'][li].mean(0) / ndims for li in range(self.num_latents)])
        bpdim = kls * self.log2_e
        self._stats_log[f'{mode}_bpdim'] = bpdim.tolist()
        self._stats_log[f'{mode}_bppix'] = (bpdim * imC).tolist()
        stats = OrderedDict()
        stats['loss'] = res['stats'][0]
        stats['kl'] = float(host[1])
        stats[self.out_net.loss_name] = float(host[2]) * self.out_net.mse_lmb      # lambda * MSE | the NLL (lmb = 1)
        stats['bppix'] = float(host[1]) * self.log2_e * imC
        stats['psnr'] = -10 * math.log10(float(host[3]))
        if return_rec:
            stats['im_hat'] = res['im_hat']
        return stats

    @torch.no_grad()
    def forward_stream(self, batches, depth=2):
        """Throughput form of forward() for evaluation loops (an extension; see
        lvae.models.qarv.model.VariableRateLossyVAE.forward_stream): one OrderedDict(loss, kl, mse | nll, bppix, psnr -- Python
        floats) per batch, batch i+1 copied to the device while batch i runs; `_stats_log` holds the last batch's rates."""
        assert not (self.training and torch.is_grad_enabled()), 'forward_stream is an inference path'
        mode = 'train' if self.training else 'eval'
        shapes = []

        def items():
            for im in batches:
                im = im[0] if isinstance(im, (tuple, list)) else im
                self._check_image(im)
                shapes.append(tuple(im.shape))
                yield im, self._lmb(im.shape[0])
        for i, res in enumerate(self.engine.run_stream(items(), mode=mode, depth=depth)):
            host, (nB, imC, imH, imW) = res['stats_host'], shapes[i]
            bpdim = [float(k) / (imC * imH * imW) * self.log2_e for k in res['kl_layers_mean_host']]
            self._stats_log[f'{mode}_bpdim'] = bpdim
            self._stats_log[f'{mode}_bppix'] = [b * imC for b in bpdim]
            yield OrderedDict([('loss', float(host[0])), ('kl', float(host[1])),
                               (self.out_net.loss_name, float(host[2]) * self.out_net.mse_lmb),
                               ('bppix', float(host[1]) * self.log2_e * imC), ('psnr', -10 * math.log10(float(host[3])))])

    @property
    def train_path(self):
        if self.__dict__.get('_train_path') is None:
            from ...training import TrainPath
            self.__dict__['_train_path'] = TrainPath(self)
        return self.__dict__['_train_path']

    def _forward_train(self, im, return_rec=False, noise=None):
        """Training step forward with the autograd graph attached to stats['loss'] (lvae.training): what
        `loss.backward()` of lvae/trainer.py:262-270 differentiates (reference model.py:517-569)."""
        assert 0 <= float(im.min()) <= float(im.max()) <= 1, 'image values must lie in [0, 1]'
        nB, imC, imH, imW = im.shape
        T = self.train_path
        if T.autograph_enabled and noise is None and not return_rec and T.autograph.usable(im, None):
            # forward + backward of this shape as two CUDA-graph replays (lvae.training.AutoGraphedTrain)
            loss, sv, klv = T.autograph(im, self._lmb(nB))
            host, klh = sv.cpu(), klv.cpu()
            ndims = imC * imH * imW
            bpdim = [float(k) / ndims * self.log2_e for k in klh]
            self._stats_log['train_bpdim'] = bpdim
            self._stats_log['train_bppix'] = [b * imC for b in bpdim]
            stats = OrderedDict()
            stats['loss'] = loss
            stats['kl'] = float(host[0])
            stats[self.out_net.loss_name] = float(host[3])
            stats['bppix'] = float(host[0]) * self.log2_e * imC
            stats['psnr'] = -10 * math.log10(float(host[2]))
            return stats
        res = T.objective(im, self._lmb(nB), noise=noise)
        ndims = imC * imH * imW
        kls = torch.stack([k.detach().reshape(nB, -1).sum(1).mean(0) / ndims for k in res['kl']])
        bpdim = kls * self.log2_e
        self._stats_log['train_bpdim'] = bpdim.tolist()
        self._stats_log['train_bppix'] = (bpdim * imC).tolist()
        stats = OrderedDict()
        stats['loss'] = res['loss']
        stats['kl'] = res['kl_mean']
        stats[self.out_net.loss_name] = res['lmb_mse']
        stats['bppix'] = res['kl_mean'] * self.log2_e * imC
        stats['psnr'] = -10 * math.log10(res['im_mse'])
        if return_rec:
            stats['im_hat'] = res['im_hat']
        return stats

    @torch.no_grad()
    def forward_eval(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def forward_get_latents(self, im, noise=None):
        """[dict(z=[B,zdim,h,w], kl=[B,zdim,h,w])] per latent layer (reference model.py:603-609)."""
        im = im.to(self._device())
        self._check_image(im)
        res = self.engine.run(im, self._lmb(im.shape[0]), mode='train' if self.training else 'eval', want_elem=True,
                              noise=noise)
        return [dict(z=z, kl=kl) for z, kl in zip(res['z'], res['kl_elem'])]

    # ------------------------------------------------------------------ sampling
    @torch.no_grad()
    def uncond_sample(self, nhw_repeat, temprature=1.0):
        nB, nH, nW = nhw_repeat
        return self.engine.sample(self._lmb(nB), [None] * self.num_latents, (nB, nH, nW), float(temprature))

    @torch.no_grad()
    def cond_sample(self, latents, nhw_repeat=None, temprature=1.0, paint_box=None):
        assert paint_box is None, 'inpainting (paint_box) is outside the rate-distortion path'
        if nhw_repeat is None:
            nB, _, nH, nW = latents[0].shape
        else:
            nB, nH, nW = nhw_repeat
        return self.engine.sample(self._lmb(nB), list(latents), (nB, nH, nW), float(temprature))

    # ------------------------------------------------------------------ real entropy coding
    @property
    def lossless(self):
        return isinstance(self.out_net, GaussianNLLOutputNet)

    def compress_mode(self, mode=True):
        if mode:
            self.decoder.update()
            if self.lossless:
                self.out_net.update()
                self.engine.invalidate()        # the engine keeps a device copy of the out-net's scale table
        self.compressing = mode

    @torch.no_grad()
    def compress(self, im):
        """[B,3,H,W] in [0,1] -> [strings_layer0, ..., strings_layer11, feature_shape] with strings_layer_i a list
        of B byte strings (reference model.py:649-668)."""
        im = im.to(self._device())
        self._check_image(im)
        nB, _, imH, imW = im.shape
        res = self.engine.run(im, self._lmb(nB), mode='compress', want_elem=False)
        out = list(res['strings'])
        width = self.decoder.dec_blocks[0].in_channels
        out.append((nB, width, imH // self.max_stride, imW // self.max_stride))
        if self.lossless:                       # the image's own residual stream, one string per image (model.py:664-667)
            out.append(res['out_strings'])
        return out

    @torch.no_grad()
    def decompress(self, compressed_object):
        """Inverse of compress -> [B,3,H,W] in [0,1] (reference model.py:670-687)."""
        out_strings = None
        if self.lossless:
            compressed_object, out_strings = compressed_object[:-1], compressed_object[-1]
        nB, _, nH, nW = compressed_object[-1]
        strings = compressed_object[:-1]
        assert len(strings) == self.num_latents, f'decoded={len(strings)}, len={len(compressed_object)}'
        return self.engine.decompress(self._lmb(nB), strings, (nB, nH, nW), out_strings=out_strings)

    @torch.no_grad()
    def compress_file(self, img_path, output_path):
        import torchvision.transforms.functional as tvf
        from PIL import Image
        img = Image.open(img_path)
        img_padded = coding.pad_divisible_by(img, div=self.max_stride)
        im = tvf.to_tensor(img_padded).unsqueeze_(0).to(device=self._device())
        compressed_obj = self.compress(im)
        compressed_obj.append((img.height, img.width))
        with open(output_path, 'wb') as f:
            pickle.dump(compressed_obj, file=f)

    @torch.no_grad()
    def decompress_file(self, bits_path):
        with open(bits_path, 'rb') as f:
            compressed_obj = pickle.load(file=f)
        img_h, img_w = compressed_obj.pop()
        im_hat = self.decompress(compressed_obj)
        return im_hat[:, :, :img_h, :img_w]
