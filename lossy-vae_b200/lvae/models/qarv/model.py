"""Variable-rate hierarchical VAE (qarv): the reference's model surface over the B200 engine.

Mirrors the public surface of `VariableRateLossyVAE` / `VRLVBlockBase`
(reference lvae/models/qarv/model.py:19-124,169-581): same constructor config, same module /
parameter names (state-dict compatible), same methods (`forward`, `forward_end2end`,
`compress_mode`, `compress`, `decompress`, `compress_file`, `decompress_file`, `self_evaluate`,
`conditional_sample`, `unconditional_sample`, `sample_lmb`, ...), same return types and the same
assertion / ValueError conventions.  The arithmetic runs in liblvae_b200 (hand-written sm_100a
CUDA behind the C ABI of include/lvae_b200.h), sequenced by `lvae.engine.QarvEngine`; there is no
ATen / CPU fallback, so every compute entry point requires the model to live on a CUDA device.
"""
import math
import struct
from collections import OrderedDict, defaultdict
from pathlib import Path

import torch
import torch.nn as nn

from .. import common
from .. import entropy_coding
from ...utils import coding


class VRLVBlockBase(nn.Module):
    """Variable-rate latent-variable block: parameter container (reference model.py:19-124).

    resnet_front -> prior 1x1 -> (posterior0 | posterior1) -> post_merge 1x1 on the K-concat ->
    posterior2 -> posterior 3x3 -> quantise/likelihood -> z_proj 1x1 fuse -> resnet_end.
    """
    default_embedding_dim = 256

    def __init__(self, width, zdim, enc_key, enc_width, embed_dim=None, kernel_size=7, mlp_ratio=2):
        super().__init__()
        self.in_channels = width
        self.out_channels = width
        self.enc_key = enc_key
        self.enc_width = enc_width
        self.zdim = zdim
        block = common.ConvNeXtBlockAdaLN
        embed_dim = embed_dim or self.default_embedding_dim
        self.resnet_front = block(width, embed_dim, kernel_size=kernel_size, mlp_ratio=mlp_ratio)
        self.resnet_end = block(width, embed_dim, kernel_size=kernel_size, mlp_ratio=mlp_ratio)
        self.posterior0 = block(enc_width, embed_dim, kernel_size=kernel_size)
        self.posterior1 = block(width, embed_dim, kernel_size=kernel_size)
        self.posterior2 = block(width, embed_dim, kernel_size=kernel_size)
        self.post_merge = common.conv_k1s1(width + enc_width, width)
        self.posterior = common.conv_k3s1(width, zdim)
        self.z_proj = common.conv_k1s1(zdim, width)
        self.prior = common.conv_k1s1(width, zdim * 2)
        self.discrete_gaussian = entropy_coding.DiscretizedGaussian()
        self.is_latent_block = True

    def update(self):
        self.discrete_gaussian.update()


class VariableRateLossyVAE(nn.Module):
    log2_e = math.log2(math.e)
    MAX_LMB = 8192

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
        self.compressing = False
        self._logging_images = config.get('log_images', [])
        self._flops_mode = False
        # arithmetic of the dense contractions (DESIGN.md "Precision modes"):
        #   'f16x3'   tcgen05, 2 fp16 planes per operand (22 significand bits), 3 MMAs, split main/cross TMEM
        #             accumulators: fp32-class, the parity mode (default)
        #   'bf16x6'  tcgen05, 3 bf16 planes per operand, 6 MMAs: fp32-class, the first parity mode (kept: wider range)
        #   'bf16x3'  tcgen05, 2 planes, 3 MMAs (~2^-17 per product)      'bf16'  tcgen05 single pass (fast, non-parity)
        #   'fp32'    fp32 FFMA on CUDA cores
        self.precision = config.get('precision', 'f16x3')
        self.__dict__['_engine'] = None   # not a submodule / not deep-copied state

    # ------------------------------------------------------------------ engine plumbing
    @property
    def engine(self):
        if self.__dict__.get('_engine') is None:
            from ...engine import QarvEngine
            self.__dict__['_engine'] = QarvEngine(self)
        return self.__dict__['_engine']

    def __deepcopy__(self, memo):
        # EMA wrappers deep-copy the model (lvae/trainer.py:311); the engine (device buffers,
        # ctypes descriptors) is rebuilt lazily by the copy.
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in ('_engine', '_train_path') else copy.deepcopy(v, memo)
        return new

    def _device(self):
        return self._dummy.device

    # ------------------------------------------------------------------ small helpers (reference surface)
    def preprocess_input(self, im: torch.Tensor):
        assert (im.shape[2] % self.max_stride == 0) and (im.shape[3] % self.max_stride == 0)
        assert (im.dim() == 4) and (0 <= im.min() <= im.max() <= 1) and not im.requires_grad
        return im

    def sample_lmb(self, n):
        low, high = self.lmb_range
        p = 3.0
        low, high = math.pow(low, 1 / p), math.pow(high, 1 / p)
        transformed = low + (high - low) * torch.rand(n, device=self._device())
        return torch.pow(transformed, exponent=p)

    def expand_to_tensor(self, input_, n):
        assert isinstance(input_, (torch.Tensor, float, int)), f'{type(input_)=}'
        if isinstance(input_, torch.Tensor) and (input_.numel() == 1):
            input_ = input_.item()
        if isinstance(input_, (float, int)):
            input_ = torch.full(size=(n,), fill_value=float(input_), device=self._device())
        assert input_.shape == (n,), f'{input_=}, {input_.shape=}'
        return input_

    def _check_image(self, im):
        # shape / grad checks here; the [0,1] range check (reference model.py:220) is evaluated on the device
        # copy by the engine and raised as AssertionError with the results (one sync instead of two)
        assert im.dim() == 4 and im.shape[1] == 3, f'expected [B,3,H,W], got {tuple(im.shape)}'
        assert (im.shape[2] % self.max_stride == 0) and (im.shape[3] % self.max_stride == 0)
        assert not im.requires_grad and im.dtype == torch.float32

    # ------------------------------------------------------------------ forward paths
    def forward_end2end(self, im: torch.Tensor, lmb: torch.Tensor, mode='trainval', get_latent=False):
        """Returns (x_hat [B,3,H,W] in about (-1,1), [stats_i] for the latent layers) -- or only the
        list when mode == 'compress' (reference model.py:294-315).  stats_i holds 'kl' [B,z,h,w]
        (nats) and, with get_latent, 'z'."""
        if mode not in ('trainval', 'compress'):
            raise ValueError(f'Unknown mode={mode}')
        self._check_image(im)
        lmb = self.expand_to_tensor(lmb, n=im.shape[0])
        if mode == 'compress':
            res = self.engine.run(im, lmb, mode='compress', want_elem=False)
            return [dict(strings=s) for s in res['strings']]
        emode = 'train' if self.training else 'eval'
        res = self.engine.run(im, lmb, mode=emode, want_elem=True)
        stats = []
        for li in range(self.num_latents):
            st = dict(kl=res['kl_elem'][li])
            if get_latent:
                st['z'] = res['z'][li]
            stats.append(st)
        return res['x_hat'], stats

    def forward(self, batch, lmb=None, return_rec=False):
        """Rate-distortion objective of a batch (reference model.py:317-363).

        Returns OrderedDict(loss: 0-d tensor, bppix, mse, psnr: float[, im_hat])."""
        im = batch[0] if isinstance(batch, (tuple, list)) else batch
        nB, imC, imH, imW = im.shape
        if self._flops_mode:
            raise NotImplementedError('_flops_mode is a profiling hook of the ATen modules; use bench.py')
        if lmb is None:
            lmb = self.sample_lmb(n=nB)
        assert isinstance(lmb, torch.Tensor) and lmb.shape == (nB,)
        self._check_image(im)
        if self.training and torch.is_grad_enabled():
            return self._forward_train(im, lmb, return_rec)
        emode = 'train' if self.training else 'eval'
        res = self.engine.run(im, lmb.to(self._device(), torch.float32), mode=emode, want_elem=False,
                              want_im_hat=return_rec)
        host = res['stats_host']              # one D2H copy: [loss, kl, mse, im_mse, ...]
        stats = OrderedDict()
        stats['loss'] = res['stats'][0]       # 0-d device tensor
        stats['bppix'] = float(host[1]) * self.log2_e * imC
        stats[self.distortion_name] = float(host[2])
        stats['psnr'] = -10 * math.log10(float(host[3]))
        if return_rec:
            stats['im_hat'] = res['im_hat']
        return stats

    @torch.no_grad()
    def forward_stream(self, batches, lmb=None, depth=2):
        """Throughput form of forward() for evaluation loops (an extension: the reference has one synchronous call per
        batch).  `batches` yields image batches ([B,3,H,W] fp32 in [0,1]; pinned host memory for asynchronous copies --
        or (im, label) pairs); one OrderedDict(loss, bppix, mse, psnr -- all Python floats) per batch comes back, in
        order.  Batch i+1 is copied to the device while batch i runs (lvae.engine.run_stream); values equal forward()'s.
        lmb: None (sampled per batch), a float, or a [B] tensor used for every batch."""
        assert not (self.training and torch.is_grad_enabled()), 'forward_stream is an inference path'
        shapes = []

        def items():
            for batch in batches:
                im = batch[0] if isinstance(batch, (tuple, list)) else batch
                self._check_image(im)
                nB = im.shape[0]
                l = self.sample_lmb(n=nB) if lmb is None else self.expand_to_tensor(lmb, n=nB)
                shapes.append(im.shape[1])
                yield im, l.to(self._device(), torch.float32)
        for i, res in enumerate(self.engine.run_stream(items(), mode='train' if self.training else 'eval', depth=depth)):
            host, imC = res['stats_host'], shapes[i]
            yield OrderedDict([('loss', float(host[0])), ('bppix', float(host[1]) * self.log2_e * imC),
                               (self.distortion_name, float(host[2])), ('psnr', -10 * math.log10(float(host[3])))])

    @property
    def train_path(self):
        if self.__dict__.get('_train_path') is None:
            from ...training import TrainPath
            self.__dict__['_train_path'] = TrainPath(self)
        return self.__dict__['_train_path']

    def _forward_train(self, im, lmb, return_rec=False, noise=None):
        """Training step forward with the autograd graph attached to stats['loss'] (lvae.training): what
        `loss.backward()` of lvae/trainer.py:262-270 differentiates."""
        im = im.to(self._device())
        assert 0 <= float(im.min()) <= float(im.max()) <= 1, 'image values must lie in [0, 1]'
        lmb = lmb.to(self._device(), torch.float32)
        T = self.train_path
        if T.autograph_enabled and noise is None and not return_rec and T.autograph.usable(im, lmb):
            # forward + backward of this shape as two CUDA-graph replays (lvae.training.AutoGraphedTrain)
            loss, sv, _ = T.autograph(im, lmb)
            host = sv.cpu()
            stats = OrderedDict()
            stats['loss'] = loss
            stats['bppix'] = float(host[0]) * self.log2_e * im.shape[1]
            stats[self.distortion_name] = float(host[1])
            stats['psnr'] = -10 * math.log10(float(host[2]))
            return stats
        res = T.objective(im, lmb, noise=noise)
        stats = OrderedDict()
        stats['loss'] = res['loss']
        stats['bppix'] = res['kl_mean'] * self.log2_e * im.shape[1]
        stats[self.distortion_name] = res['mse']
        stats['psnr'] = -10 * math.log10(res['im_mse'])
        if return_rec:
            stats['im_hat'] = res['im_hat']
        return stats

    def process_output(self, x: torch.Tensor):
        assert not x.requires_grad
        return x.clone().clamp_(min=-1.0, max=1.0).mul_(0.5).add_(0.5)

    # ------------------------------------------------------------------ sampling
    def conditional_sample(self, lmb, latents, emb=None, bhw_repeat=None, t=1.0):
        """Top-down pass with given latents (or prior samples where None) -> images in [0,1]
        (reference model.py:365-398)."""
        if latents[0] is None:
            assert bhw_repeat is not None, 'bhw_repeat should be provided'
            nB, nH, nW = bhw_repeat
        else:
            assert len(latents) == self.num_latents
            nB, _, nH, nW = latents[0].shape
        assert emb is None, 'pre-computed embeddings are not part of the engine interface'
        lmb = self.expand_to_tensor(lmb, n=nB)
        return self.engine.sample(lmb, latents, (nB, nH, nW), float(t))

    def unconditional_sample(self, lmb, bhw_repeat, t=1.0):
        return self.conditional_sample(lmb, [None] * self.num_latents, bhw_repeat=bhw_repeat, t=t)

    @torch.no_grad()
    def study(self, save_dir, **kwargs):
        import torchvision as tv
        import torchvision.transforms.functional as tvf
        from PIL import Image
        save_dir = Path(save_dir)
        save_dir.mkdir(parents=False, exist_ok=True)
        lmb = self.expand_to_tensor(self.default_lmb, n=1)
        for k in [1, 2]:
            num = 6
            im_samples = self.unconditional_sample(self.default_lmb, bhw_repeat=(num, k, k))
            tv.utils.save_image(im_samples, fp=save_dir / f'samples_k{k}_hw{im_samples.shape[2]}.png',
                                nrow=math.ceil(num ** 0.5))
        for imname in self._logging_images:
            impath = Path('images') / imname
            if not impath.is_file():
                continue
            im = tvf.to_tensor(Image.open(impath)).unsqueeze_(0).to(device=self._device())
            x_hat, _ = self.forward_end2end(im, lmb=lmb)
            tv.utils.save_image(torch.cat([im, self.process_output(x_hat)], dim=0), fp=save_dir / imname)

    # ------------------------------------------------------------------ self evaluation (estimated rate)
    @torch.no_grad()
    def _self_evaluate(self, img_paths, lmb: float, pbar=False, log_dir=None):
        import torchvision.transforms.functional as tvf
        import torch.nn.functional as tnf
        from PIL import Image
        was_training = self.training
        self.eval()
        acc = defaultdict(float)
        ch_stats = defaultdict(list)
        for impath in img_paths:
            img = Image.open(impath)
            imgh, imgw = img.height, img.width
            im = tvf.to_tensor(coding.pad_divisible_by(img, div=self.max_stride)).unsqueeze_(0).to(self._device())
            x_hat, stats_all = self.forward_end2end(im, lmb=self.expand_to_tensor(lmb, n=1))
            x_hat = x_hat[:, :, :imgh, :imgw]
            _, imC, imH, imW = im.shape
            kl = sum(st['kl'].sum(dim=(1, 2, 3)) for st in stats_all).mean(0) / (imC * imgh * imgw)
            real = tvf.to_tensor(img)
            x_target = real.unsqueeze(0).to(self._device()).add(-0.5).mul(2.0)
            distortion = tnf.mse_loss(x_hat, x_target, reduction='none').mean(dim=(1, 2, 3)).item()
            fake = self.process_output(x_hat).cpu().squeeze(0)
            mse = tnf.mse_loss(real, fake, reduction='mean').item()
            acc['count'] += 1
            acc['loss'] += float(kl.item() + lmb * distortion)
            acc['bpp'] += kl.item() * self.log2_e * imC
            acc['psnr'] += float(-10 * math.log10(mse))
            if log_dir is not None:
                for i, st in enumerate(stats_all):
                    ch_stats[i].append(st['kl'].sum(dim=(2, 3)).mean(0).cpu() / (imH * imW) * self.log2_e)
        self.train(was_training)
        count = acc.pop('count')
        avg = {k: v / count for k, v in acc.items()}
        avg['lambda'] = lmb
        if log_dir is not None:
            self._log_channel_stats({k: torch.stack(v).mean(0) for k, v in ch_stats.items()}, Path(log_dir), lmb)
        return avg

    @staticmethod
    def _log_channel_stats(channel_bpp, log_dir, lmb):
        msg = '=' * 64 + '\n---- row: latent blocks, colums: channels, avg over images ----\n'
        keys = sorted(channel_bpp.keys())
        for k in keys:
            msg += ''.join(f'{a:<7.4f} ' for a in channel_bpp[k].tolist()) + '\n'
        msg += '---- colums: latent blocks, avg over images ----\n'
        msg += ''.join(f'{channel_bpp[k].sum().item():<7.4f} ' for k in keys) + '\n'
        for name in (f'channel-bppix-lmb{round(lmb)}.txt', 'all_lmb_channel_stats.txt'):
            with open(log_dir / name, mode='a') as f:
                print(msg, file=f)

    @torch.no_grad()
    def self_evaluate(self, img_dir, lmb_range=None, steps=8, log_dir=None):
        """dict with lists 'loss', 'bpp', 'psnr', 'lambda' over `steps` log-spaced lambdas
        (reference model.py:491-507)."""
        img_paths = sorted(Path(img_dir).rglob('*.*'))
        start, end = self.lmb_range if (lmb_range is None) else lmb_range
        lambdas = torch.linspace(math.log(start), math.log(end), steps=steps).exp()
        if log_dir is not None:
            (Path(log_dir) / 'all_lmb_channel_stats.txt').unlink(missing_ok=True)
        out = defaultdict(list)
        for lmb in lambdas.tolist():
            for k, v in self._self_evaluate(img_paths, lmb, log_dir=log_dir).items():
                out[k].append(v)
        return out

    # ------------------------------------------------------------------ real entropy coding
    def compress_mode(self, mode=True):
        if mode:
            for block in self.dec_blocks:
                if hasattr(block, 'update'):
                    block.update()
        self.compressing = mode

    @torch.no_grad()
    def compress(self, im, lmb=None):
        """One image [1,3,H,W] in [0,1] -> bytes: 'f' lambda | '3H' (nB, H/64, W/64) | packed strings
        (reference model.py:516-529)."""
        lmb = lmb or self.default_lmb
        assert im.shape[0] == 1, f'Right now only support a single image, got {im.shape=}'
        results = self.forward_end2end(im, lmb=lmb, mode='compress')
        assert len(results) == self.num_latents
        string = coding.pack_byte_strings([res['strings'][0] for res in results])
        nB, _, imH, imW = im.shape
        return struct.pack('f', lmb) + struct.pack('3H', nB, imH // self.max_stride, imW // self.max_stride) + string

    @torch.no_grad()
    def decompress(self, string):
        """bytes -> reconstruction [1,3,H,W] in [0,1] (reference model.py:531-557)."""
        lmb = struct.unpack('f', string[:4])[0]
        nB, nH, nW = struct.unpack('3H', string[4:10])
        all_lv_strings = coding.unpack_byte_string(string[10:])
        assert len(all_lv_strings) == self.num_latents, f'str_i={self.num_latents}, len={len(all_lv_strings)}'
        lmb = self.expand_to_tensor(lmb, n=nB)
        return self.engine.decompress(lmb, all_lv_strings, (nB, nH, nW))

    # ---- batched variants (an extension: SURVEY 8(f)-2; the reference codes one image per call).  Every blob is a
    #      standard single-image bit stream, byte-identical to what compress() gives for that image (the kernels are
    #      batch-invariant), so blobs can be mixed freely between the two APIs.
    @torch.no_grad()
    def compress_batch(self, im, lmb=None):
        """[B,3,H,W] in [0,1] -> list of B byte strings.  lmb: None (default_lmb), a float, or a [B] tensor."""
        nB, _, imH, imW = im.shape
        lmb = self.default_lmb if lmb is None else lmb
        lmb_t = self.expand_to_tensor(lmb, n=nB)
        results = self.forward_end2end(im, lmb=lmb_t, mode='compress')
        assert len(results) == self.num_latents
        lmbs = lmb_t.tolist()
        head = struct.pack('3H', 1, imH // self.max_stride, imW // self.max_stride)
        return [struct.pack('f', lmbs[b]) + head + coding.pack_byte_strings([res['strings'][b] for res in results])
                for b in range(nB)]

    @torch.no_grad()
    def decompress_batch(self, blobs):
        """list of B byte strings of same-size images -> [B,3,H,W] in [0,1]."""
        lmbs, shapes, layers = [], set(), []
        for blob in blobs:
            lmbs.append(struct.unpack('f', blob[:4])[0])
            nB, nH, nW = struct.unpack('3H', blob[4:10])
            assert nB == 1
            shapes.add((nH, nW))
            layers.append(coding.unpack_byte_string(blob[10:]))
        assert len(shapes) == 1, f'decompress_batch needs images of one size, got {sorted(shapes)}'
        nH, nW = shapes.pop()
        strings = [[layers[b][li] for b in range(len(blobs))] for li in range(self.num_latents)]
        lmb = torch.tensor(lmbs, device=self._device())
        return self.engine.decompress(lmb, strings, (len(blobs), nH, nW))

    @torch.no_grad()
    def compress_file(self, img_path, output_path, lmb=None):
        import torchvision.transforms.functional as tvf
        from PIL import Image
        img = Image.open(img_path)
        img_padded = coding.pad_divisible_by(img, div=self.max_stride)
        im = tvf.to_tensor(img_padded).unsqueeze_(0).to(device=self._device())
        body = self.compress(im, lmb=lmb)
        with open(output_path, 'wb') as f:
            f.write(struct.pack('2H', img.height, img.width) + body)

    @torch.no_grad()
    def decompress_file(self, bits_path):
        with open(bits_path, 'rb') as f:
            header = f.read(4)
            body = f.read()
        img_h, img_w = struct.unpack('2H', header)
        return self.decompress(body)[:, :, :img_h, :img_w]
