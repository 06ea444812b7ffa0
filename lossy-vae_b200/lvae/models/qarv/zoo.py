"""qarv model zoo (reference: lvae/models/qarv/zoo.py:9-99).

`qarv_base`: 5-stage bottom-up encoder (30 ConvNeXt-AdaLN blocks), top-down decoder with 9 latent
blocks + 15 plain blocks, z_dims [32, 32, 96, 8]; 93.4 M parameters.  The factory keeps the
reference's signature and state-dict layout; the compute is liblvae_b200 (see lvae/engine.py).
"""
import torch

from ..registry import register_model
from .. import common
from . import model as qarv


def _stage(width, n, kernel_size, mlp_ratio=2):
    return [common.ConvNeXtBlockAdaLN(width, kernel_size=kernel_size, mlp_ratio=mlp_ratio) for _ in range(n)]


def _latents(n, width, zdim, enc_key, enc_width, kernel_size, mlp_ratio):
    return [qarv.VRLVBlockBase(width, zdim, enc_key=enc_key, enc_width=enc_width,
                               kernel_size=kernel_size, mlp_ratio=mlp_ratio) for _ in range(n)]


@register_model
def qarv_base(lmb_range=(16, 2048), pretrained=False):
    cfg = dict(
        im_shift=-0.4546259594901961, im_scale=3.67572653978347,   # imagenet mean / std
        max_stride=64,
        log_images=['collie64.png', 'gun128.png', 'motor256.png'],
        lmb_range=(float(lmb_range[0]), float(lmb_range[1])),
        lmb_embed_dim=(256, 256), sin_period=64,
    )
    common.ConvNeXtBlockAdaLN.default_embedding_dim = cfg['lmb_embed_dim'][1]
    ch = 128
    e = [192, ch * 3, ch * 4, ch * 4, ch * 4]        # encoder widths at strides 4, 8, 16, 32, 64
    d = [ch * 4, ch * 4, ch * 3, ch * 2, ch * 1]     # decoder widths at strides 64, 32, 16, 8, 4
    z = [32, 32, 96, 8]

    cfg['enc_blocks'] = [
        common.patch_downsample(3, e[0], rate=4),
        *_stage(e[0], 7, 7),
        common.patch_downsample(e[0], e[1]),
        *_stage(e[1], 6, 7), common.SetKey('enc_s8'), *_stage(e[1], 1, 7),
        common.patch_downsample(e[1], e[2]),
        *_stage(e[2], 6, 5), common.SetKey('enc_s16'), *_stage(e[2], 1, 7),
        common.patch_downsample(e[2], e[3]),
        *_stage(e[3], 4, 3), common.SetKey('enc_s32'), *_stage(e[3], 1, 7),
        common.patch_downsample(e[3], e[4]),
        *_stage(e[4], 4, 1), common.SetKey('enc_s64'),
    ]
    cfg['dec_blocks'] = [
        *_latents(1, d[0], z[0], 'enc_s64', e[4], 1, 4),
        *_stage(d[0], 1, 1, 4),
        common.patch_upsample(d[0], d[1], rate=2),
        *_stage(d[1], 1, 3, 3), *_latents(2, d[1], z[1], 'enc_s32', e[3], 3, 3), *_stage(d[1], 1, 3, 3),
        common.patch_upsample(d[1], d[2], rate=2),
        *_stage(d[2], 1, 5, 2), *_latents(3, d[2], z[2], 'enc_s16', e[2], 5, 2), *_stage(d[2], 1, 5, 2),
        common.patch_upsample(d[2], d[3], rate=2),
        *_stage(d[3], 1, 7, 1.75), *_latents(3, d[3], z[3], 'enc_s8', e[1], 7, 1.75),
        common.CompresionStopFlag(),
        *_stage(d[3], 1, 7, 1.75),
        common.patch_upsample(d[3], d[4], rate=2),
        *_stage(d[4], 8, 7, 1.5),
        common.patch_upsample(d[4], 3, rate=4),
    ]
    model = qarv.VariableRateLossyVAE(cfg)

    if pretrained is True:
        from torch.hub import load_state_dict_from_url
        url = 'https://huggingface.co/duanzh0/my-model-weights/resolve/main/qarv_base-2022-dec-12.pt'
        model.load_state_dict(load_state_dict_from_url(url)['model'])
    elif pretrained:   # str or Path
        model.load_state_dict(torch.load(pretrained, map_location='cpu')['model'])
    return model
