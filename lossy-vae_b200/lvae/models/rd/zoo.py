"""rd model zoo (reference: lvae/models/rd/zoo.py:9-77): `rd_model_base`, 186.7 M parameters -- 5-stage
ConvNeXt-AdaLN encoder (widths 256/512/640/768/768, every stage closed by a block + patch conv), 15 latent
blocks (zdim 32) in the top-down decoder."""
import torch

from ..registry import register_model
from .. import common
from . import model as lib


@register_model
def rd_model_base(lmb_range=(4, 2048), pretrained=False):
    cfg = dict(lmb_range=(float(lmb_range[0]), float(lmb_range[1])), lmb_embed_dim=(256, 256), sin_period=64,
               im_shift=-0.4546259594901961, im_scale=3.67572653978347, max_stride=64,
               log_images=['collie64.png', 'gun128.png', 'motor256.png'])
    emb = cfg['lmb_embed_dim'][1]
    e = [256, 512, 640, 768, 768]
    d = [768, 768, 640, 512, 256]
    z = [32, 32, 32, 32, 32]
    blk = common.ConvNeXtBlockAdaLN

    enc = [common.patch_downsample(3, e[0], rate=4)]
    for stage, n in enumerate([6, 6, 6, 4]):
        enc += [blk(e[stage], emb) for _ in range(n)]
        enc.append(lib.ConvNeXtAdaLNPatchDown(e[stage], e[min(stage + 1, 3)] if stage < 3 else e[3], embed_dim=emb))
    enc += [blk(e[3], emb) for _ in range(4)]
    cfg['enc_blocks'] = enc

    dec = []
    for stage in range(5):
        dec += [lib.LatentVariableBlock(d[stage], z[stage], emb, enc_width=e[4 - stage]) for _ in range(stage + 1)]
        dec.append(common.patch_upsample(d[stage], d[stage + 1], rate=2) if stage < 4 else common.patch_upsample(d[4], 3, rate=4))
    cfg['dec_blocks'] = dec

    model = lib.VariableRateLossyVAE(cfg)
    if pretrained is True:
        from torch.hub import load_state_dict_from_url
        url = 'https://huggingface.co/duanzh0/my-model-weights/resolve/main/rd_model_base-200k-feb14-2023.pt'
        model.load_state_dict(load_state_dict_from_url(url)['model'])
    elif isinstance(pretrained, str):
        model.load_state_dict(torch.load(pretrained, map_location='cpu')['model'])
    return model
