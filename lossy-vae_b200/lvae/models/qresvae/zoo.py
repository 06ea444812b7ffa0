"""qresvae model zoo (reference: lvae/models/qresvae/zoo.py:9-58): `qres34m`, 34.0 M parameters -- 5-stage ConvNeXt
encoder (widths 192/384/384/384/384, stages closed by a block + patch conv), 12 latent blocks (zdim 16/14/12/10/8)
in the top-down decoder, lambda fixed at construction."""
import torch

from ..registry import register_model
from .. import common
from . import model as qres


def _qres34m_blocks(final_up):
    enc_nums = [6, 6, 6, 4, 2]
    dec_nums = [1, 2, 3, 3, 3]
    z_dims = [16, 14, 12, 10, 8]
    ch = 96
    kernels = [7, 7, 5, 3, 1]
    widths = [ch * 2, ch * 4, ch * 4, ch * 4, ch * 4]
    enc = [common.patch_downsample(3, widths[0], rate=4)]
    for s in range(5):
        enc += [qres.MyConvNeXtBlock(widths[s], kernel_size=kernels[s]) for _ in range(enc_nums[s])]
        if s < 4:
            enc.append(qres.MyConvNeXtPatchDown(widths[s], widths[s + 1]))
    dec = []
    for s in range(5):
        w = widths[4 - s]
        dec += [qres.QLatentBlockX(w, z_dims[s], kernel_size=kernels[4 - s]) for _ in range(dec_nums[s])]
        if s < 4:
            dec.append(common.patch_upsample(w, widths[3 - s], rate=2))
        elif final_up:
            dec.append(common.patch_upsample(w, 3, rate=4))
    return enc, dec, widths[0]


@register_model
def qres34m_lossless(pretrained=False):
    """Lossless variant (reference zoo.py:63-114): the qres34m trunk without the final up-sampler; the decoder's H/4
    feature feeds GaussianNLLOutputNet's mean / log-scale heads."""
    cfg = dict()
    cfg['enc_blocks'], cfg['dec_blocks'], w0 = _qres34m_blocks(final_up=False)
    cfg['out_net'] = qres.GaussianNLLOutputNet(conv_mean=common.patch_upsample(w0, 3, rate=4),
                                               conv_scale=common.patch_upsample(w0, 3, rate=4))
    cfg['im_shift'] = -0.4546259594901961
    cfg['im_scale'] = 3.67572653978347
    cfg['max_stride'] = 64
    model = qres.HierarchicalVAE(cfg)
    if pretrained is True:
        from torch.hub import load_state_dict_from_url
        url = 'https://huggingface.co/duanzh0/my-model-weights/resolve/main/qres34m/qres34m-lossless.pt'
        model.load_state_dict(load_state_dict_from_url(url)['model'])
    elif isinstance(pretrained, str):
        model.load_state_dict(torch.load(pretrained, map_location='cpu')['model'])
    else:
        assert pretrained is False, f'Invalid {pretrained=}'
    return model


@register_model
def qres34m(lmb=32, pretrained=False):
    cfg = dict()
    enc_nums = [6, 6, 6, 4, 2]
    dec_nums = [1, 2, 3, 3, 3]
    z_dims = [16, 14, 12, 10, 8]
    ch = 96
    kernels = [7, 7, 5, 3, 1]
    widths = [ch * 2, ch * 4, ch * 4, ch * 4, ch * 4]
    enc = [common.patch_downsample(3, widths[0], rate=4)]
    for s in range(5):
        enc += [qres.MyConvNeXtBlock(widths[s], kernel_size=kernels[s]) for _ in range(enc_nums[s])]
        if s < 4:
            enc.append(qres.MyConvNeXtPatchDown(widths[s], widths[s + 1]))
    cfg['enc_blocks'] = enc
    dec = []
    for s in range(5):
        w = widths[4 - s]
        dec += [qres.QLatentBlockX(w, z_dims[s], kernel_size=kernels[4 - s]) for _ in range(dec_nums[s])]
        dec.append(common.patch_upsample(w, widths[3 - s], rate=2) if s < 4 else common.patch_upsample(w, 3, rate=4))
    cfg['dec_blocks'] = dec
    cfg['out_net'] = qres.MSEOutputNet(mse_lmb=lmb)
    cfg['im_shift'] = -0.4546259594901961
    cfg['im_scale'] = 3.67572653978347
    cfg['max_stride'] = 64

    model = qres.HierarchicalVAE(cfg)
    if (pretrained is True) and (lmb in {16, 32, 64, 128, 256, 512, 1024, 2048}):
        from torch.hub import load_state_dict_from_url
        url = f'https://huggingface.co/duanzh0/my-model-weights/resolve/main/qres34m/qres34m-lmb{lmb}.pt'
        model.load_state_dict(load_state_dict_from_url(url)['model'])
    elif isinstance(pretrained, str):
        model.load_state_dict(torch.load(pretrained, map_location='cpu')['model'])
    else:
        assert pretrained is False, f'Invalid {pretrained=} and {lmb=}'
    return model
