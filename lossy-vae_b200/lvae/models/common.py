"""Building blocks of the bottom-up / top-down networks, as PARAMETER CONTAINERS.

Module and parameter names, shapes and default initialisation mirror the reference
(lvae/models/common.py:8-38,48-66,84-161 and timm's `Mlp`) so that state dicts, optimizer
parameter groups (keyed on '.weight' / '.bias' substrings, lvae/trainer.py:182-194), EMA deep
copies and `torch.manual_seed`-reproducible initial weights are interchangeable with it.
None of these modules computes with ATen: the forward pass of a whole model is compiled by
`lvae.engine` into a sequence of liblvae_b200 kernel launches over NHWC buffers.
"""
import torch
import torch.nn as nn


class ParamConv2d(nn.Conv2d):
    """Holds `weight`/`bias` of a convolution (torch default init); compute lives in liblvae_b200."""
    def forward(self, x):
        raise RuntimeError('ParamConv2d is a parameter container; run the owning model (lvae.engine)')


class ParamLinear(nn.Linear):
    """Holds `weight`/`bias` of a linear layer (torch default init); compute lives in liblvae_b200."""
    def forward(self, x):
        raise RuntimeError('ParamLinear is a parameter container; run the owning model (lvae.engine)')


def get_conv(in_ch, out_ch, kernel_size, stride, padding, zero_bias=True, zero_weights=False):
    conv = ParamConv2d(in_ch, out_ch, kernel_size, stride, padding)
    if zero_bias:
        conv.bias.data.mul_(0.0)
    if zero_weights:
        conv.weight.data.mul_(0.0)
    return conv


def conv_k1s1(in_ch, out_ch, zero_bias=True, zero_weights=False):
    return get_conv(in_ch, out_ch, 1, 1, 0, zero_bias, zero_weights)


def conv_k3s1(in_ch, out_ch, zero_bias=True, zero_weights=False):
    return get_conv(in_ch, out_ch, 3, 1, 1, zero_bias, zero_weights)


def patch_downsample(in_ch, out_ch, rate=2):
    conv = get_conv(in_ch, out_ch, kernel_size=rate, stride=rate, padding=0)
    conv.op_kind = 'down'
    conv.rate = rate
    return conv


class PixelShuffleMarker(nn.Module):
    def __init__(self, rate):
        super().__init__()
        self.rate = rate

    def extra_repr(self):
        return f'upscale_factor={self.rate}'


def patch_upsample(in_ch, out_ch, rate=2):
    seq = nn.Sequential(get_conv(in_ch, out_ch * (rate ** 2), kernel_size=1, stride=1, padding=0),
                        PixelShuffleMarker(rate))
    seq.op_kind = 'up'
    seq.rate = rate
    return seq


class SetKey(nn.Module):
    """Marks the position where an encoder feature is tapped (reference common.py:48-56)."""
    def __init__(self, key):
        super().__init__()
        self.key = key


class CompresionStopFlag(nn.Module):
    """Marks where the top-down pass may stop when only bits are needed (common.py:59-66)."""


class Mlp(nn.Module):
    """timm.layers.mlp.Mlp key set: fc1, fc2."""
    def __init__(self, in_features, hidden_features, out_features):
        super().__init__()
        self.fc1 = ParamLinear(in_features, hidden_features)
        self.fc2 = ParamLinear(hidden_features, out_features)


class ConvNeXtBlockAdaLN(nn.Module):
    """dwconv kxk -> LayerNorm(C) -> AdaLN(lambda embedding) -> Linear -> GELU -> Linear -> layer scale
    -> residual (reference common.py:110-161)."""
    default_embedding_dim = 256

    def __init__(self, dim, embed_dim=None, out_dim=None, kernel_size=7, mlp_ratio=2, residual=True,
                 ls_init_value=1e-6):
        super().__init__()
        assert out_dim is None or out_dim == dim, 'out_dim != dim is not used by any registered model'
        assert residual and ls_init_value >= 0
        pad = (kernel_size - 1) // 2
        self.conv_dw = ParamConv2d(dim, dim, kernel_size=kernel_size, padding=pad, groups=dim)
        embed_dim = embed_dim or self.default_embedding_dim
        # index 1 keeps the reference key `embedding_layer.1.weight` (index 0 is the GELU)
        self.embedding_layer = nn.Sequential(nn.Identity(), ParamLinear(embed_dim, 2 * dim))
        hidden = int(mlp_ratio * dim)
        self.mlp = Mlp(dim, hidden, dim)
        self.gamma = nn.Parameter(torch.full(size=(1, dim, 1, 1), fill_value=1e-6))
        self.dim, self.hidden, self.kernel_size = dim, hidden, kernel_size
        self.requires_embedding = True


class ParamLayerNorm(nn.LayerNorm):
    """Holds the affine parameters of a LayerNorm (eps 1e-6); compute lives in liblvae_b200 (dwln kernel)."""
    def forward(self, x):
        raise RuntimeError('ParamLayerNorm is a parameter container; run the owning model (lvae.engine)')


class ConvNeXtBlockLN(nn.Module):
    """timm ConvNeXtBlock key set (conv_dw, norm, mlp.fc1/fc2, gamma [C]) as used by qresvae's MyConvNeXtBlock:
    dwconv kxk -> affine LayerNorm(C) -> Linear -> GELU -> Linear -> layer scale -> residual
    (reference lvae/models/qresvae/model.py:163-182; timm.models.convnext.ConvNeXtBlock)."""
    def __init__(self, dim, mlp_ratio=2, kernel_size=7, ls_init_value=1e-6):
        super().__init__()
        pad = (kernel_size - 1) // 2
        self.conv_dw = ParamConv2d(dim, dim, kernel_size=kernel_size, padding=pad, groups=dim)
        self.norm = ParamLayerNorm(dim, eps=1e-6)
        hidden = int(mlp_ratio * dim)
        self.mlp = Mlp(dim, hidden, dim)
        self.gamma = nn.Parameter(ls_init_value * torch.ones(dim))
        self.dim, self.hidden, self.kernel_size = dim, hidden, kernel_size
        self.requires_embedding = False


class FeatureExtractorWithEmbedding(nn.Module):
    """Bottom-up path container (reference common.py:84-98)."""
    def __init__(self, blocks):
        super().__init__()
        self.enc_blocks = nn.ModuleList(blocks)


def sinusoidal_frequencies(dim=256, max_period=64):
    """freqs of `sinusoidal_embedding` (common.py:101-107), computed by torch on the host in fp32."""
    exponents = torch.linspace(0, 1, steps=(dim // 2))
    return torch.pow(max_period, -1.0 * exponents)
