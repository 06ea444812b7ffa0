"""timm.models.convnext.ConvNeXtBlock restated (timm 0.8/0.9 attribute names:
conv_dw, norm, mlp, gamma, use_conv_mlp, drop_path), channels-last MLP variant only."""
import torch
import torch.nn as nn
from timm.layers.mlp import Mlp


class ConvNeXtBlock(nn.Module):
    def __init__(self, in_chs, out_chs=None, kernel_size=7, stride=1, dilation=1, mlp_ratio=4,
                 conv_mlp=False, conv_bias=True, ls_init_value=1e-6, act_layer='gelu',
                 norm_layer=None, drop_path=0.0):
        super().__init__()
        out_chs = out_chs or in_chs
        assert not conv_mlp and stride == 1 and dilation == 1 and out_chs == in_chs
        self.use_conv_mlp = False
        pad = (kernel_size - 1) // 2
        self.conv_dw = nn.Conv2d(in_chs, out_chs, kernel_size=kernel_size, padding=pad,
                                 groups=in_chs, bias=conv_bias)
        self.norm = nn.LayerNorm(out_chs, eps=1e-6)
        self.mlp = Mlp(out_chs, int(mlp_ratio * out_chs), act_layer=nn.GELU)
        self.gamma = nn.Parameter(ls_init_value * torch.ones(out_chs)) if ls_init_value > 0 else None
        self.drop_path = nn.Identity()

    def forward(self, x):
        shortcut = x
        x = self.conv_dw(x)
        x = x.permute(0, 2, 3, 1)
        x = self.norm(x)
        x = self.mlp(x)
        x = x.permute(0, 3, 1, 2)
        if self.gamma is not None:
            x = x.mul(self.gamma.reshape(1, -1, 1, 1))
        x = self.drop_path(x) + shortcut
        return x
