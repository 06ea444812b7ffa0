"""Discretised-Gaussian entropy model state (scale table, quantised CDF tables).

Mirrors the buffer set of the reference's `DiscretizedGaussian(GaussianConditional)`
(lvae/models/entropy_coding.py:52-82 on top of CompressAI's EntropyModel) so upstream checkpoints
load with strict=True: `_offset`, `_quantized_cdf`, `_cdf_length` (int32, empty until `update()`),
`likelihood_lower_bound.bound`, `lower_bound_scale.bound`; `scale_table` is non-persistent.
The float math (quantise, likelihood, index) runs in the fused CUDA latent kernel
(csrc/latent.cu); `update()` builds the CDF tables with the host C coder (csrc/rans.cpp).
"""
import math
import numpy as np
import scipy.stats
import torch
import torch.nn as nn

from .. import _native


class _Bound(nn.Module):
    def __init__(self, bound):
        super().__init__()
        self.register_buffer('bound', torch.Tensor([float(bound)]))


def default_scale_table():
    return torch.exp(torch.linspace(math.log(0.11), math.log(20.0), steps=64))


class DiscretizedGaussian(nn.Module):
    cdf_kind = 'normal'
    tail_mass = 1e-9
    entropy_coder_precision = 16

    def __init__(self, scale_table=None):
        super().__init__()
        self.likelihood_lower_bound = _Bound(1e-9)
        self.register_buffer('_offset', torch.IntTensor())
        self.register_buffer('_quantized_cdf', torch.IntTensor())
        self.register_buffer('_cdf_length', torch.IntTensor())
        if scale_table is None:
            scale_table = default_scale_table()
        assert scale_table.dim() == 1 and scale_table.numel() >= 1 and scale_table.min() > 0
        assert torch.equal(scale_table, torch.sort(scale_table)[0])
        self.register_buffer('scale_table', scale_table, persistent=False)
        self.lower_bound_scale = _Bound(scale_table[0])

    def cdf_ready(self):
        return self._quantized_cdf.numel() > 0

    def update(self):
        """Build the [n_scales, max_len+2] quantised CDF table (CompressAI GaussianConditional.update)."""
        cdf, length, offset = build_cdf_tables(self.scale_table.detach().cpu(), self.tail_mass,
                                               self.entropy_coder_precision)
        dev = self.scale_table.device
        self._quantized_cdf = cdf.to(dev)
        self._cdf_length = length.to(dev)
        self._offset = offset.to(dev)

    def host_tables(self):
        if not self.cdf_ready():
            raise ValueError('Uninitialized CDFs. Run update() first')
        return (np.ascontiguousarray(self._quantized_cdf.cpu().numpy()),
                np.ascontiguousarray(self._cdf_length.cpu().numpy()),
                np.ascontiguousarray(self._offset.cpu().numpy()))


class GaussianConditional(DiscretizedGaussian):
    """CompressAI's stock `GaussianConditional(None)` as qres34m uses it (lvae/models/qresvae/model.py:241,317-325):
    buffer set `_offset, _quantized_cdf, _cdf_length, scale_table (persistent, empty until update_scale_table),
    scale_bound, likelihood_lower_bound.bound, lower_bound_scale.bound`; scale lower bound 0.11; the CDF is
    0.5 * erfc(-x / sqrt 2) (`cdf_kind = 'erfc'`: the fused latent kernel and the table builder follow it)."""
    cdf_kind = 'erfc'

    def __init__(self, scale_table=None, scale_bound=0.11):
        nn.Module.__init__(self)
        assert scale_table is None, 'qres34m constructs the entropy model without a table (set in update())'
        self.likelihood_lower_bound = _Bound(1e-9)
        self.register_buffer('_offset', torch.IntTensor())
        self.register_buffer('_quantized_cdf', torch.IntTensor())
        self.register_buffer('_cdf_length', torch.IntTensor())
        self.register_buffer('scale_table', torch.Tensor())
        self.register_buffer('scale_bound', torch.Tensor([float(scale_bound)]))
        self.lower_bound_scale = _Bound(scale_bound)

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = torch.Tensor(tuple(float(s) for s in scale_table)).to(device)
        self.update()
        return True

    def update(self):
        cdf, length, offset = build_cdf_tables(self.scale_table.detach().cpu(), self.tail_mass,
                                               self.entropy_coder_precision, cdf_kind='erfc')
        dev = self.scale_table.device
        self._quantized_cdf = cdf.to(dev)
        self._cdf_length = length.to(dev)
        self._offset = offset.to(dev)


_table_cache = {}


def build_cdf_tables(scale_table, tail_mass=1e-9, precision=16, cdf_kind='normal'):
    """pmf of the integer offsets around each table scale -> 16-bit quantised CDF rows."""
    key = (scale_table.numpy().tobytes(), tail_mass, precision, cdf_kind)
    if key in _table_cache:
        return tuple(t.clone() for t in _table_cache[key])
    lib = _native.lib()
    multiplier = -scipy.stats.norm.ppf(tail_mass / 2)
    center = torch.ceil(scale_table * multiplier).int()
    length = 2 * center + 1
    max_length = int(length.max())
    samples = torch.abs(torch.arange(max_length).int() - center[:, None]).float()
    sc = scale_table.unsqueeze(1).float()
    if cdf_kind == 'erfc':        # CompressAI GaussianConditional._standardized_cumulative
        def cdf_fn(t):
            return float(0.5) * torch.erfc(float(-(2 ** -0.5)) * t)
    else:                         # td.Normal(0, 1).cdf (lvae/models/entropy_coding.py:77-82)
        cdf_fn = torch.distributions.Normal(0.0, 1.0).cdf
    upper = cdf_fn((0.5 - samples) / sc)
    lower = cdf_fn((-0.5 - samples) / sc)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    cdf = np.zeros((len(length), max_length + 2), dtype=np.int32)
    for i in range(len(length)):
        n = int(length[i])
        prob = np.ascontiguousarray(torch.cat((pmf[i, :n], tail[i])).numpy(), dtype=np.float32)
        row = np.zeros(n + 2, dtype=np.int32)
        _native.check(lib.lvae_pmf_to_quantized_cdf(prob.ctypes.data, n + 1, precision, row.ctypes.data),
                      'pmf_to_quantized_cdf')
        cdf[i, :n + 2] = row
    out = (torch.from_numpy(cdf), (length + 2).int(), (-center).int())
    _table_cache[key] = out
    return tuple(t.clone() for t in out)
