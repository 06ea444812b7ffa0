"""EntropyModel / GaussianConditional restated from CompressAI entropy_models/entropy_models.py."""
import scipy.stats
import torch
import torch.nn as nn

from compressai.ops import LowerBound
from compressai._cxx import pmf_to_quantized_cdf as _pmf_to_quantized_cdf
from compressai import ans


def pmf_to_quantized_cdf(pmf, precision=16):
    cdf = _pmf_to_quantized_cdf(pmf.tolist(), precision)
    return torch.IntTensor(cdf)


class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound=1e-9, entropy_coder=None, entropy_coder_precision=16):
        super().__init__()
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    def quantize(self, inputs, mode, means=None):
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            half = float(0.5)
            noise = torch.empty_like(inputs).uniform_(-half, half)
            return inputs + noise
        outputs = inputs.clone()
        if means is not None:
            outputs -= means
        outputs = torch.round(outputs)
        if mode == "dequantize":
            if means is not None:
                outputs += means
            return outputs
        return outputs.int()

    @staticmethod
    def dequantize(inputs, means=None, dtype=torch.float):
        if means is not None:
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.type(dtype)
        return outputs

    def _check_cdf(self):
        if self._quantized_cdf.numel() == 0 or self._offset.numel() == 0 or self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")

    def compress(self, inputs, indexes, means=None):
        symbols = self.quantize(inputs, "symbols", means)
        self._check_cdf()
        cdfs = self._quantized_cdf.tolist()
        sizes = self._cdf_length.reshape(-1).int().tolist()
        offsets = self._offset.reshape(-1).int().tolist()
        strings = []
        for i in range(symbols.size(0)):
            rv = ans.RansEncoder().encode_with_indexes(
                symbols[i].reshape(-1).int().tolist(), indexes[i].reshape(-1).int().tolist(),
                cdfs, sizes, offsets)
            strings.append(rv)
        return strings

    def decompress(self, strings, indexes, dtype=torch.float, means=None):
        self._check_cdf()
        cdf = self._quantized_cdf
        outputs = cdf.new_empty(indexes.size())
        cdfs = cdf.tolist()
        sizes = self._cdf_length.reshape(-1).int().tolist()
        offsets = self._offset.reshape(-1).int().tolist()
        for i, s in enumerate(strings):
            values = ans.RansDecoder().decode_with_indexes(
                s, indexes[i].reshape(-1).int().tolist(), cdfs, sizes, offsets)
            outputs[i] = torch.tensor(values, device=outputs.device, dtype=outputs.dtype).reshape(outputs[i].size())
        return self.dequantize(outputs, means, dtype)


class GaussianConditional(EntropyModel):
    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if scale_table is not None and len(scale_table) > 0:
            scale_table = torch.Tensor(tuple(float(s) for s in scale_table))
            if scale_bound is None:
                scale_bound = float(scale_table[0])
        else:
            scale_table = torch.Tensor()
        self.register_buffer("scale_table", scale_table)
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)
        self.tail_mass = float(tail_mass)
        self.lower_bound_scale = LowerBound(scale_bound if scale_bound is not None else 0.11)

    @staticmethod
    def _standardized_quantile(quantile):
        return scipy.stats.norm.ppf(quantile)

    def _standardized_cumulative(self, inputs):
        half = float(0.5)
        const = float(-(2 ** -0.5))
        return half * torch.erfc(const * inputs)

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = torch.Tensor(tuple(float(s) for s in scale_table)).to(device)
        self.update()
        return True

    def update(self):
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(self.scale_table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = torch.max(pmf_length).item()
        device = pmf_center.device
        samples = torch.abs(torch.arange(max_length, device=device).int() - pmf_center[:, None])
        samples_scale = self.scale_table.unsqueeze(1)
        samples = samples.float()
        samples_scale = samples_scale.float()
        upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
        lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32, device=device)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            _cdf = pmf_to_quantized_cdf(prob, self.entropy_coder_precision)
            cdf[i, : _cdf.size(0)] = _cdf
        self._quantized_cdf = cdf
        self._offset = -pmf_center
        self._cdf_length = pmf_length + 2

    def _likelihood(self, inputs, scales, means=None):
        half = float(0.5)
        values = inputs - means if means is not None else inputs
        scales = self.lower_bound_scale(scales)
        values = torch.abs(values)
        upper = self._standardized_cumulative((half - values) / scales)
        lower = self._standardized_cumulative((-half - values) / scales)
        return upper - lower

    def forward(self, inputs, scales, means=None, training=None):
        if training is None:
            training = self.training
        outputs = self.quantize(inputs, "noise" if training else "dequantize", means)
        likelihood = self._likelihood(outputs, scales, means)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        return outputs, likelihood

    def build_indexes(self, scales):
        scales = self.lower_bound_scale(scales)
        indexes = scales.new_full(scales.size(), len(self.scale_table) - 1).int()
        for s in self.scale_table[:-1]:
            indexes -= (scales <= s).int()
        return indexes
