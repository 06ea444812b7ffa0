"""`timm.utils` subset the reference harness uses: AverageMeter, unwrap_model, random_seed, ModelEmaV2 (same call
signatures and attribute names as timm 0.9; see ../__init__.py for when this module is active)."""
import copy
import random

import numpy as np
import torch

__all__ = ['AverageMeter', 'unwrap_model', 'random_seed', 'ModelEmaV2']


class AverageMeter:
    """Running value / sum / count / mean of a scalar series (lvae/evaluation.py:25-33 keeps one per logged quantity)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val, self.sum, self.count, self.avg = 0, 0, 0, 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class ModelEmaV2(torch.nn.Module):
    """Exponential moving average of a model's state dict: ema <- decay * ema + (1 - decay) * model, every entry of the
    state dict (parameters and buffers), kept in `.module` (lvae/trainer.py:205-212,374-377)."""

    def __init__(self, model, decay=0.9999, device=None):
        super().__init__()
        self.module = copy.deepcopy(model).eval()
        self.decay, self.device = decay, device
        if device is not None:
            self.module.to(device=device)

    @torch.no_grad()
    def _apply_update(self, model, fn):
        for dst, src in zip(self.module.state_dict().values(), model.state_dict().values()):
            if self.device is not None:
                src = src.to(device=self.device)
            dst.copy_(fn(dst, src))

    def update(self, model):
        d = self.decay
        self._apply_update(model, lambda e, m: d * e + (1.0 - d) * m)

    def set(self, model):
        self._apply_update(model, lambda e, m: m)


def unwrap_model(model):
    """The underlying module of a DistributedDataParallel / DataParallel / ModelEmaV2 wrapper."""
    if isinstance(model, ModelEmaV2):
        return unwrap_model(model.module)
    return model.module if hasattr(model, 'module') else model


def random_seed(seed=42, rank=0):
    torch.manual_seed(seed + rank)
    np.random.seed(seed + rank)
    random.seed(seed + rank)
