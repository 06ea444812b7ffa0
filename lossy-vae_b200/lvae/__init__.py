"""lvae -- B200-native drop-in for the hot path of duanzhiihao/lossy-vae.

Same import surface as the reference package (`/root/reference/lvae/__init__.py:1-2`):
`lvae.get_model`, `lvae.known_datasets`; the compute underneath is hand-written sm_100a CUDA in
`liblvae_b200.so`, reached through the C ABI of `include/lvae_b200.h`.
"""
import os as _os

# Overlay mode: with LVAE_REFERENCE_ROOT=/path/to/lossy-vae, sub-modules this package does not have --
# the reference's trainer, datasets, logging utilities: callers of the path, not the path -- are loaded, unmodified, from
# that checkout, while lvae.models / lvae.evaluation / lvae.utils.coding / lvae.paths stay the ones of this package.
_ref_root = _os.environ.get('LVAE_REFERENCE_ROOT')
if _ref_root and _os.path.isdir(_os.path.join(_ref_root, 'lvae')):
    __path__.append(_os.path.join(_ref_root, 'lvae'))

# The reference's harness (lvae/trainer.py:14, lvae/evaluation.py:9, train-var-rate.py:5) imports four helpers from
# `timm.utils`.  With timm installed, that is what they get; on a box without it (this image), the restatement under
# ../compat is put on the path so that the reference's scripts run unchanged.
import importlib.util as _ilu
import sys as _sys
if _ilu.find_spec('timm') is None:
    _sys.path.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'compat'))

from .paths import known_datasets
from .models.registry import get_model, register_model
from . import models
from . import evaluation

__all__ = ['get_model', 'register_model', 'known_datasets', 'models', 'evaluation']
