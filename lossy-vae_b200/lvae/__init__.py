"""lvae -- B200-native drop-in for the hot path of duanzhiihao/lossy-vae.

Same import surface as the reference package (`/root/reference/lvae/__init__.py:1-2`):
`lvae.get_model`, `lvae.known_datasets`; the compute underneath is hand-written sm_100a CUDA in
`liblvae_b200.so`, reached through the C ABI of `include/lvae_b200.h`.
"""
from .paths import known_datasets
from .models.registry import get_model, register_model
from . import models
from . import evaluation

__all__ = ['get_model', 'register_model', 'known_datasets', 'models', 'evaluation']
