"""qresvae: fixed-rate hierarchical VAE family (reference lvae/models/qresvae)."""
from . import zoo
