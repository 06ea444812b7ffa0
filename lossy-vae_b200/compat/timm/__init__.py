"""Minimal stand-in for the `timm` package, used ONLY when the real one is not installed (lvae/__init__.py adds this
directory to sys.path in that case).  The reference's harness imports four helpers from `timm.utils`
(lvae/trainer.py:14, lvae/evaluation.py:9, train-var-rate.py:5); they are callers of the rate-distortion path, not part
of it, and are restated here so that the reference's scripts run unchanged on a box without timm."""
from . import utils  # noqa: F401

__version__ = '0.0-lvae-b200-compat'
