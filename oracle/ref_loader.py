"""Import the UNMODIFIED reference `lvae` package: from `/root/reference` in the build container, else from the
byte-identical staged copy `oracle/_ref/reference` (made by the committed recipe oracle/make_ref.py; git-ignored, it
travels to the GPU box).

TEST / MEASUREMENT INFRASTRUCTURE ONLY. Puts `oracle/shims` (timm/compressai stand-ins) and the reference root
on sys.path under a private import so that the reference's own model code is the thing that runs.
Used by `oracle/gen_golden.py` (fixture generation), by the `not gpu` tests that pin `oracle/lvae_oracle.py` against
it, and by `bench.py --impl reference` / its `cpu_baseline` leg (the reference's CPU path timed on the box's host cores).
"""
import importlib
import sys
from pathlib import Path

SHIMS = Path(__file__).resolve().parent / 'shims'
STAGED = Path(__file__).resolve().parent / '_ref' / 'reference'
REFERENCE_ROOT = Path('/root/reference') if (Path('/root/reference') / 'lvae' / '__init__.py').is_file() else STAGED


def available():
    return (REFERENCE_ROOT / 'lvae' / '__init__.py').is_file()


def load_reference():
    """Returns the reference `lvae` package (module object). Evicts any other `lvae` first."""
    if not available():
        raise RuntimeError('reference tree not present')
    for name in list(sys.modules):
        if name == 'lvae' or name.startswith('lvae.'):
            mod = sys.modules[name]
            f = getattr(mod, '__file__', '') or ''
            if not f.startswith(str(REFERENCE_ROOT)):
                del sys.modules[name]
    for p in (str(SHIMS), str(REFERENCE_ROOT)):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, str(REFERENCE_ROOT))
    sys.path.insert(0, str(SHIMS))
    ref = importlib.import_module('lvae')
    assert ref.__file__.startswith(str(REFERENCE_ROOT)), ref.__file__
    return ref


def unload_reference():
    for name in list(sys.modules):
        if name == 'lvae' or name.startswith('lvae.') or name.split('.')[0] in ('timm', 'compressai'):
            del sys.modules[name]
    for p in (str(SHIMS), str(REFERENCE_ROOT)):
        while p in sys.path:
            sys.path.remove(p)
