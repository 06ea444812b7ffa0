"""Utilities on the coding path (reference: lvae/utils/__init__.py:1)."""
import os as _os

_ref_root = _os.environ.get('LVAE_REFERENCE_ROOT')      # overlay mode, see lvae/__init__.py
if _ref_root and _os.path.isdir(_os.path.join(_ref_root, 'lvae', 'utils')):
    __path__.append(_os.path.join(_ref_root, 'lvae', 'utils'))
    from .general import *      # noqa: F401,F403  (the reference's lvae/utils/__init__.py:1 -- its trainer uses these helpers)
