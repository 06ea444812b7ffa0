"""Minimal `compressai` stand-in (TEST INFRASTRUCTURE ONLY).

CompressAI is an un-vendored, un-pinned dependency of the reference and is not installed
in this image. The classes the reference's hot path touches are restated here from
CompressAI's published behaviour (entropy_models/entropy_models.py, ops/bound_ops.py,
cpp_exts/ops/ops.cpp, cpp_exts/rans/rans_interface.cpp + ryg_rans rans64.h) so the
unmodified reference can run as the parity oracle. Bitstream byte-compatibility with a real
CompressAI build cannot be checked here: "parity unpinned" for rANS bytes and CDF tables.
Nothing in the product package imports this directory.
"""
__version__ = "shim"
