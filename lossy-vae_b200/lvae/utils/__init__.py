"""Utilities on the coding path (reference: lvae/utils/__init__.py:1)."""
