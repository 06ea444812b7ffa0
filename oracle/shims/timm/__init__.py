"""Minimal `timm` stand-in (TEST INFRASTRUCTURE ONLY).

`timm` is not installed in this image. The reference (`/root/reference/lvae`) needs a
handful of its classes to import; they are restated here from timm's published
behaviour so the *unmodified* reference can run as the parity oracle in the build
container. Nothing in the product package imports this directory.
"""
__version__ = "0.9.0-shim"
