"""b200timm: B200 (sm_100a) kernels behind timm's ViT / ConvNeXt hot path.

Layout
  _lib.py     ctypes binding of libb200timm.so (C ABI, include/b200timm.h)
  ops.py      tensor-level wrappers of the raw kernels
"""
from ._lib import B200Error, load as load_library  # noqa: F401

__version__ = "0.1.0"
