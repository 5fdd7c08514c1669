"""rehrseg_b200 -- B200-native (sm_100a) implementation of the REHRSeg dense-3D-conv hot path.

Public surface (mirrors the reference's seams):
  seg_model.SegModel / seg_model.convert      <- models/seg_model.py:153-210
  functional.*                                 autograd.Function wrappers over the C-ABI (include/rehrseg_b200.h)
The CUDA library `librehrseg_b200.so` is loaded lazily on first use and there is no CPU fallback.
"""
from ._lib import RehrError, declared_symbols, lib  # noqa: F401

__version__ = "0.1.0"
