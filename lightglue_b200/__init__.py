"""lightglue_b200 -- a B200-native (sm_100a) implementation of the LightGlue matcher forward path.

``LightGlue`` is a drop-in for ``lightglue.LightGlue`` (cvg/LightGlue): same constructor, same
``forward({"image0": ..., "image1": ...})`` -> output dict; the math runs in hand-written CUDA
kernels behind the C ABI declared in ``include/lightglue_b200.h``.
"""
from .matcher import LightGlue  # noqa: F401

__all__ = ["LightGlue"]
