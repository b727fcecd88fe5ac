"""B200-native (sm_100a) train / infer step of the keypoint-interpolation transformer.

Drop-in for the reference's Python surface: ``model.KeypointCompleter``, ``euclidean_loss``,
``augmentation``, ``dataloader`` helpers and the A1 train / eval step, running hand-written CUDA
through the C ABI in ``include/kit.h`` (``libkit_b200.so``).  No CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
