"""libfacedetection.train_b200 — B200-native (sm_100a) YuNet training / inference hot path.

Drop-in for the mmdet plugin surface of ShiqiYu/libfacedetection.train: the YuNet backbone + TFPN
neck + head forward/backward, SimOTA assignment + multi-task loss and decode + NMS run as
hand-written CUDA kernels in ``libyunet_b200.so`` (C ABI: ``include/yunet_b200.h``), driven from
Python through ctypes.  Importing this package requires the built shared library; using it requires
a CUDA device.  There is deliberately no CPU fallback.
"""
from . import _capi  # noqa: F401  (raises ImportError when libyunet_b200.so is missing)
from .engine import YuNetEngine, ARCHS  # noqa: F401
from . import synthetic  # noqa: F401

__all__ = ['YuNetEngine', 'ARCHS', 'synthetic']
