"""kitti_motion_compensation_b200 — B200-native per-point LiDAR deskew (motion compensation).

The product is the CUDA shared library built from csrc/ (C ABI: include/kmc_b200.h) and the C++ mirror of the
reference API (include/kitti_motion_compensation/*.hpp).  `capi` is the ctypes plumbing used by tests and bench.py.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
