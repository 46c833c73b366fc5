"""ctypes binding of oracle/_ref/libkmc_ref.so — the reference's OWN hot-path sources (motion_compensation.cpp,
trajectory_interpolation.cpp, lie_algebra.cpp, timestamp_mocking.cpp), compiled unmodified from /root/reference by
`make -C oracle ref` behind the C entry points of oracle/ref_shim.cpp.

TEST INFRASTRUCTURE ONLY (same rule as oracle/binding.py): tests/, smoke() and bench.py's CPU legs may load it, the
product never does.  The library is built in the development container (where /root/reference exists) and travels to
the GPU box as a prebuilt file; `available()` says whether it is there.  `eigen_provider()` says what supplied
namespace Eigen at build time: "eigen3" (real Eigen) or "shim" (this repository's eigen_shim.hpp — the reference's
statements are then compiled as they stand, but the 3x3 product / inverse / polar-rotation arithmetic underneath them is
the shim's).

The reference ABORTS the process on a time outside the interpolation interval (trajectory_interpolation.cpp:9,32);
every wrapper here checks the range first and raises instead of calling into the library.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from oracle.binding import ReferenceWouldAbort, _colmajor, _d, _dp, _fp, _from_colmajor, _ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libkmc_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{_LIB_PATH} is not built (run `make -C oracle ref` where /root/reference exists)")
        _lib = C.CDLL(_LIB_PATH)
        _lib.kmc_ref_eigen_provider.restype = C.c_char_p
        _lib.kmc_ref_fraction_of_scan_completed.restype = C.c_double
        _lib.kmc_ref_fraction_of_scan_completed.argtypes = [_dp]
        _lib.kmc_ref_pseudo_time_stamp.restype = C.c_double
        _lib.kmc_ref_pseudo_time_stamp.argtypes = [_dp, C.c_double, C.c_double]
        _lib.kmc_ref_pose_at_time.argtypes = [C.c_double, _dp, C.c_double, _dp, C.c_double, _dp]
        _lib.kmc_ref_relative_pose_between_times.argtypes = [C.c_double, _dp, C.c_double, _dp, C.c_double, C.c_double, _dp]
        _lib.kmc_ref_motion_compensate_point.argtypes = [C.c_double, _dp, C.c_double, _dp, C.c_double, _dp, C.c_double, _dp]
        _lib.kmc_ref_motion_compensate_frame.argtypes = [_dp, _dp, C.c_int64, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp]
        _lib.kmc_ref_deskew_xyzi_scan.argtypes = [_fp, C.c_int64, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp]
        _lib.kmc_ref_timed_frames.restype = C.c_double
        _lib.kmc_ref_timed_frames.argtypes = [_fp, C.c_int64, C.c_int32, _dp, _dp, _dp, C.c_int32, _dp]
    return _lib


def eigen_provider() -> str:
    return lib().kmc_ref_eigen_provider().decode()


def _mat_fn(name, arg, in_matrix, n_out, out_matrix):
    a = _colmajor(arg) if in_matrix else _d(arg).reshape(-1)
    out = np.empty(n_out * n_out if out_matrix else n_out)
    fn = getattr(lib(), name)
    fn.argtypes = [_dp, _dp]
    fn.restype = None
    fn(_ptr(a), _ptr(out))
    return _from_colmajor(out, n_out) if out_matrix else out


# ---- lie_algebra.cpp:7-103 ------------------------------------------------------------------------------
def hat(phi):
    return _mat_fn("kmc_ref_hat", phi, False, 3, True)


def vee(m):
    return _mat_fn("kmc_ref_vee", m, True, 3, False)


def so3_exp(phi):
    return _mat_fn("kmc_ref_so3_exp", phi, False, 3, True)


def so3_log(R):
    return _mat_fn("kmc_ref_so3_log", R, True, 3, False)


def left_jacobian(phi):
    return _mat_fn("kmc_ref_left_jacobian", phi, False, 3, True)


def inverse_left_jacobian(phi):
    return _mat_fn("kmc_ref_inverse_left_jacobian", phi, False, 3, True)


def se3_exp(xi):
    return _mat_fn("kmc_ref_se3_exp", xi, False, 4, True)


def se3_log(T):
    return _mat_fn("kmc_ref_se3_log", T, True, 6, False)


# ---- trajectory_interpolation.cpp:27-51 -------------------------------------------------------------------
def _check(t1, t2, *times):
    for t in times:
        if not (t1 <= t <= t2):
            raise ReferenceWouldAbort("time outside [t1, t2] (trajectory_interpolation.cpp:32)")


def pose_at_time(t1, P1, t2, P2, t):
    _check(t1, t2, t)
    out = np.empty(16)
    lib().kmc_ref_pose_at_time(t1, _ptr(_colmajor(P1)), t2, _ptr(_colmajor(P2)), t, _ptr(out))
    return _from_colmajor(out, 4)


def relative_pose_between_times(t1, P1, t2, P2, anchor, query):
    _check(t1, t2, anchor, query)
    out = np.empty(16)
    lib().kmc_ref_relative_pose_between_times(t1, _ptr(_colmajor(P1)), t2, _ptr(_colmajor(P2)), anchor, query, _ptr(out))
    return _from_colmajor(out, 4)


# ---- timestamp_mocking.cpp:46-63 ----------------------------------------------------------------------------
def fraction_of_scan_completed(point4):
    return lib().kmc_ref_fraction_of_scan_completed(_ptr(_d(point4)))


def pseudo_time_stamp(point4, start, end):
    return lib().kmc_ref_pseudo_time_stamp(_ptr(_d(point4)), start, end)


# ---- motion_compensation.cpp:9-28 -----------------------------------------------------------------------------
def motion_compensate_point(t1, P1, t2, P2, point_stamp, point4, requested_time):
    _check(t1, t2, point_stamp, requested_time)
    out = np.empty(4)
    lib().kmc_ref_motion_compensate_point(t1, _ptr(_colmajor(P1)), t2, _ptr(_colmajor(P2)), point_stamp, _ptr(_d(point4)),
                                          requested_time, _ptr(out))
    return out


def motion_compensate_frame(cloud_n4, timestamps, T_start, T_end, stamp_start, stamp_end, requested_time):
    cloud = np.asarray(cloud_n4, dtype=np.float64)
    n = cloud.shape[0]
    ts = _d(timestamps)
    _check(stamp_start, stamp_end, requested_time, float(ts.min()) if n else stamp_start, float(ts.max()) if n else stamp_start)
    out = np.empty(4 * n)
    lib().kmc_ref_motion_compensate_frame(_ptr(_colmajor(cloud)), _ptr(ts), n, _ptr(_colmajor(T_start)), _ptr(_colmajor(T_end)),
                                          stamp_start, stamp_end, requested_time, _ptr(out))
    return out.reshape(4, n).T.copy()


def _stamps_in_range(pts, stamp_start, stamp_end):
    """GetPseudoTimeStamps in numpy, to predict the reference's abort before calling it (same formula, same libm)."""
    frac = (np.pi - np.arctan2(pts[:, 1].astype(np.float64), pts[:, 0].astype(np.float64))) / (2.0 * np.pi)
    st = stamp_start + frac * (stamp_end - stamp_start)
    return bool(np.all((st >= stamp_start) & (st <= stamp_end)))


def deskew_xyzi_scan(xyzi_f32, T_start, T_end, stamp_start, stamp_end, requested_time):
    """Loader conversion (data_io.cpp:124-135) + GetPseudoTimeStamps + MotionCompensateFrame, all reference code."""
    pts = np.ascontiguousarray(xyzi_f32, dtype=np.float32).reshape(-1, 4)
    n = pts.shape[0]
    _check(stamp_start, stamp_end, requested_time)
    if not _stamps_in_range(pts, stamp_start, stamp_end):
        raise ReferenceWouldAbort("a pseudo stamp rounds outside [stamp_start, stamp_end]")
    out = np.empty((n, 4))
    lib().kmc_ref_deskew_xyzi_scan(pts.ctypes.data_as(_fp), n, _ptr(_colmajor(T_start)), _ptr(_colmajor(T_end)), stamp_start,
                                   stamp_end, requested_time, _ptr(out))
    return out


def timed_frames(xyzi_f32, points_per_frame, T_start, T_end, stamps3, n_threads):
    """Time the reference's MotionCompensateFrame (+ loader conversion and GetPseudoTimeStamps) over many scans."""
    pts = np.ascontiguousarray(xyzi_f32, dtype=np.float32).reshape(-1)
    n_frames = pts.size // (4 * points_per_frame)
    ts = np.ascontiguousarray(np.stack([np.asarray(T).T.reshape(-1) for T in T_start]), dtype=np.float64)
    te = np.ascontiguousarray(np.stack([np.asarray(T).T.reshape(-1) for T in T_end]), dtype=np.float64)
    st = _d(stamps3).reshape(-1, 3)
    for f in range(n_frames):
        _check(st[f, 0], st[f, 1], st[f, 2])
        if not _stamps_in_range(pts.reshape(-1, 4)[f * points_per_frame:(f + 1) * points_per_frame], st[f, 0], st[f, 1]):
            raise ReferenceWouldAbort(f"frame {f}: a pseudo stamp rounds outside the scan interval")
    chk = C.c_double(0.0)
    sec = lib().kmc_ref_timed_frames(pts.ctypes.data_as(_fp), points_per_frame, n_frames, _ptr(ts), _ptr(te),
                                     _ptr(st.reshape(-1)), n_threads, C.byref(chk))
    return sec, chk.value
