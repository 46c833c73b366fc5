"""ctypes binding of oracle/libkmc_oracle.so (the CPU restatement of the reference deskew path).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (kitti_motion_compensation_b200/) never imports this module.

All matrices cross this boundary column-major (Eigen's default, `Affine3d::matrix().data()`); the helpers
here take / return ordinary numpy row-major arrays and transpose at the edge.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkmc_oracle.so")

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe in oracle/Makefile (g++ -O3, the reference's flags)."""
    src = os.path.join(_HERE, "kmc_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "libkmc_oracle.so"] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.kmc_oracle_fraction_of_scan_completed.restype = C.c_double
        _lib.kmc_oracle_pseudo_time_stamp.restype = C.c_double
        _lib.kmc_oracle_pseudo_time_stamp.argtypes = [_dp, C.c_double, C.c_double]
        _lib.kmc_oracle_pseudo_time_stamps.argtypes = [_dp, C.c_int64, C.c_double, C.c_double, _dp]
        _lib.kmc_oracle_pose_at_time.argtypes = [C.c_double, _dp, C.c_double, _dp, C.c_double, _dp]
        _lib.kmc_oracle_relative_pose_between_times.argtypes = [C.c_double, _dp, C.c_double, _dp, C.c_double,
                                                                C.c_double, _dp]
        _lib.kmc_oracle_oxts_to_pose.argtypes = [_dp, C.c_double, _dp]
        _lib.kmc_oracle_interpolate_trajectory.argtypes = [_dp, _dp, C.c_double, _dp]
        _lib.kmc_oracle_make_frame_poses.argtypes = [_dp, _dp, _dp, C.c_double, C.c_double, _dp, _dp]
        _lib.kmc_oracle_motion_compensate_point.argtypes = [C.c_double, _dp, C.c_double, _dp, C.c_double, _dp,
                                                            C.c_double, _dp]
        _lib.kmc_oracle_motion_compensate_frame.argtypes = [_dp, _dp, C.c_int64, _dp, _dp, C.c_double, C.c_double,
                                                            C.c_double, _dp]
        _lib.kmc_oracle_deskew_xyzi_scan.argtypes = [_fp, C.c_int64, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp]
        _lib.kmc_oracle_timed_frames.restype = C.c_double
        _lib.kmc_oracle_timed_frames.argtypes = [_fp, C.c_int64, C.c_int32, _dp, _dp, _dp, C.c_int32, _dp]
        _lib.kmc_oracle_hardware_threads.restype = C.c_int
    return _lib


def _d(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def _colmajor(m) -> np.ndarray:
    """row-major numpy matrix -> flat column-major buffer"""
    return _d(np.asarray(m, dtype=np.float64).T).reshape(-1)


def _from_colmajor(buf: np.ndarray, n: int) -> np.ndarray:
    return buf.reshape(n, n).T.copy()


def _mat_fn(name: str, arg, n_in_matrix: int | None, n_out: int, out_matrix: bool):
    a = _colmajor(arg) if n_in_matrix else _d(arg).reshape(-1)
    out = np.empty(n_out * n_out if out_matrix else n_out, dtype=np.float64)
    fn = getattr(lib(), name)
    fn.argtypes = [_dp, _dp]
    fn.restype = None
    fn(_ptr(a), _ptr(out))
    return _from_colmajor(out, n_out) if out_matrix else out


# ---- lie_algebra.cpp ---------------------------------------------------------------------------------
def hat(phi):
    return _mat_fn("kmc_oracle_hat", phi, None, 3, True)


def vee(m):
    return _mat_fn("kmc_oracle_vee", m, 3, 3, False)


def so3_exp(phi):
    return _mat_fn("kmc_oracle_so3_exp", phi, None, 3, True)


def so3_log(R):
    return _mat_fn("kmc_oracle_so3_log", R, 3, 3, False)


def left_jacobian(phi):
    return _mat_fn("kmc_oracle_left_jacobian", phi, None, 3, True)


def inverse_left_jacobian(phi):
    return _mat_fn("kmc_oracle_inverse_left_jacobian", phi, None, 3, True)


def se3_exp(xi):
    return _mat_fn("kmc_oracle_se3_exp", xi, None, 4, True)


def se3_log(T):
    return _mat_fn("kmc_oracle_se3_log", T, 4, 6, False)


def polar_rotation(L):
    return _mat_fn("kmc_oracle_polar_rotation", L, 3, 3, True)


def affine_inverse(T):
    return _mat_fn("kmc_oracle_affine_inverse", T, 4, 4, True)


# ---- trajectory_interpolation.cpp ----------------------------------------------------------------------
class ReferenceWouldAbort(AssertionError):
    """The reference asserts (aborts, even in Release) on a time outside the interpolation interval."""


def pose_at_time(t1, P1, t2, P2, t, allow_abort=False):
    out = np.empty(16)
    rc = lib().kmc_oracle_pose_at_time(t1, _ptr(_colmajor(P1)), t2, _ptr(_colmajor(P2)), t, _ptr(out))
    if rc and not allow_abort:
        raise ReferenceWouldAbort("time outside [t1, t2] (trajectory_interpolation.cpp:32)")
    return _from_colmajor(out, 4)


def relative_pose_between_times(t1, P1, t2, P2, anchor, query):
    out = np.empty(16)
    rc = lib().kmc_oracle_relative_pose_between_times(t1, _ptr(_colmajor(P1)), t2, _ptr(_colmajor(P2)), anchor, query,
                                                      _ptr(out))
    if rc:
        raise ReferenceWouldAbort("time outside [t1, t2] (trajectory_interpolation.cpp:32)")
    return _from_colmajor(out, 4)


# ---- timestamp_mocking.cpp -----------------------------------------------------------------------------
def fraction_of_scan_completed(point4):
    p = _d(point4)
    fn = lib().kmc_oracle_fraction_of_scan_completed
    fn.argtypes = [_dp]
    return fn(_ptr(p))


def pseudo_time_stamp(point4, start, end):
    p = _d(point4)
    return lib().kmc_oracle_pseudo_time_stamp(_ptr(p), start, end)


def pseudo_time_stamps(cloud_n4, start, end):
    """cloud_n4: (n, 4) array (x y z 1); returns (n,) stamps — GetPseudoTimeStamps."""
    cloud = np.asarray(cloud_n4, dtype=np.float64)
    n = cloud.shape[0]
    cm = _colmajor(cloud)
    out = np.empty(n)
    lib().kmc_oracle_pseudo_time_stamps(_ptr(cm), n, start, end, _ptr(out))
    return out


# ---- data_io.cpp fixture builders -------------------------------------------------------------------------
def oxts_to_pose(oxts7, scale=1.0):
    out = np.empty(16)
    lib().kmc_oracle_oxts_to_pose(_ptr(_d(oxts7)), scale, _ptr(out))
    return _from_colmajor(out, 4)


def interpolate_trajectory(o1, o2, t):
    out = np.empty(16)
    rc = lib().kmc_oracle_interpolate_trajectory(_ptr(_d(o1)), _ptr(_d(o2)), t, _ptr(out))
    if rc:
        raise ReferenceWouldAbort("time outside [t1, t2]")
    return _from_colmajor(out, 4)


def make_frame_poses(o_prev, o_cur, o_next, stamp_start, stamp_end):
    a, b = np.empty(16), np.empty(16)
    rc = lib().kmc_oracle_make_frame_poses(_ptr(_d(o_prev)), _ptr(_d(o_cur)), _ptr(_d(o_next)), stamp_start, stamp_end,
                                           _ptr(a), _ptr(b))
    if rc:
        raise ReferenceWouldAbort("scan stamps outside the oxts interval")
    return _from_colmajor(a, 4), _from_colmajor(b, 4)


# ---- motion_compensation.cpp -------------------------------------------------------------------------------
def motion_compensate_point(t1, P1, t2, P2, point_stamp, point4, requested_time):
    out = np.empty(4)
    rc = lib().kmc_oracle_motion_compensate_point(t1, _ptr(_colmajor(P1)), t2, _ptr(_colmajor(P2)), point_stamp,
                                                  _ptr(_d(point4)), requested_time, _ptr(out))
    if rc:
        raise ReferenceWouldAbort("time outside [t1, t2]")
    return out


def motion_compensate_frame(cloud_n4, timestamps, T_start, T_end, stamp_start, stamp_end, requested_time):
    """The reference's MotionCompensateFrame: (n,4) double cloud (x y z w) + (n,) stamps -> (n,4) double cloud."""
    cloud = np.asarray(cloud_n4, dtype=np.float64)
    n = cloud.shape[0]
    cm = _colmajor(cloud)
    ts = _d(timestamps)
    out = np.empty(4 * n)
    rc = lib().kmc_oracle_motion_compensate_frame(_ptr(cm), _ptr(ts), n, _ptr(_colmajor(T_start)), _ptr(_colmajor(T_end)),
                                                  stamp_start, stamp_end, requested_time, _ptr(out))
    if rc:
        raise ReferenceWouldAbort("a point stamp or the requested time is outside [stamp_start, stamp_end]")
    return out.reshape(4, n).T.copy()


def deskew_xyzi_scan(xyzi_f32, T_start, T_end, stamp_start, stamp_end, requested_time, allow_abort=False):
    """Loader conversion + GetPseudoTimeStamps + MotionCompensateFrame for one float32 (n,4) xyzi scan.

    Returns the reference's double result as (n,4) (x', y', z', 1)."""
    pts = np.ascontiguousarray(xyzi_f32, dtype=np.float32).reshape(-1, 4)
    n = pts.shape[0]
    out = np.empty((n, 4), dtype=np.float64)
    rc = lib().kmc_oracle_deskew_xyzi_scan(pts.ctypes.data_as(_fp), n, _ptr(_colmajor(T_start)), _ptr(_colmajor(T_end)),
                                           stamp_start, stamp_end, requested_time, _ptr(out))
    if rc and not allow_abort:
        raise ReferenceWouldAbort("a point stamp or the requested time is outside [stamp_start, stamp_end]")
    return out


def timed_frames(xyzi_f32, points_per_frame, T_start, T_end, stamps3, n_threads):
    """Time the restated reference path over n_frames scans; returns (seconds, checksum)."""
    pts = np.ascontiguousarray(xyzi_f32, dtype=np.float32).reshape(-1)
    n_frames = pts.size // (4 * points_per_frame)
    ts = np.ascontiguousarray(np.stack([np.asarray(T).T.reshape(-1) for T in T_start]), dtype=np.float64)
    te = np.ascontiguousarray(np.stack([np.asarray(T).T.reshape(-1) for T in T_end]), dtype=np.float64)
    st = _d(stamps3).reshape(-1)
    chk = C.c_double(0.0)
    sec = lib().kmc_oracle_timed_frames(pts.ctypes.data_as(_fp), points_per_frame, n_frames, _ptr(ts), _ptr(te),
                                        _ptr(st), n_threads, C.byref(chk))
    return sec, chk.value


def project_pointcloud(cloud_n4, tf_c00_lo, R_rect_00, P_rect_3x4, max_range=15.0):
    """camera_model.cpp:5-36,38-95 for one camera, without the drawing.  Returns (uv (n,2), valid (n,) bool,
    color_scale (n,), xyz_rect (n,3))."""
    cloud = np.asarray(cloud_n4, dtype=np.float64)
    n = cloud.shape[0]
    cm = _colmajor(cloud)
    uv = np.empty((n, 2))
    valid = np.empty(n, dtype=np.int32)
    color = np.empty(n)
    rect = np.empty((n, 3))
    fn = lib().kmc_oracle_project_pointcloud
    fn.argtypes = [_dp, C.c_int64, _dp, _dp, _dp, C.c_double, _dp, C.POINTER(C.c_int32), _dp, _dp]
    fn.restype = None
    fn(_ptr(cm), n, _ptr(_colmajor(tf_c00_lo)), _ptr(_colmajor(R_rect_00)), _ptr(_colmajor(P_rect_3x4)), max_range, _ptr(uv),
       valid.ctypes.data_as(C.POINTER(C.c_int32)), _ptr(color), _ptr(rect))
    return uv, valid.astype(bool), color, rect


def hardware_threads() -> int:
    return int(lib().kmc_oracle_hardware_threads())
