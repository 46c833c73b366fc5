"""ctypes binding of oracle/_ref/libkmc_ref.so — the reference's OWN sources (motion_compensation.cpp,
trajectory_interpolation.cpp, lie_algebra.cpp, timestamp_mocking.cpp of the hot path, plus data_io.cpp, handlers.cpp,
camera_model.cpp and utils.cpp either side of it), compiled unmodified from /root/reference by `make -C oracle ref`
behind the C entry points of oracle/ref_shim.cpp.  OpenCV is a stub whose cv::circle records a draw list.

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
        _lib.kmc_ref_oxts_to_pose.argtypes = [_dp, C.c_double, _dp]
        _lib.kmc_ref_make_frame_poses.argtypes = [_dp, _dp, _dp, C.c_double, C.c_double, _dp, _dp]
        _lib.kmc_ref_interpolate_trajectory.argtypes = [_dp, _dp, C.c_double, _dp]
        _lib.kmc_ref_load_time_stamp.restype = C.c_double
        _lib.kmc_ref_load_time_stamp.argtypes = [C.c_char_p, C.c_int64]
        _lib.kmc_ref_load_oxts.argtypes = [C.c_char_p, C.c_int64, _dp]
        _lib.kmc_ref_load_pointcloud.restype = C.c_int64
        _lib.kmc_ref_load_pointcloud.argtypes = [C.c_char_p, _dp, _dp, C.c_int64]
        _lib.kmc_ref_write_pointcloud.argtypes = [C.c_char_p, C.c_int64, _dp, _dp, C.c_int64]
        _lib.kmc_ref_motion_compensate_run.restype = C.c_double
        _lib.kmc_ref_motion_compensate_run.argtypes = [C.c_char_p]
        _lib.kmc_ref_load_calibration.argtypes = [C.c_char_p, _dp, _dp, _dp, _dp]
        _lib.kmc_ref_project_pointcloud_on_frame.argtypes = [_dp, C.c_int64, _dp, _dp, _dp, C.POINTER(C.POINTER(C.c_int32)),
                                                             C.POINTER(_dp), C.POINTER(C.c_int64)]
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


# ---- data_io.cpp / handlers.cpp (the rows either side of the path) -------------------------------------------------
def oxts_to_pose(oxts7, scale=1.0):
    """OxtsToPose (data_io.cpp:68-88); oxts7 = [stamp, lat, lon, alt, roll, pitch, yaw]."""
    out = np.empty(16)
    lib().kmc_ref_oxts_to_pose(_ptr(_d(oxts7)), scale, _ptr(out))
    return _from_colmajor(out, 4)


def interpolate_trajectory(o1, o2, t):
    _check(o1[0], o2[0], t)
    out = np.empty(16)
    lib().kmc_ref_interpolate_trajectory(_ptr(_d(o1)), _ptr(_d(o2)), t, _ptr(out))
    return _from_colmajor(out, 4)


def make_frame_poses(o_prev, o_cur, o_next, stamp_start, stamp_end):
    """MakeFrame (data_io.cpp:253-269) -> (T_start, T_end)."""
    _check(o_prev[0], o_cur[0], stamp_start)
    _check(o_cur[0], o_next[0], stamp_end)
    a, b = np.empty(16), np.empty(16)
    lib().kmc_ref_make_frame_poses(_ptr(_d(o_prev)), _ptr(_d(o_cur)), _ptr(_d(o_next)), stamp_start, stamp_end, _ptr(a), _ptr(b))
    return _from_colmajor(a, 4), _from_colmajor(b, 4)


def load_time_stamp(path, frame_id):
    if not os.path.exists(path):
        raise FileNotFoundError(path)  # the reference prints and exit(0)s
    return lib().kmc_ref_load_time_stamp(os.fsencode(path), frame_id)


def load_oxts(run_folder, frame_id):
    out = np.empty(7)
    lib().kmc_ref_load_oxts(os.fsencode(run_folder), frame_id, _ptr(out))
    return out


def load_pointcloud(path):
    """KittiPclLoader::LoadPointcloud (data_io.cpp:101-138) -> ((n,4) double cloud x y z 1, (n,) double intensities)."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)  # the reference throws std::runtime_error, which would cross the C boundary
    n = os.path.getsize(path) // 16
    cloud, inten = np.empty((n, 4)), np.empty(n)
    got = lib().kmc_ref_load_pointcloud(os.fsencode(path), _ptr(cloud), _ptr(inten), n)
    assert got == n
    return cloud, inten


def write_pointcloud(folder, frame_id, cloud_n4, intensities):
    cloud = _d(np.asarray(cloud_n4, dtype=np.float64).reshape(-1, 4))
    inten = _d(intensities)
    lib().kmc_ref_write_pointcloud(os.fsencode(folder), frame_id, _ptr(cloud), _ptr(inten), cloud.shape[0])


def motion_compensate_run(run_folder) -> float:
    """handlers.cpp:41-65 on a KITTI run folder; returns wall seconds.  The caller guarantees the folder is well formed
    (the reference exit(0)s / aborts on missing files and out-of-range stamps)."""
    return lib().kmc_ref_motion_compensate_run(os.fsencode(run_folder))


def project_pointcloud_on_frame(cloud_n4, tf_c00_lo, R_rect_00, P_rects):
    """ProjectPointcloudOnFrame (camera_model.cpp:38-95) with cv::circle recording instead of drawing.
    P_rects: four 3x4 matrices (cameras 00..03).  Returns, per camera, (uv int32 (m,2), colour (m,3)) of the points the
    reference draws, in point order (max_range = the reference's default 15 m)."""
    cloud = np.asarray(cloud_n4, dtype=np.float64)
    n = cloud.shape[0]
    P = np.ascontiguousarray(np.stack([np.asarray(p, dtype=np.float64).T.reshape(-1) for p in P_rects]))  # 4 x 12, column-major each
    uv = [np.empty((n, 2), dtype=np.int32) for _ in range(4)]
    col = [np.empty((n, 3)) for _ in range(4)]
    uv_ptrs = (C.POINTER(C.c_int32) * 4)(*[a.ctypes.data_as(C.POINTER(C.c_int32)) for a in uv])
    col_ptrs = (_dp * 4)(*[_ptr(a) for a in col])
    counts = (C.c_int64 * 4)()
    lib().kmc_ref_project_pointcloud_on_frame(_ptr(_colmajor(cloud)), n, _ptr(_colmajor(tf_c00_lo)), _ptr(_colmajor(R_rect_00)), _ptr(P),
                                              uv_ptrs, col_ptrs, counts)
    return [(uv[k][:counts[k]].copy(), col[k][:counts[k]].copy()) for k in range(4)]


def load_calibration(folder):
    """LoadLidarExtrinsics(folder, to_cam=True) + LoadCameraCalibrations(folder) (data_io.cpp:168-210, 321-406).
    Returns (T_velo_to_cam 4x4, R_rect_00 3x3, [P_rect_00..03] 3x4, S_rect_00 (2,))."""
    for name in ("calib_velo_to_cam.txt", "calib_cam_to_cam.txt"):
        if not os.path.exists(os.path.join(folder, name)):
            raise FileNotFoundError(os.path.join(folder, name))  # the reference prints and exit(0)s
    T, R, P, S = np.empty(16), np.empty(9), np.empty(48), np.empty(2)
    lib().kmc_ref_load_calibration(os.fsencode(folder), _ptr(T), _ptr(R), _ptr(P), _ptr(S))
    return _from_colmajor(T, 4), _from_colmajor(R, 3), [P[12 * k:12 * k + 12].reshape(4, 3).T.copy() for k in range(4)], S
