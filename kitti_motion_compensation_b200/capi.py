"""ctypes binding of libkmc_b200.so — the C ABI declared in include/kmc_b200.h.

This module is plumbing for the Python-side tests and bench.py; the product is the shared library and the C++ mirror
of the reference API (include/kitti_motion_compensation/*.hpp).  It never falls back to a CPU implementation: if the
library has not been built, importing `lib()` raises, and every compute call needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libkmc_b200.so")

OK = 0
ERR_NULL_POINTER = -1
ERR_BAD_SIZE = -2
ERR_TIME_OUT_OF_RANGE = -3
ERR_EMPTY_INTERVAL = -4
ERR_NOT_RIGID = -5
ERR_CUDA = -6
ERR_NO_DEVICE = -7
ERR_BAD_MODE = -8
ERR_CAPACITY = -9
ERR_IO = -10
ERR_INTERNAL = -11
WARN_ACCURACY = 1  # positive = warning: outputs valid

TIME_FROM_AZIMUTH = 0
TIME_FROM_W = 1


class FrameParams(C.Structure):
    """kmc_b200_frame_params (64 bytes)."""
    _fields_ = [
        ("phi", C.c_float * 3), ("theta2", C.c_float),
        ("rho_perp", C.c_float * 3), ("c0", C.c_float),
        ("rho_par", C.c_float * 3), ("x_req", C.c_float),
        ("phi_x_rho", C.c_float * 3), ("wide", C.c_float),
    ]


class CameraParams(C.Structure):
    """kmc_b200_camera_params (112 bytes)."""
    _fields_ = [("rect", C.c_float * 12), ("pix", C.c_float * 12), ("min_depth", C.c_float), ("max_range", C.c_float),
                ("max_below", C.c_float), ("color_gain", C.c_float)]


FRAME_PARAMS_DTYPE = np.dtype([
    ("phi", np.float32, 3), ("theta2", np.float32), ("rho_perp", np.float32, 3), ("c0", np.float32),
    ("rho_par", np.float32, 3), ("x_req", np.float32), ("phi_x_rho", np.float32, 3), ("wide", np.float32)])
assert FRAME_PARAMS_DTYPE.itemsize == C.sizeof(FrameParams) == 64


class RunStats(C.Structure):
    """kmc_b200_run_stats"""
    _fields_ = [("frames", C.c_int64), ("frames_deskewed", C.c_int64), ("points_deskewed", C.c_int64),
                ("seconds_prepare", C.c_double), ("seconds_pipeline", C.c_double), ("seconds_total", C.c_double)]


class KmcError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"kmc_b200 status {status} ({message})")
        self.status = status


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); every exported symbol of include/kmc_b200.h appears here (tests check the two agree).
SIGNATURES = {
    "kmc_b200_version": (C.c_int, []),
    "kmc_b200_status_string": (C.c_char_p, [C.c_int]),
    "kmc_b200_last_error": (C.c_char_p, []),
    "kmc_b200_device_count": (C.c_int, []),
    "kmc_b200_launch_count": (C.c_uint64, []),
    "kmc_b200_frame_params_from_poses": (C.c_int, [_dp, _dp, C.c_double, C.c_double, C.c_double, C.POINTER(FrameParams)]),
    "kmc_b200_frame_params_from_twist": (C.c_int, [_dp, C.c_double, C.POINTER(FrameParams)]),
    "kmc_b200_frame_accuracy_bound": (C.c_int, [C.POINTER(FrameParams), C.c_double, _dp]),
    "kmc_b200_so3_hat": (C.c_int, [_dp, _dp]),
    "kmc_b200_so3_vee": (C.c_int, [_dp, _dp]),
    "kmc_b200_so3_exp": (C.c_int, [_dp, _dp]),
    "kmc_b200_so3_log": (C.c_int, [_dp, _dp]),
    "kmc_b200_so3_left_jacobian": (C.c_int, [_dp, _dp]),
    "kmc_b200_so3_inverse_left_jacobian": (C.c_int, [_dp, _dp]),
    "kmc_b200_se3_exp": (C.c_int, [_dp, _dp]),
    "kmc_b200_se3_log": (C.c_int, [_dp, _dp]),
    "kmc_b200_pose_at_time": (C.c_int, [C.c_double, _dp, C.c_double, _dp, C.c_double, _dp]),
    "kmc_b200_relative_pose_between_times": (C.c_int, [C.c_double, _dp, C.c_double, _dp, C.c_double, C.c_double, _dp]),
    "kmc_b200_fraction_of_scan_completed": (C.c_double, [C.c_double, C.c_double]),
    "kmc_b200_pseudo_time_stamp": (C.c_double, [C.c_double, C.c_double, C.c_double, C.c_double]),
    "kmc_b200_oxts_to_pose": (C.c_int, [C.c_double] * 7 + [_dp]),
    "kmc_b200_camera_params_from_calibration": (C.c_int, [_dp, _dp, _dp, C.c_double, C.POINTER(CameraParams)]),
    "kmc_b200_shard_range": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "kmc_b200_deskew_frame_device": (C.c_int, [_vp, _vp, C.c_int64, C.POINTER(FrameParams), C.c_int, _vp]),
    "kmc_b200_deskew_batch_device": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int32, C.c_int64, C.c_int, _vp]),
    "kmc_b200_project_frame_device": (C.c_int, [_vp, _vp, C.c_int64, C.POINTER(CameraParams), _vp]),
    "kmc_b200_deskew_project_frame_device": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.POINTER(FrameParams), C.POINTER(CameraParams),
                                                       C.c_int, _vp]),
    "kmc_b200_deskew_cloud_f64_device": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_double, C.c_double, C.c_double,
                                                   C.POINTER(FrameParams), _vp, _vp]),
    "kmc_b200_deskew_project_frame4_device": (C.c_int, [_vp, _vp, C.POINTER(_vp), C.c_int64, C.POINTER(FrameParams),
                                                        C.POINTER(CameraParams), C.c_int, _vp]),
    "kmc_b200_deskew_project_batch_device": (C.c_int, [_vp, _vp, C.POINTER(_vp), C.c_int32, _vp, _vp, C.c_int32, C.c_int64,
                                                       C.POINTER(CameraParams), C.c_int, _vp]),
    "kmc_b200_deskew_cloud_f64_batch_device": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int64, _vp, _vp]),
    "kmc_b200_pseudo_time_stamps_device": (C.c_int, [_vp, _vp, C.c_int64, C.c_double, C.c_double, _vp]),
    "kmc_b200_pseudo_time_stamps_xy_device": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_double, C.c_double, _vp]),
    "kmc_b200_check_fractions_device": (C.c_int, [_vp, C.c_int64, _vp, _vp]),
    "kmc_b200_frame_checksums_device": (C.c_int, [_vp, _vp, C.c_int32, C.c_int64, _vp, _vp]),
    "kmc_b200_synth_scans_device": (C.c_int, [_vp, C.c_int64, C.c_int32, C.c_int32, C.c_uint64, C.c_int64, _vp]),
    "kmc_b200_synth_frame_params": (C.c_int, [C.c_int32, C.c_uint64, C.c_int64, C.c_double, _vp, _vp]),
    "kmc_b200_handle_create": (C.c_int, [C.c_int, C.c_int64, C.POINTER(_vp)]),
    "kmc_b200_handle_destroy": (C.c_int, [_vp]),
    "kmc_b200_default_handle": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "kmc_b200_handle_set_file_callback": (C.c_int, [_vp, _vp, _vp]),
    "kmc_b200_handle_device": (C.c_int, [_vp]),
    "kmc_b200_handle_capacity": (C.c_int64, [_vp]),
    "kmc_b200_deskew_frame_host": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.POINTER(FrameParams), C.c_int]),
    "kmc_b200_deskew_batch_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int]),
    "kmc_b200_deskew_batch_multi_gpu": (C.c_int, [C.POINTER(_vp), C.c_int32, _vp, _vp, _vp, _vp, C.c_int32, C.c_int]),
    "kmc_b200_deskew_cloud_f64_host": (C.c_int, [_vp, _dp, _dp, _dp, C.c_int64, C.c_double, C.c_double, C.c_double,
                                                 C.POINTER(FrameParams), C.POINTER(C.c_int)]),
    "kmc_b200_deskew_cloud_f64_batch_host": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _vp, C.c_int32, _vp]),
    "kmc_b200_pseudo_time_stamps_xy_host": (C.c_int, [_vp, _dp, _dp, C.c_int64, C.c_double, C.c_double, _dp]),
    "kmc_b200_project_frame_host": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.POINTER(CameraParams)]),
    "kmc_b200_deskew_bin_file": (C.c_int, [_vp, C.c_char_p, C.c_char_p, C.POINTER(FrameParams), C.POINTER(C.c_int64)]),
    "kmc_b200_deskew_bin_files": (C.c_int, [_vp, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), _vp, C.c_int, C.c_int32,
                                            C.POINTER(C.c_int64)]),
    "kmc_b200_motion_compensate_run": (C.c_int, [_vp, C.c_char_p, C.c_int32, _vp]),
    "kmc_b200_run_prepare": (C.c_int, [C.c_char_p, C.c_int64, _vp, C.POINTER(C.c_int64)]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads the in-tree shared library; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m kitti_motion_compensation_b200.build` "
                "(or __graft_entry__.build()).  There is no CPU fallback for the deskew path.")
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().kmc_b200_last_error().decode(errors="replace")  # messages may quote bytes of a corrupt input file


last_warning = 0  # the most recent positive (warning) status seen by check()


def check(status: int) -> int:
    """Raises on a negative status; a positive one is a warning (outputs valid) and is returned and remembered."""
    global last_warning
    if status < 0:
        raise KmcError(status, last_error() or lib().kmc_b200_status_string(status).decode())
    if status > 0:
        last_warning = status
    return status


def _colmajor(m, n) -> np.ndarray:
    a = np.asarray(m, dtype=np.float64)
    assert a.shape == (n, n), a.shape
    return np.ascontiguousarray(a.T).reshape(-1)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


# ---- host prep ----------------------------------------------------------------------------------------------------
def frame_params_from_poses(T_start, T_end, t_start: float, t_end: float, t_req: float) -> FrameParams:
    out = FrameParams()
    a, b = _colmajor(T_start, 4), _colmajor(T_end, 4)
    check(lib().kmc_b200_frame_params_from_poses(_ptr(a), _ptr(b), t_start, t_end, t_req, C.byref(out)))
    return out


def frame_params_from_twist(xi, x_req: float) -> FrameParams:
    out = FrameParams()
    v = np.ascontiguousarray(xi, dtype=np.float64)
    check(lib().kmc_b200_frame_params_from_twist(_ptr(v), x_req, C.byref(out)))
    return out


def frame_accuracy_bound(params: FrameParams, max_range_m: float = 128.0) -> tuple[float, int]:
    """(upper bound in metres of max |dxyz| of the fp32 kernels vs the reference for this frame, status OK | WARN_ACCURACY)"""
    b = C.c_double()
    rc = check(lib().kmc_b200_frame_accuracy_bound(C.byref(params), max_range_m, C.byref(b)))
    return b.value, rc


def params_array(params) -> np.ndarray:
    """list of FrameParams -> structured numpy array (n,) with the 64-byte record layout."""
    arr = np.zeros(len(params), dtype=FRAME_PARAMS_DTYPE)
    for i, p in enumerate(params):
        C.memmove(arr[i:i + 1].ctypes.data, C.byref(p), 64)
    return arr


def _lie(name: str, arg, n_in: int | None, n_out: int, out_matrix: bool):
    a = _colmajor(arg, n_in) if n_in else np.ascontiguousarray(arg, dtype=np.float64).reshape(-1)
    out = np.empty(n_out * n_out if out_matrix else n_out)
    check(getattr(lib(), name)(_ptr(a), _ptr(out)))
    return out.reshape(n_out, n_out).T.copy() if out_matrix else out


def so3_hat(phi): return _lie("kmc_b200_so3_hat", phi, None, 3, True)
def so3_vee(m): return _lie("kmc_b200_so3_vee", m, 3, 3, False)
def so3_exp(phi): return _lie("kmc_b200_so3_exp", phi, None, 3, True)
def so3_log(R): return _lie("kmc_b200_so3_log", R, 3, 3, False)
def so3_left_jacobian(phi): return _lie("kmc_b200_so3_left_jacobian", phi, None, 3, True)
def so3_inverse_left_jacobian(phi): return _lie("kmc_b200_so3_inverse_left_jacobian", phi, None, 3, True)
def se3_exp(xi): return _lie("kmc_b200_se3_exp", xi, None, 4, True)
def se3_log(T): return _lie("kmc_b200_se3_log", T, 4, 6, False)


def pose_at_time(t1, P1, t2, P2, t):
    out = np.empty(16)
    check(lib().kmc_b200_pose_at_time(t1, _ptr(_colmajor(P1, 4)), t2, _ptr(_colmajor(P2, 4)), t, _ptr(out)))
    return out.reshape(4, 4).T.copy()


def relative_pose_between_times(t1, P1, t2, P2, anchor, query):
    out = np.empty(16)
    check(lib().kmc_b200_relative_pose_between_times(t1, _ptr(_colmajor(P1, 4)), t2, _ptr(_colmajor(P2, 4)), anchor,
                                                     query, _ptr(out)))
    return out.reshape(4, 4).T.copy()


def oxts_to_pose(lat, lon, alt, roll, pitch, yaw, scale: float = 1.0):
    out = np.empty(16)
    check(lib().kmc_b200_oxts_to_pose(lat, lon, alt, roll, pitch, yaw, scale, _ptr(out)))
    return out.reshape(4, 4).T.copy()


def camera_params_from_calibration(P_rect_3x4, R_rect_00, T_velo_to_cam, max_range: float = 15.0) -> CameraParams:
    out = CameraParams()
    P = np.ascontiguousarray(np.asarray(P_rect_3x4, dtype=np.float64).T).reshape(-1)  # 3x4 -> column-major
    check(lib().kmc_b200_camera_params_from_calibration(_ptr(P), _ptr(_colmajor(R_rect_00, 3)), _ptr(_colmajor(T_velo_to_cam, 4)),
                                                        max_range, C.byref(out)))
    return out


def fraction_of_scan_completed(x: float, y: float) -> float:
    return lib().kmc_b200_fraction_of_scan_completed(x, y)


def pseudo_time_stamp(x: float, y: float, start: float, end: float) -> float:
    return lib().kmc_b200_pseudo_time_stamp(x, y, start, end)


def shard_range(n_items: int, n_parts: int, index: int) -> tuple[int, int]:
    b, e = C.c_int64(), C.c_int64()
    check(lib().kmc_b200_shard_range(n_items, n_parts, index, C.byref(b), C.byref(e)))
    return b.value, e.value


def synth_frame_params(n_frames: int, seed: int, first_scan_index: int = 0, x_req: float = 0.5):
    """-> (structured params array (n,), twists (n, 6))"""
    params = np.zeros(n_frames, dtype=FRAME_PARAMS_DTYPE)
    xi = np.zeros((n_frames, 6))
    check(lib().kmc_b200_synth_frame_params(n_frames, seed, first_scan_index, x_req, params.ctypes.data, xi.ctypes.data))
    return params, xi


# ---- device entry points (raw device pointers, e.g. torch.Tensor.data_ptr()) ---------------------------------------
def deskew_frame_device(in_ptr: int, out_ptr: int, n_points: int, params: FrameParams, mode: int = TIME_FROM_AZIMUTH,
                        stream: int = 0) -> None:
    check(lib().kmc_b200_deskew_frame_device(in_ptr, out_ptr, n_points, C.byref(params), mode, stream))


def deskew_batch_device(in_ptr: int, out_ptr: int, offsets_ptr: int, params_ptr: int, n_frames: int, n_points_total: int,
                        mode: int = TIME_FROM_AZIMUTH, stream: int = 0) -> None:
    check(lib().kmc_b200_deskew_batch_device(in_ptr, out_ptr, offsets_ptr, params_ptr, n_frames, n_points_total, mode,
                                             stream))


def project_frame_device(in_ptr: int, uvzc_ptr: int, n_points: int, camera: CameraParams, stream: int = 0) -> None:
    check(lib().kmc_b200_project_frame_device(in_ptr, uvzc_ptr, n_points, C.byref(camera), stream))


def deskew_project_frame_device(in_ptr: int, out_ptr: int, uvzc_ptr: int, n_points: int, params: FrameParams, camera: CameraParams,
                                mode: int = TIME_FROM_AZIMUTH, stream: int = 0) -> None:
    check(lib().kmc_b200_deskew_project_frame_device(in_ptr, out_ptr, uvzc_ptr, n_points, C.byref(params), C.byref(camera), mode,
                                                     stream))


def deskew_project_frame4_device(in_ptr: int, out_ptr: int, uvzc_ptrs, n_points: int, params, cameras, mode: int = TIME_FROM_AZIMUTH,
                                 stream: int = 0) -> None:
    """Four cameras in one pass.  params may be None (project the input as it is); out_ptr may be 0."""
    planes = (_vp * 4)(*uvzc_ptrs)
    cams = (CameraParams * 4)(*cameras)
    check(lib().kmc_b200_deskew_project_frame4_device(in_ptr, out_ptr, planes, n_points, C.byref(params) if params is not None else None,
                                                      cams, mode, stream))


def deskew_project_batch_device(in_ptr: int, out_ptr: int, uvzc_ptrs, offsets_ptr: int, params_ptr: int, n_frames: int, n_points_total: int,
                                cameras, mode: int = TIME_FROM_AZIMUTH, stream: int = 0) -> None:
    """Deskew + projection onto len(cameras) (1 or 4) cameras for a whole batch of frames; out_ptr may be 0."""
    n_cam = len(cameras)
    planes = (_vp * n_cam)(*uvzc_ptrs)
    cams = (CameraParams * n_cam)(*cameras)
    check(lib().kmc_b200_deskew_project_batch_device(in_ptr, out_ptr, planes, n_cam, offsets_ptr, params_ptr, n_frames, n_points_total, cams,
                                                     mode, stream))


def deskew_cloud_f64_batch_device(cloud_ptr: int, stamps_ptr: int, out_ptr: int, offsets_ptr: int, params_ptr: int, times_ptr: int,
                                  n_frames: int, n_points_total: int, flags_ptr: int, stream: int = 0) -> None:
    check(lib().kmc_b200_deskew_cloud_f64_batch_device(cloud_ptr, stamps_ptr, out_ptr, offsets_ptr, params_ptr, times_ptr, n_frames,
                                                       n_points_total, flags_ptr, stream))


def pseudo_time_stamps_device(in_ptr: int, out_ptr: int, n_points: int, start: float, end: float, stream: int = 0) -> None:
    check(lib().kmc_b200_pseudo_time_stamps_device(in_ptr, out_ptr, n_points, start, end, stream))


def synth_scans_device(out_ptr: int, points_per_scan: int, n_scans: int, n_rings: int, seed: int,
                       first_scan_index: int = 0, stream: int = 0) -> None:
    check(lib().kmc_b200_synth_scans_device(out_ptr, points_per_scan, n_scans, n_rings, seed, first_scan_index, stream))


def check_fractions_device(xyzi_ptr: int, n_points: int, flags_ptr: int, stream: int = 0) -> None:
    check(lib().kmc_b200_check_fractions_device(xyzi_ptr, n_points, flags_ptr, stream))


def frame_checksums_device(xyzi_ptr: int, offsets_ptr: int, n_frames: int, n_points_total: int, sums_ptr: int, stream: int = 0) -> None:
    """Per-frame 64-bit position-weighted checksums of a device batch into a device uint64 array (n_frames)."""
    check(lib().kmc_b200_frame_checksums_device(xyzi_ptr, offsets_ptr, n_frames, n_points_total, sums_ptr, stream))


def frame_checksums_numpy(xyzi: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """The same checksum on the host (definition in include/kmc_b200.h), for tests."""
    words = np.ascontiguousarray(xyzi, dtype=np.float32).reshape(-1, 4).view(np.uint32).astype(np.uint64)
    out = np.zeros(len(offsets) - 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for f in range(len(out)):
            w = words[offsets[f]:offsets[f + 1]].reshape(-1)
            j = np.arange(w.size, dtype=np.uint64)
            out[f] = np.sum((w + np.uint64(0x9E3779B9)) * (np.uint64(2) * j + np.uint64(1)), dtype=np.uint64)
    return out


def launch_count() -> int:
    return int(lib().kmc_b200_launch_count())


# ---- handle + host entry points ---------------------------------------------------------------------------------------
class Handle:
    """Owns a device, streams and pinned/device staging buffers (kmc_b200_handle); non-copyable like KittiPclLoader."""

    def __init__(self, device: int = 0, capacity_points: int = 250_000):
        self._h = _vp()
        check(lib().kmc_b200_handle_create(device, capacity_points, C.byref(self._h)))

    def close(self) -> None:
        if self._h:
            lib().kmc_b200_handle_destroy(self._h)
            self._h = _vp()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def raw(self):
        return self._h

    @property
    def device(self) -> int:
        return lib().kmc_b200_handle_device(self._h)

    @property
    def capacity(self) -> int:
        return lib().kmc_b200_handle_capacity(self._h)

    def deskew_frame(self, xyzi: np.ndarray, params: FrameParams, mode: int = TIME_FROM_AZIMUTH,
                     out: np.ndarray | None = None) -> np.ndarray:
        pts = np.ascontiguousarray(xyzi, dtype=np.float32).reshape(-1, 4)
        if out is None:
            out = np.empty_like(pts)
        check(lib().kmc_b200_deskew_frame_host(self._h, pts.ctypes.data, out.ctypes.data, pts.shape[0], C.byref(params),
                                               mode))
        return out

    def deskew_frame_ptr(self, in_ptr: int, out_ptr: int, n_points: int, params: FrameParams,
                         mode: int = TIME_FROM_AZIMUTH) -> None:
        check(lib().kmc_b200_deskew_frame_host(self._h, in_ptr, out_ptr, n_points, C.byref(params), mode))

    def deskew_batch(self, xyzi: np.ndarray, offsets: np.ndarray, params: np.ndarray, mode: int = TIME_FROM_AZIMUTH,
                     out: np.ndarray | None = None) -> np.ndarray:
        pts = np.ascontiguousarray(xyzi, dtype=np.float32).reshape(-1, 4)
        offs = np.ascontiguousarray(offsets, dtype=np.int64)
        prm = np.ascontiguousarray(params, dtype=FRAME_PARAMS_DTYPE)
        assert offs.size == prm.size + 1 and offs[-1] == pts.shape[0]
        if out is None:
            out = np.empty_like(pts)
        check(lib().kmc_b200_deskew_batch_host(self._h, pts.ctypes.data, out.ctypes.data, offs.ctypes.data,
                                               prm.ctypes.data, prm.size, mode))
        return out

    def deskew_batch_ptr(self, in_ptr: int, out_ptr: int, offsets: np.ndarray, params: np.ndarray,
                         mode: int = TIME_FROM_AZIMUTH) -> None:
        offs = np.ascontiguousarray(offsets, dtype=np.int64)
        prm = np.ascontiguousarray(params, dtype=FRAME_PARAMS_DTYPE)
        check(lib().kmc_b200_deskew_batch_host(self._h, in_ptr, out_ptr, offs.ctypes.data, prm.ctypes.data, prm.size, mode))

    def deskew_cloud_f64(self, cloud_n4: np.ndarray, stamps: np.ndarray, t_start: float, t_end: float, t_req: float,
                         params: FrameParams):
        """The reference's MotionCompensateFrame layout: (n,4) double cloud (x y z 1) + (n,) stamps -> ((n,4) double, flags, status)."""
        cm = np.ascontiguousarray(np.asarray(cloud_n4, dtype=np.float64).T)  # column-major n x 4 == C-order 4 x n
        ts = np.ascontiguousarray(stamps, dtype=np.float64)
        out = np.empty_like(cm)
        flags = C.c_int(0)
        rc = lib().kmc_b200_deskew_cloud_f64_host(self._h, _ptr(cm), _ptr(ts), _ptr(out), ts.size, t_start, t_end, t_req,
                                                  C.byref(params), C.byref(flags))
        return out.T.copy(), flags.value, rc

    def deskew_cloud_f64_batch(self, clouds_n4, stamps, times, params: np.ndarray):
        """Several frames in the reference's layout through ONE pipeline: clouds_n4[f] (n_f,4) double, stamps[f] (n_f,),
        times (F,3) = (t_start, t_end, t_req), params structured array (F,) -> (list of (n_f,4) results, flags (F,), status)."""
        F = len(clouds_n4)
        cms = [np.ascontiguousarray(np.asarray(c, dtype=np.float64).T) for c in clouds_n4]
        tss = [np.ascontiguousarray(t, dtype=np.float64) for t in stamps]
        outs = [np.empty_like(c) for c in cms]
        n = np.array([t.size for t in tss], dtype=np.int64)
        tm = np.ascontiguousarray(times, dtype=np.float64).reshape(F, 3)
        prm = np.ascontiguousarray(params, dtype=FRAME_PARAMS_DTYPE)
        flags = np.zeros(F, dtype=np.int32)
        pa = (_vp * F)(*[c.ctypes.data for c in cms])
        ps = (_vp * F)(*[t.ctypes.data for t in tss])
        po = (_vp * F)(*[o.ctypes.data for o in outs])
        rc = lib().kmc_b200_deskew_cloud_f64_batch_host(self._h, pa, ps, po, n.ctypes.data, tm.ctypes.data, prm.ctypes.data, F, flags.ctypes.data)
        return [o.T.copy() for o in outs], flags, rc

    def pseudo_time_stamps(self, x: np.ndarray, y: np.ndarray, start: float, end: float) -> np.ndarray:
        xs = np.ascontiguousarray(x, dtype=np.float64)
        ys = np.ascontiguousarray(y, dtype=np.float64)
        out = np.empty_like(xs)
        check(lib().kmc_b200_pseudo_time_stamps_xy_host(self._h, _ptr(xs), _ptr(ys), xs.size, start, end, _ptr(out)))
        return out

    def project_frame(self, xyzi: np.ndarray, camera: CameraParams) -> np.ndarray:
        """(n,4) float32 xyzi -> (n,4) float32 (u, v, z_rect, colour | -1)."""
        pts = np.ascontiguousarray(xyzi, dtype=np.float32).reshape(-1, 4)
        out = np.empty_like(pts)
        check(lib().kmc_b200_project_frame_host(self._h, pts.ctypes.data, out.ctypes.data, pts.shape[0], C.byref(camera)))
        return out

    def deskew_bin_file(self, path_in: str, path_out: str, params: FrameParams) -> int:
        n = C.c_int64()
        check(lib().kmc_b200_deskew_bin_file(self._h, path_in.encode(), path_out.encode(), C.byref(params), C.byref(n)))
        return n.value

    def deskew_bin_files(self, paths_in: list[str], paths_out: list[str], params: np.ndarray, mode: int = TIME_FROM_AZIMUTH,
                         io_threads: int = 0) -> np.ndarray:
        """Many .bin files through one overlapped read -> H2D -> kernel -> D2H -> write pipeline; returns points per file."""
        prm = np.ascontiguousarray(params, dtype=FRAME_PARAMS_DTYPE)
        n = len(paths_in)
        assert len(paths_out) == n and prm.size == n
        a = (C.c_char_p * n)(*[os.fsencode(p) for p in paths_in])
        b = (C.c_char_p * n)(*[os.fsencode(p) for p in paths_out])
        points = np.zeros(n, dtype=np.int64)
        check(lib().kmc_b200_deskew_bin_files(self._h, n, a, b, prm.ctypes.data, mode, io_threads,
                                              points.ctypes.data_as(C.POINTER(C.c_int64))))
        return points

    def motion_compensate_run(self, run_folder: str, io_threads: int = 0) -> dict:
        """MotionCompensateRun (handlers.cpp:41-65) on a KITTI raw run folder; returns the kmc_b200_run_stats fields."""
        stats = RunStats()
        check(lib().kmc_b200_motion_compensate_run(self._h, os.fsencode(run_folder), io_threads, C.byref(stats)))
        return {name: getattr(stats, name) for name, _ in RunStats._fields_}


def run_prepare(run_folder: str):
    """Host half of MotionCompensateRun (no GPU): (n_frames, per-frame records of frames 1 .. n-2)."""
    n = C.c_int64(0)
    check(lib().kmc_b200_run_prepare(os.fsencode(run_folder), 0, None, C.byref(n)))
    params = np.zeros(max(n.value - 2, 0), dtype=FRAME_PARAMS_DTYPE)
    if params.size:
        check(lib().kmc_b200_run_prepare(os.fsencode(run_folder), params.size, params.ctypes.data, C.byref(n)))
    return n.value, params


def deskew_batch_multi_gpu(handles: list[Handle], xyzi: np.ndarray, offsets: np.ndarray, params: np.ndarray,
                           mode: int = TIME_FROM_AZIMUTH) -> np.ndarray:
    pts = np.ascontiguousarray(xyzi, dtype=np.float32).reshape(-1, 4)
    offs = np.ascontiguousarray(offsets, dtype=np.int64)
    prm = np.ascontiguousarray(params, dtype=FRAME_PARAMS_DTYPE)
    out = np.empty_like(pts)
    arr = (_vp * len(handles))(*[h.raw for h in handles])
    check(lib().kmc_b200_deskew_batch_multi_gpu(arr, len(handles), pts.ctypes.data, out.ctypes.data, offs.ctypes.data,
                                                prm.ctypes.data, prm.size, mode))
    return out
