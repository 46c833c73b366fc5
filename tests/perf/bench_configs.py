#!/usr/bin/env python
"""Measures the BASELINE.json configs that are not bench.py's headline line (run under gpurun, one GPU):

  config 1  real KITTI scan (drive_0005 frame 0, 123 397 points) — single-frame kernel latency + parity vs the oracle
  config 2  one synthetic 130 000-point scan — single-frame kernel latency (L2-resident: a latency, not a bandwidth, number)
  config 5  one dense 10 M-point 128-beam frame — single-frame kernel bandwidth (320 MB of traffic > L2)
  + the C ABI host call (H2D + kernel + D2H) for one scan from pageable and from pinned memory.

Writes gpurun_out/configs.json.  The oracle is used here only as the checker / CPU timing of config 1.
"""
import ctypes as C
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from kitti_motion_compensation_b200 import capi  # noqa: E402


def time_launches(fn, reps=200, flush=None):
    for _ in range(5):
        fn()
    ms = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()  # evict the working set from L2 between launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return statistics.median(ms), min(ms)


def main():
    import helpers
    from oracle import binding as ob
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    out = {}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # 256 MB > 126 MB L2

    # ---- config 1 ---------------------------------------------------------------------------------------------------
    pts = helpers.real_scan()
    T_start, T_end, t0, t1, t2 = helpers.config1_frame()
    p = capi.frame_params_from_poses(T_start, T_end, t0, t2, t1)
    d_in = torch.from_numpy(pts).cuda()
    d_out = torch.empty_like(d_in)
    fn = lambda: capi.deskew_frame_device(d_in.data_ptr(), d_out.data_ptr(), len(pts), p, 0, stream)  # noqa: E731
    warm_med, warm_min = time_launches(fn)
    cold_med, cold_min = time_launches(fn, reps=50, flush=flush)
    tc = time.perf_counter()
    ref = ob.deskew_xyzi_scan(pts, T_start, T_end, t0, t2, t1)
    cpu_s = time.perf_counter() - tc
    err = float(np.abs(d_out.cpu().numpy()[:, :3].astype(np.float64) - ref[:, :3]).max())
    out["config1_real_scan"] = {"points": len(pts), "kernel_us_l2_warm_median": warm_med * 1e3, "kernel_us_l2_warm_best": warm_min * 1e3,
                                "kernel_us_l2_flushed_median": cold_med * 1e3, "mpoints_per_s_flushed": len(pts) / (cold_med * 1e-3) / 1e6,
                                "oracle_cpu_seconds_1_thread": cpu_s, "oracle_mpoints_per_s_1_thread": len(pts) / cpu_s / 1e6,
                                "max_abs_err_m": err}

    # the reference-layout entry point (what kmc::MotionCompensateFrame(Frame, t) calls): column-major double cloud +
    # per-point stamps from pageable host memory, result back into pageable host memory
    cloud = np.concatenate([pts[:, :3].astype(np.float64), np.ones((len(pts), 1))], axis=1)
    stamps = ob.pseudo_time_stamps(cloud, t0, t2)
    with capi.Handle(0, 1024) as h:
        for _ in range(3):
            res64, flags, rc = h.deskew_cloud_f64(cloud, stamps, t0, t2, t1, p)
        cm = np.ascontiguousarray(cloud.T)
        out64 = np.empty_like(cm)
        dp = C.POINTER(C.c_double)
        t = []
        for _ in range(30):
            a = time.perf_counter()
            capi.lib().kmc_b200_deskew_cloud_f64_host(h.raw, cm.ctypes.data_as(dp), stamps.ctypes.data_as(dp), out64.ctypes.data_as(dp),
                                                      len(pts), t0, t2, t1, C.byref(p), None)
            t.append(time.perf_counter() - a)
        ref64 = ob.motion_compensate_frame(cloud[::5], stamps[::5], T_start, T_end, t0, t2, t1)
        out["config1_reference_layout_f64_host_call"] = {
            "us_median": statistics.median(t) * 1e6, "mpoints_per_s": len(pts) / statistics.median(t) / 1e6,
            "max_abs_err_m": float(np.abs(res64[::5, :3] - ref64[:, :3]).max()),
            "speedup_vs_oracle_1_thread": cpu_s / statistics.median(t)}

    # ---- config 2 ---------------------------------------------------------------------------------------------------
    n = 130_000
    d_in = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    capi.synth_scans_device(d_in.data_ptr(), n, 1, 64, 20110926, 0, stream)
    d_out = torch.empty_like(d_in)
    params, xi = capi.synth_frame_params(1, 20110926, 0, 0.5)
    p = capi.FrameParams.from_buffer_copy(params.tobytes())
    fn = lambda: capi.deskew_frame_device(d_in.data_ptr(), d_out.data_ptr(), n, p, 0, stream)  # noqa: E731
    warm_med, warm_min = time_launches(fn)
    cold_med, _ = time_launches(fn, reps=50, flush=flush)
    host_pts = d_in.cpu().numpy()
    ref = ob.deskew_xyzi_scan(host_pts[::4], np.eye(4), ob.se3_exp(xi[0]), 0.0, 0.1, 0.05)
    err = float(np.abs(d_out.cpu().numpy()[::4, :3].astype(np.float64) - ref[:, :3]).max())
    out["config2_synthetic_130k"] = {"points": n, "kernel_us_l2_warm_median": warm_med * 1e3, "kernel_us_l2_warm_best": warm_min * 1e3,
                                     "kernel_us_l2_flushed_median": cold_med * 1e3, "mpoints_per_s_l2_warm": n / (warm_med * 1e-3) / 1e6,
                                     "max_abs_err_m": err}
    # host entry point for one scan: pageable and pinned
    with capi.Handle(0, 250_000) as h:
        res = np.empty_like(host_pts)
        for _ in range(3):
            h.deskew_frame(host_pts, p, out=res)
        t = []
        for _ in range(30):
            a = time.perf_counter()
            h.deskew_frame(host_pts, p, out=res)
            t.append(time.perf_counter() - a)
        pin_in = torch.from_numpy(host_pts).pin_memory()
        pin_out = torch.empty_like(pin_in).pin_memory()
        for _ in range(3):
            h.deskew_frame_ptr(pin_in.data_ptr(), pin_out.data_ptr(), n, p)
        tp = []
        for _ in range(30):
            a = time.perf_counter()
            h.deskew_frame_ptr(pin_in.data_ptr(), pin_out.data_ptr(), n, p)
            tp.append(time.perf_counter() - a)
        out["config2_host_call"] = {"pageable_us_median": statistics.median(t) * 1e6, "pinned_us_median": statistics.median(tp) * 1e6,
                                    "pinned_mpoints_per_s": n / statistics.median(tp) / 1e6}

    # ---- config 5 ---------------------------------------------------------------------------------------------------
    n = 10_000_000
    d_in = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    capi.synth_scans_device(d_in.data_ptr(), n, 1, 128, 20110926, 0, stream)
    d_out = torch.empty_like(d_in)
    fn = lambda: capi.deskew_frame_device(d_in.data_ptr(), d_out.data_ptr(), n, p, 0, stream)  # noqa: E731
    med, best = time_launches(fn, reps=50, flush=flush)
    out["config5_dense_10m"] = {"points": n, "kernel_us_median": med * 1e3, "kernel_us_best": best * 1e3,
                                "mpoints_per_s": n / (med * 1e-3) / 1e6, "gb_per_s_32B_per_point": 32 * n / (med * 1e-3) / 1e9}
    # ---- projection onto a camera, standalone and fused behind the deskew (100 M points, 1.6 GB) ---
    n = 100_000_000
    del d_in, d_out
    d_in = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    capi.synth_scans_device(d_in.data_ptr(), n, 1, 128, 20110926, 0, stream)
    d_out = torch.empty_like(d_in)
    d_pix = torch.empty_like(d_in)
    T = np.eye(4)
    T[:3, :3] = np.array([7.533745e-03, -9.999714e-01, -6.166020e-04, 1.480249e-02, 7.280733e-04, -9.998902e-01, 9.998621e-01,
                          7.523790e-03, 1.480755e-02]).reshape(3, 3)
    T[:3, 3] = [-4.069766e-03, -7.631618e-02, -2.717806e-01]
    R_rect = np.array([9.999239e-01, 9.837760e-03, -7.445048e-03, -9.869795e-03, 9.999421e-01, -4.278459e-03, 7.402527e-03,
                       4.351614e-03, 9.999631e-01]).reshape(3, 3)
    P2 = np.array([7.215377e+02, 0, 6.095593e+02, 4.485728e+01, 0, 7.215377e+02, 1.728540e+02, 2.163791e-01, 0, 0, 1, 2.745884e-03]).reshape(3, 4)
    cam = capi.camera_params_from_calibration(P2, R_rect, T, 15.0)
    proj = {}
    for name, bpp, fn in (
            ("deskew_only", 32, lambda: capi.deskew_frame_device(d_in.data_ptr(), d_out.data_ptr(), n, p, 0, stream)),
            ("project_only", 32, lambda: capi.project_frame_device(d_in.data_ptr(), d_pix.data_ptr(), n, cam, stream)),
            ("deskew_project_fused_cloud_and_pixels", 48,
             lambda: capi.deskew_project_frame_device(d_in.data_ptr(), d_out.data_ptr(), d_pix.data_ptr(), n, p, cam, 0, stream)),
            ("deskew_project_fused_pixels_only", 32,
             lambda: capi.deskew_project_frame_device(d_in.data_ptr(), 0, d_pix.data_ptr(), n, p, cam, 0, stream))):
        med, best = time_launches(fn, reps=20)
        proj[name] = {"bytes_per_point": bpp, "kernel_us_median": med * 1e3, "mpoints_per_s": n / (med * 1e-3) / 1e6,
                      "gb_per_s": bpp * n / (med * 1e-3) / 1e9}
    # all four cameras: four launches (4 x 32 B/point) vs one pass (16 + 64 B/point)
    cams = [cam] * 4
    planes = [torch.empty_like(d_in) for _ in range(4)]

    def four_launches():
        for c in range(4):
            capi.project_frame_device(d_in.data_ptr(), planes[c].data_ptr(), n, cams[c], stream)

    for name, bpp, fn in (
            ("project_4_cameras_four_launches", 128, four_launches),
            ("project_4_cameras_one_pass", 80,
             lambda: capi.deskew_project_frame4_device(d_in.data_ptr(), 0, [q.data_ptr() for q in planes], n, None, cams, 0, stream)),
            ("deskew_project_4_cameras_one_pass_with_cloud", 96,
             lambda: capi.deskew_project_frame4_device(d_in.data_ptr(), d_out.data_ptr(), [q.data_ptr() for q in planes], n, p, cams, 0, stream))):
        med, best = time_launches(fn, reps=10)
        proj[name] = {"bytes_per_point": bpp, "kernel_us_median": med * 1e3, "mpoints_per_s": n / (med * 1e-3) / 1e6,
                      "gb_per_s": bpp * n / (med * 1e-3) / 1e9}
    out["projection_100m_points"] = proj
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
