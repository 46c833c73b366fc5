#!/usr/bin/env python
"""Grid shapes of the secondary kernels on one B200: persistent (round 1) against one CTA per tile, for
  * DeskewCloudF64Kernel / DeskewCloudF64BatchKernel   (reference layout, 72 B/point)
  * PseudoTimeStampsKernel / PseudoTimeStampsXyKernel  (24 B/point)
  * DeskewProjectBatchKernel, 1 and 4 cameras, with and without the deskewed cloud (32 ... 96 B/point)
GB/s = algorithmic bytes / CUDA-event time, K launches back to back; torch copy_ on the same box as the ceiling."""
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from kitti_motion_compensation_b200 import capi  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b) / reps)
    return statistics.median(out)


def main():
    import ctypes as C
    from test_projection import calibration
    torch.cuda.set_device(0)
    st = torch.cuda.current_stream().cuda_stream
    a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    b = torch.empty_like(a)
    ms = timed(lambda: b.copy_(a))
    print(f"torch copy_ (1 GiB): {2 * a.numel() / ms / 1e6:7.0f} GB/s", flush=True)
    del a, b

    def sweep(name, bpp, n, fn, tunes):
        for t in tunes:
            os.environ.pop("KMC_B200_TUNE", None)
            if t:
                os.environ["KMC_B200_TUNE"] = t
            print(f"{name:46s} {t or 'default':34s} {bpp * n / timed(fn) / 1e6:7.0f} GB/s", flush=True)
        os.environ.pop("KMC_B200_TUNE", None)

    # ---- reference layout (double, column-major), one frame of 40 M points and a batch of 300 x 130 000
    n = 40_000_000
    cloud = torch.rand(4 * n, dtype=torch.float64, device="cuda") * 100 - 50
    cloud[3 * n:] = 1.0
    stamps = torch.rand(n, dtype=torch.float64, device="cuda") * 0.1
    out = torch.empty_like(cloud)
    flags = torch.zeros(4096, dtype=torch.int32, device="cuda")
    params, _ = capi.synth_frame_params(300, 20110926, 0, 0.5)
    p = capi.FrameParams.from_buffer_copy(params[:1].tobytes())
    f64_tunes = ["", "f64_ctas=8", "f64_ctas=4096", "f64_ctas=4096,f64_block=128", "f64_ctas=9,f64_block=128", "f64_ctas=16,f64_block=128"]
    sweep("DeskewCloudF64Kernel 40 M points", 72, n,
          lambda: capi.check(capi.lib().kmc_b200_deskew_cloud_f64_device(cloud.data_ptr(), stamps.data_ptr(), out.data_ptr(), n, 0.0, 0.1, 0.05, C.byref(p),
                                                                          flags.data_ptr(), st)), f64_tunes[:3])
    F, pts = 300, 130_000
    nb = F * pts
    offs = torch.arange(0, (F + 1) * pts, pts, dtype=torch.int64, device="cuda")
    d_par = torch.from_numpy(params.view(np.uint8).copy()).cuda()
    times = torch.tensor([[0.0, 0.1, 0.05]] * F, dtype=torch.float64, device="cuda").reshape(-1)
    # the batch layout: frame f's 4 N doubles at 4 offsets[f]; any content will do for timing, w = 1 keeps the common path
    for f in range(F):
        cloud[4 * f * pts + 3 * pts:4 * (f + 1) * pts] = 1.0
    sweep("DeskewCloudF64BatchKernel 300 x 130 000", 72, nb,
          lambda: capi.deskew_cloud_f64_batch_device(cloud.data_ptr(), stamps.data_ptr(), out.data_ptr(), offs.data_ptr(), d_par.data_ptr(),
                                                     times.data_ptr(), F, nb, flags.data_ptr(), st),
          ["", "f64_item_tiles=4", "f64_item_tiles=8", "f64_item_tiles=32", "f64_item_tiles=64", "f64_block=128", "f64_block=128,f64_item_tiles=8",
           "f64_block=128,f64_item_tiles=32", "f64_block=128,f64_item_tiles=64"])
    # ---- stamps
    xy_n = 100_000_000
    del cloud, out
    x = torch.rand(xy_n, dtype=torch.float64, device="cuda") - 0.5
    y = torch.rand(xy_n, dtype=torch.float64, device="cuda") - 0.5
    ts = torch.empty(xy_n, dtype=torch.float64, device="cuda")
    sweep("PseudoTimeStampsXyKernel 100 M points", 24, xy_n,
          lambda: capi.check(capi.lib().kmc_b200_pseudo_time_stamps_xy_device(x.data_ptr(), y.data_ptr(), ts.data_ptr(), xy_n, 0.0, 0.1, st)),
          ["", "stamp_vec=1", "stamp_ctas=4", "stamp_ctas=5", "stamp_ctas=8", "stamp_ctas=4096", "stamp_ctas=3"])
    del x, y
    d_in = torch.empty((xy_n, 4), dtype=torch.float32, device="cuda")
    capi.synth_scans_device(d_in.data_ptr(), xy_n, 1, 128, 20110926, 0, st)
    sweep("PseudoTimeStampsKernel 100 M points", 24, xy_n,
          lambda: capi.pseudo_time_stamps_device(d_in.data_ptr(), ts.data_ptr(), xy_n, 0.0, 0.1, st), ["", "stamp_ctas=12", "stamp_ctas=4096"])
    del ts
    # ---- deskew + projection over a batch of 700 x 130 000 points
    T, R_rect, P = calibration()
    cams = [capi.camera_params_from_calibration(P[k], R_rect, T, 15.0) for k in ("00", "01", "02", "03")]
    F = 700
    nb = F * 130_000
    offs = torch.arange(0, (F + 1) * 130_000, 130_000, dtype=torch.int64, device="cuda")
    params, _ = capi.synth_frame_params(F, 20110926, 0, 0.5)
    d_par = torch.from_numpy(params.view(np.uint8).copy()).cuda()
    cloud_out = torch.empty((nb, 4), dtype=torch.float32, device="cuda")
    planes = [torch.empty((nb, 4), dtype=torch.float32, device="cuda") for _ in range(4)]
    pp = [q.data_ptr() for q in planes]
    ptunes = ["", "pctas=65536", "pctas=65536,pitem_tiles=4", "pctas=65536,pitem_tiles=1", "pctas=65536,pblock=256", "pctas=65536,pblock=128"]
    for n_cam, with_cloud in ((1, False), (1, True), (4, False), (4, True)):
        bpp = 16 + 16 * n_cam + (16 if with_cloud else 0)
        sweep(f"DeskewProjectBatchKernel {n_cam} cam{' + cloud' if with_cloud else ''}", bpp, nb,
              lambda: capi.deskew_project_batch_device(d_in.data_ptr(), cloud_out.data_ptr() if with_cloud else 0, pp[:n_cam], offs.data_ptr(),
                                                       d_par.data_ptr(), F, nb, cams[:n_cam] if n_cam == 4 else [cams[2]], 0, st), ptunes)


if __name__ == "__main__":
    main()
