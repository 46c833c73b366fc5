#!/usr/bin/env python
"""DeskewCloudF64BatchKernel (reference layout, 72 B/point) on one B200: one point per thread against two (f64_pair), by
frame shape (one huge frame ... thousands of small ones, even and odd point counts) and CTA size; every variant's output is
compared bit for bit with the first variant's.  GB/s = algorithmic bytes / CUDA-event time.
  python tools/sweep_f64_batch.py ["tune a;tune b;..."]"""
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
    torch.cuda.set_device(0)
    st = torch.cuda.current_stream().cuda_stream
    a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    b = torch.empty_like(a)
    print(f"torch copy_ (1 GiB): {2 * a.numel() / timed(lambda: b.copy_(a)) / 1e6:7.0f} GB/s", flush=True)
    del a, b
    tunes = ["f64_pair=0,f64_item_tiles=16", "", "f64_item_tiles=2", "f64_item_tiles=4", "f64_item_tiles=8", "f64_min_ctas=3,f64_item_tiles=4",
             "f64_block=128,f64_min_ctas=8,f64_item_tiles=4"]
    shapes = [(1, 39_000_000), (30, 1_300_000), (300, 130_000), (300, 130_001), (3000, 13_001)]
    if len(sys.argv) > 1:
        tunes = sys.argv[1].split(";")
    for F, pts in shapes:
        n = F * pts
        g = torch.Generator(device="cuda").manual_seed(7)
        cloud = torch.rand(4 * n, dtype=torch.float64, device="cuda", generator=g) * 100 - 50
        cloud.view(F, 4, pts)[:, 3, :] = 1.0
        stamps = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 0.1
        out = torch.empty_like(cloud)
        flags = torch.zeros(F, dtype=torch.int32, device="cuda")
        params, _ = capi.synth_frame_params(F, 20110926, 0, 0.5)
        d_par = torch.from_numpy(params.view(np.uint8).copy()).cuda()
        offs = torch.arange(0, (F + 1) * pts, pts, dtype=torch.int64, device="cuda")
        times = torch.tensor([[0.0, 0.1, 0.05]] * F, dtype=torch.float64, device="cuda").reshape(-1)

        def run():
            capi.deskew_cloud_f64_batch_device(cloud.data_ptr(), stamps.data_ptr(), out.data_ptr(), offs.data_ptr(), d_par.data_ptr(),
                                               times.data_ptr(), F, n, flags.data_ptr(), st)

        if F == 1:  # the single-frame kernel on the same frame
            import ctypes as C
            p = capi.FrameParams.from_buffer_copy(params[:1].tobytes())
            for t in ["f64_pair=0", "f64_pair=1", "f64_pair=1,f64_item_tiles=4", "f64_pair=1,f64_item_tiles=1", "f64_pair=1,f64_ctas=4",
                      "f64_pair=1,f64_ctas=8"]:
                os.environ["KMC_B200_TUNE"] = t
                ms = timed(lambda: capi.check(capi.lib().kmc_b200_deskew_cloud_f64_device(
                    cloud.data_ptr(), stamps.data_ptr(), out.data_ptr(), n, 0.0, 0.1, 0.05, C.byref(p), flags.data_ptr(), st)))
                print(f"single-frame kernel, {pts} points  {t:50s} {72 * n / ms / 1e6:7.0f} GB/s", flush=True)
            os.environ.pop("KMC_B200_TUNE", None)
        want = None
        for t in tunes:
            os.environ.pop("KMC_B200_TUNE", None)
            if t:
                os.environ["KMC_B200_TUNE"] = t
            out.zero_()
            ms = timed(run)
            torch.cuda.synchronize()
            if want is None:
                want = out.clone()
                same = "first"
            else:
                same = "bit-equal" if torch.equal(out.view(torch.int64), want.view(torch.int64)) else "DIFFERENT"
            print(f"{F:5d} x {pts:9d}  {t or 'default':62s} {72 * n / ms / 1e6:7.0f} GB/s  {same}", flush=True)
        os.environ.pop("KMC_B200_TUNE", None)
        del cloud, out, stamps, want


if __name__ == "__main__":
    main()
