#!/usr/bin/env python
"""Launch-shape sweep of the projection kernels on one B200: threads per CTA x resident CTAs per SM x
128/256-bit accesses, for the single-camera kernel (project only, fused with and without the cloud output) and the
four-camera kernel (project only, fused with the cloud output).  Baselines on the same buffers: torch copy_ (1 read :
1 write) and fill_ (write only).  GB/s are algorithmic bytes (16 B per point read or written) / CUDA-event time.
"""
import argparse
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kitti_motion_compensation_b200 import capi  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=100_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--blocks", default="128,256")
    ap.add_argument("--ctas", default="2,3,4,5,6,8,9,10,12,16")
    ap.add_argument("--vecs", default="1,2")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    n = args.points
    stream = torch.cuda.current_stream().cuda_stream
    d_in = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    capi.synth_scans_device(d_in.data_ptr(), n, 1, 128, 20110926, 0, stream)
    d_out = torch.empty_like(d_in)
    planes = [torch.empty_like(d_in) for _ in range(4)]
    T = np.eye(4)
    T[:3, :3] = np.array([7.533745e-03, -9.999714e-01, -6.166020e-04, 1.480249e-02, 7.280733e-04, -9.998902e-01, 9.998621e-01,
                          7.523790e-03, 1.480755e-02]).reshape(3, 3)
    T[:3, 3] = [-4.069766e-03, -7.631618e-02, -2.717806e-01]
    R_rect = np.array([9.999239e-01, 9.837760e-03, -7.445048e-03, -9.869795e-03, 9.999421e-01, -4.278459e-03, 7.402527e-03,
                       4.351614e-03, 9.999631e-01]).reshape(3, 3)
    P2 = np.array([7.215377e+02, 0, 6.095593e+02, 4.485728e+01, 0, 7.215377e+02, 1.728540e+02, 2.163791e-01, 0, 0, 1, 2.745884e-03]).reshape(3, 4)
    cam = capi.camera_params_from_calibration(P2, R_rect, T, 15.0)
    cams = [cam] * 4
    params, _ = capi.synth_frame_params(1, 20110926, 0, 0.5)
    p = capi.FrameParams.from_buffer_copy(params.tobytes())
    pp = [q.data_ptr() for q in planes]

    ms = timed(lambda: d_out.copy_(d_in), args.reps)
    print(f"torch copy_ (16 B read + 16 B written per point): {32 * n / ms / 1e6:7.0f} GB/s", flush=True)
    ms = timed(lambda: d_out.fill_(1.0), args.reps)
    print(f"torch fill_ (16 B written per point)            : {16 * n / ms / 1e6:7.0f} GB/s", flush=True)

    kernels = [
        ("project_only", 32, lambda: capi.project_frame_device(d_in.data_ptr(), planes[0].data_ptr(), n, cam, stream)),
        ("fused_cloud+pixels", 48, lambda: capi.deskew_project_frame_device(d_in.data_ptr(), d_out.data_ptr(), planes[0].data_ptr(), n, p, cam, 0, stream)),
        ("fused_pixels_only", 32, lambda: capi.deskew_project_frame_device(d_in.data_ptr(), 0, planes[0].data_ptr(), n, p, cam, 0, stream)),
        ("4cam_project_only", 80, lambda: capi.deskew_project_frame4_device(d_in.data_ptr(), 0, pp, n, None, cams, 0, stream)),
        ("4cam_fused_with_cloud", 96, lambda: capi.deskew_project_frame4_device(d_in.data_ptr(), d_out.data_ptr(), pp, n, p, cams, 0, stream)),
    ]
    blocks = [int(x) for x in args.blocks.split(",")]
    ctas = [int(x) for x in args.ctas.split(",")]
    vecs = [int(x) for x in args.vecs.split(",")]
    best = {}
    print("shape                  " + "".join(f"{name:>24s}" for name, _, _ in kernels), flush=True)
    for block, c, v in itertools.product(blocks, ctas, vecs):
        if block * c > 2048 and c <= 16:  # c > 16: one CTA per tile (not persistent), residency is the hardware's business
            continue
        os.environ["KMC_B200_TUNE"] = f"pblock={block},pctas={c},pvec={v}"
        row = []
        for name, bpp, fn in kernels:
            gbs = bpp * n / timed(fn, args.reps) / 1e6
            row.append(gbs)
            if gbs > best.get(name, (0, ""))[0]:
                best[name] = (gbs, os.environ["KMC_B200_TUNE"])
        print(f"{os.environ['KMC_B200_TUNE']:23s}" + "".join(f"{g:24.0f}" for g in row), flush=True)
    os.environ.pop("KMC_B200_TUNE", None)
    row = [bpp * n / timed(fn, args.reps) / 1e6 for _, bpp, fn in kernels]
    print(f"{'default':23s}" + "".join(f"{g:24.0f}" for g in row), flush=True)
    for name, (gbs, tune) in best.items():
        print(f"best {name:24s} {gbs:7.0f} GB/s  {tune}", flush=True)


if __name__ == "__main__":
    main()
