#!/usr/bin/env python
"""End-to-end alternatives for pinned host buffers (one GPU): does the deskew kernel reading and/or writing the caller's
pinned memory DIRECTLY over PCIe (UVA zero-copy, no staging in HBM, no copy engines) beat the 3-slot
H2D -> kernel -> D2H pipeline of kmc_b200_deskew_batch_host?

  a  pipeline        kmc_b200_deskew_batch_host (copy engines both ways, staged in HBM)          [the shipped path]
  b  zero-copy both  kmc_b200_deskew_batch_device(in = pinned host, out = pinned host), one launch
  c  zero-copy out   H2D by copy engine in chunks, kernel writes the caller's pinned memory
  d  zero-copy in    kernel reads the caller's pinned memory, D2H by copy engine in chunks

Every variant is checked bit-for-bit against the resident result.  Results go to stdout (tee into profiles/).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kitti_motion_compensation_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=1000)
    ap.add_argument("--points", type=int, default=130_000)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--chunk-scans", type=int, default=32)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    points, scans = args.points, args.scans
    n = points * scans
    d_in = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    d_out = torch.empty_like(d_in)
    capi.synth_scans_device(d_in.data_ptr(), points, scans, 64, 20110926, 0)
    pin_in = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
    pin_out = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
    pin_in.copy_(d_in)
    params, _ = capi.synth_frame_params(scans, 20110926, 0, 0.5)
    offs = np.arange(0, (scans + 1) * points, points, dtype=np.int64)
    d_offs = torch.from_numpy(offs).cuda()
    d_par = torch.from_numpy(params.view(np.uint8)).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    capi.deskew_batch_device(d_in.data_ptr(), d_out.data_ptr(), d_offs.data_ptr(), d_par.data_ptr(), scans, n, 0, stream)
    torch.cuda.synchronize()
    want = d_out.cpu()

    def report(name, fn, check=True):
        pin_out.zero_()
        fn()
        torch.cuda.synchronize()
        ok = bool(torch.equal(pin_out, want)) if check else None
        t0 = time.perf_counter()
        for _ in range(args.reps):
            fn()
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / args.reps
        print(f"{name:58s} {n / sec / 1e6:9.1f} Mpoints/s  {n * 16 / sec / 1e9:6.1f} GB/s each way  {sec * 1e3:8.2f} ms  bit-equal={ok}", flush=True)
        return n / sec / 1e6

    # a: the shipped pipeline
    with capi.Handle(0, args.chunk_scans * points) as h:
        report(f"a pipeline (chunks of {args.chunk_scans} scans)", lambda: h.deskew_batch_ptr(pin_in.data_ptr(), pin_out.data_ptr(), offs, params))

    # b: zero-copy both ways, one launch, for a few launch shapes
    for tune in (None, "ctas=4,block=256", "ctas=2,block=256", "ctas=1,block=256", "ctas=9,block=128,unroll=2"):
        if tune is None:
            os.environ.pop("KMC_B200_TUNE", None)
        else:
            os.environ["KMC_B200_TUNE"] = tune
        try:
            report(f"b zero-copy in+out, tune={tune or 'default'}",
                   lambda: capi.deskew_batch_device(pin_in.data_ptr(), pin_out.data_ptr(), d_offs.data_ptr(), d_par.data_ptr(), scans, n, 0, stream))
        except Exception as e:  # a tune string this build does not know
            print(f"b tune={tune}: {e}", flush=True)
    os.environ.pop("KMC_B200_TUNE", None)

    # c / d: one side by copy engine in chunks, the other side direct
    chunk = args.chunk_scans
    copy_stream = torch.cuda.Stream()
    n_chunks = -(-scans // chunk)
    stage = [torch.empty((chunk * points, 4), dtype=torch.float32, device="cuda") for _ in range(3)]
    evs_ready = [torch.cuda.Event() for _ in range(3)]
    evs_free = [torch.cuda.Event() for _ in range(3)]
    compute = torch.cuda.current_stream()

    def sub_tables(c):
        f0, f1 = c * chunk, min(scans, (c + 1) * chunk)
        return f0, f1, f0 * points, f1 * points

    # per-chunk tables (offsets relative to the chunk) prepared once
    chunk_offs = [torch.from_numpy(offs[sub_tables(c)[0]:sub_tables(c)[1] + 1] - offs[sub_tables(c)[0]]).cuda() for c in range(n_chunks)]
    par_bytes = params.view(np.uint8).reshape(scans, -1)
    chunk_pars = [torch.from_numpy(np.ascontiguousarray(par_bytes[sub_tables(c)[0]:sub_tables(c)[1]])).cuda() for c in range(n_chunks)]

    def variant_c():
        for c in range(n_chunks):
            f0, f1, p0, p1 = sub_tables(c)
            s = c % 3
            with torch.cuda.stream(copy_stream):
                if c >= 3:
                    copy_stream.wait_event(evs_free[s])
                stage[s][: p1 - p0].copy_(pin_in[p0:p1], non_blocking=True)
                evs_ready[s].record(copy_stream)
            compute.wait_event(evs_ready[s])
            capi.deskew_batch_device(stage[s].data_ptr(), pin_out[p0:p1].data_ptr(), chunk_offs[c].data_ptr(), chunk_pars[c].data_ptr(),
                                     f1 - f0, p1 - p0, 0, compute.cuda_stream)
            evs_free[s].record(compute)

    def variant_d():
        for c in range(n_chunks):
            f0, f1, p0, p1 = sub_tables(c)
            s = c % 3
            if c >= 3:
                compute.wait_event(evs_free[s])
            capi.deskew_batch_device(pin_in[p0:p1].data_ptr(), stage[s].data_ptr(), chunk_offs[c].data_ptr(), chunk_pars[c].data_ptr(),
                                     f1 - f0, p1 - p0, 0, compute.cuda_stream)
            evs_ready[s].record(compute)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(evs_ready[s])
                pin_out[p0:p1].copy_(stage[s][: p1 - p0], non_blocking=True)
                evs_free[s].record(copy_stream)

    report(f"c H2D copy engine + kernel writes pinned host (chunks {chunk})", variant_c)
    report(f"d kernel reads pinned host + D2H copy engine (chunks {chunk})", variant_d)

    # single KITTI-size frame latency: pipeline vs zero-copy (what a MotionCompensateFrame caller with pinned memory sees)
    one = points
    p1 = capi.frame_params_from_twist([1.3, 0.02, -0.01, 0.003, -0.004, 0.05], 0.5)
    with capi.Handle(0, 250_000) as h:
        for name, fn in (("frame_host pipeline", lambda: h.deskew_frame_ptr(pin_in.data_ptr(), pin_out.data_ptr(), one, p1) if hasattr(h, "deskew_frame_ptr") else None),
                         ("frame zero-copy", lambda: (capi.deskew_frame_device(pin_in.data_ptr(), pin_out.data_ptr(), one, p1, 0, stream), torch.cuda.synchronize()))):
            fn(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(200):
                fn()
            torch.cuda.synchronize()
            print(f"one {one}-point frame, {name}: {(time.perf_counter() - t0) / 200 * 1e6:8.1f} us per call", flush=True)


if __name__ == "__main__":
    main()
