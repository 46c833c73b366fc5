#!/usr/bin/env python
"""BASELINE configs[4]: one dense 10 M-point 128-beam frame through the single-frame kernel (320 MB of traffic, a ~50 us
kernel).  Times launch shapes three ways — (a) one launch between two events with the L2 flushed before it (what r01
reported), (b) K launches back to back between two events (launch gaps amortised, input still 2.5x the L2), (c) under ncu
this script is the workload for the per-launch gpu__time_duration.  Prints GB/s at 32 B/point.
  python tools/config5_probe.py [--tunes "a;b;c"] [--once]     (--once: 3 launches of the default shape, for ncu)"""
import argparse
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kitti_motion_compensation_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--tunes", default="")
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--rounds", type=int, default=3)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    n = args.points
    stream = torch.cuda.current_stream().cuda_stream
    d_in = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    capi.synth_scans_device(d_in.data_ptr(), n, 1, 128, 20110926, 0, stream)
    d_out = torch.empty_like(d_in)
    params, _ = capi.synth_frame_params(1, 20110926, 0, 0.5)
    p = capi.FrameParams.from_buffer_copy(params.tobytes())
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def launch():
        capi.deskew_frame_device(d_in.data_ptr(), d_out.data_ptr(), n, p, 0, stream)

    if args.once:
        for _ in range(3):
            flush.zero_()
            launch()
        torch.cuda.synchronize()
        return

    def gbs(us):
        return 32 * n / (us * 1e-6) / 1e9

    def single_flushed(fn, reps=40):
        us = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            us.append(a.elapsed_time(b) * 1e3)
        return statistics.median(us), min(us)

    def back_to_back(fn, k=20, reps=5):
        us = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(k):
                fn()
            b.record()
            torch.cuda.synchronize()
            us.append(a.elapsed_time(b) * 1e3 / k)
        return statistics.median(us), min(us)

    def row(name, fn):
        for _ in range(3):
            fn()
        m1, b1 = single_flushed(fn)
        m2, b2 = back_to_back(fn)
        print(f"{name:58s} single+flush {m1:6.1f} us = {gbs(m1):6.0f} GB/s (best {gbs(b1):6.0f}) | back-to-back {m2:6.1f} us = {gbs(m2):6.0f} GB/s (best {gbs(b2):6.0f})",
              flush=True)
        return gbs(m1), gbs(m2)

    os.environ.pop("KMC_B200_TUNE", None)
    tunes = [t.strip() for t in args.tunes.split(";") if t.strip()]
    table = {}
    for r in range(args.rounds):
        row("torch copy_ (same 160 + 160 MB)", lambda: d_out.copy_(d_in))
        os.environ.pop("KMC_B200_TUNE", None)
        table.setdefault("default", []).append(row("default shape", launch))
        for t in tunes:
            os.environ["KMC_B200_TUNE"] = t
            table.setdefault(t, []).append(row(t, launch))
        os.environ.pop("KMC_B200_TUNE", None)
    for t, vals in table.items():
        print(f"MEAN {t:54s} single+flush {statistics.fmean(v[0] for v in vals):6.0f} GB/s | back-to-back {statistics.fmean(v[1] for v in vals):6.0f} GB/s")


if __name__ == "__main__":
    main()
