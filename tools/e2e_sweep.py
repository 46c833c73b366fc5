#!/usr/bin/env python
"""End-to-end (host buffers -> kmc_b200_deskew_batch_host -> host buffers) throughput vs staging chunk size (one GPU)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kitti_motion_compensation_b200 import capi  # noqa: E402


def main():
    torch.cuda.set_device(0)
    points, scans = 130_000, 1000
    n = points * scans
    d_in = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    capi.synth_scans_device(d_in.data_ptr(), points, scans, 64, 20110926, 0)
    pin_in = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
    pin_out = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
    pin_in.copy_(d_in)
    torch.cuda.synchronize()
    params, _ = capi.synth_frame_params(scans, 20110926, 0, 0.5)
    offs = np.arange(0, (scans + 1) * points, points, dtype=np.int64)
    # raw PCIe ceilings
    for name, fn in (("H2D only", lambda: d_in.copy_(pin_in, non_blocking=True)), ("D2H only", lambda: pin_out.copy_(d_in, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        print(f"{name}: {5 * n * 16 / (time.perf_counter() - t0) / 1e9:.1f} GB/s", flush=True)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1):
            d_in.copy_(pin_in, non_blocking=True)
        with torch.cuda.stream(s2):
            pin_out.copy_(d_in, non_blocking=True)
    torch.cuda.synchronize()
    print(f"H2D + D2H concurrently: {5 * n * 16 / (time.perf_counter() - t0) / 1e9:.1f} GB/s each way", flush=True)
    for chunk_scans in (4, 8, 16, 32, 64, 128, 256):
        with capi.Handle(0, chunk_scans * points) as h:
            for _ in range(2):
                h.deskew_batch_ptr(pin_in.data_ptr(), pin_out.data_ptr(), offs, params)
            t0 = time.perf_counter()
            reps = 10
            for _ in range(reps):
                h.deskew_batch_ptr(pin_in.data_ptr(), pin_out.data_ptr(), offs, params)
            sec = time.perf_counter() - t0
        print(f"chunk {chunk_scans:4d} scans ({chunk_scans * points * 16 / 1e6:7.1f} MB): {reps * n / sec / 1e6:8.1f} Mpoints/s "
              f"= {reps * n * 16 / sec / 1e9:5.1f} GB/s each way", flush=True)
    # pageable caller memory (numpy arrays): staged through the handle's pinned slots on the host side
    page_in = pin_in.numpy().copy()
    page_out = np.empty_like(page_in)
    for chunk_scans, tune in ((1, ""), (8, ""), (32, ""), (8, "nt_copy=0"), (32, "nt_copy=0"), (8, "host_threads=16"), (32, "host_threads=8")):
        os.environ.pop("KMC_B200_TUNE", None)
        if tune:
            os.environ["KMC_B200_TUNE"] = tune
            print(f"KMC_B200_TUNE={tune} (host_threads only takes effect in a fresh process)", flush=True)
        with capi.Handle(0, chunk_scans * points) as h:
            h.deskew_batch_ptr(page_in.ctypes.data, page_out.ctypes.data, offs, params)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                h.deskew_batch_ptr(page_in.ctypes.data, page_out.ctypes.data, offs, params)
            sec = time.perf_counter() - t0
        ok = bool(np.array_equal(page_out, pin_out.numpy()))
        print(f"pageable, chunk {chunk_scans:4d} scans: {reps * n / sec / 1e6:8.1f} Mpoints/s = {reps * n * 16 / sec / 1e9:5.1f} GB/s each way  "
              f"equal_to_pinned_result={ok}", flush=True)


if __name__ == "__main__":
    main()
