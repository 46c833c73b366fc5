#!/usr/bin/env python
"""Multi-GPU version of tools/zero_copy_probe.py: the host path of this pool does not scale device->host copy-engine
traffic with the number of GPUs (tools/pcie_multi_probe.py).  Does it treat SM-issued writes the same way?  One process per
GPU, all running at the same time, per variant:

  a  pipeline       kmc_b200_deskew_batch_host (copy engines both ways)                       [the shipped path]
  b  zero-copy      the batch kernel reading and writing the caller's pinned memory directly
  c  zero-copy out  H2D by copy engine in chunks, the kernel writes the caller's pinned memory

Prints per-GPU and aggregate Mpoints/s and GB/s each way.  Not a bench.py number.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, world, scans, points, chunk, seconds, barrier, results):
    try:
        from kitti_motion_compensation_b200 import capi
        torch.cuda.set_device(rank)
        n = points * scans
        d_in = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        capi.synth_scans_device(d_in.data_ptr(), points, scans, 64, 20110926, rank * scans)
        pin_in = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
        pin_out = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
        pin_in.copy_(d_in)
        params, _ = capi.synth_frame_params(scans, 20110926, rank * scans, 0.5)
        offs = np.arange(0, (scans + 1) * points, points, dtype=np.int64)
        d_offs = torch.from_numpy(offs).cuda()
        d_par = torch.from_numpy(params.view(np.uint8)).cuda()
        compute = torch.cuda.current_stream()
        copy_stream = torch.cuda.Stream()
        n_chunks = -(-scans // chunk)
        stage = [torch.empty((chunk * points, 4), dtype=torch.float32, device="cuda") for _ in range(3)]
        ready = [torch.cuda.Event() for _ in range(3)]
        free = [torch.cuda.Event() for _ in range(3)]
        par_bytes = params.view(np.uint8).reshape(scans, -1)
        chunk_offs, chunk_pars, bounds = [], [], []
        for c in range(n_chunks):
            f0, f1 = c * chunk, min(scans, (c + 1) * chunk)
            bounds.append((f0, f1, f0 * points, f1 * points))
            chunk_offs.append(torch.from_numpy(offs[f0:f1 + 1] - offs[f0]).cuda())
            chunk_pars.append(torch.from_numpy(np.ascontiguousarray(par_bytes[f0:f1])).cuda())
        handle = capi.Handle(rank, chunk * points)

        def variant_a():
            handle.deskew_batch_ptr(pin_in.data_ptr(), pin_out.data_ptr(), offs, params)

        def variant_b():
            capi.deskew_batch_device(pin_in.data_ptr(), pin_out.data_ptr(), d_offs.data_ptr(), d_par.data_ptr(), scans, n, 0, compute.cuda_stream)
            torch.cuda.synchronize()

        def variant_c():
            for c in range(n_chunks):
                f0, f1, p0, p1 = bounds[c]
                s = c % 3
                with torch.cuda.stream(copy_stream):
                    if c >= 3:
                        copy_stream.wait_event(free[s])
                    stage[s][: p1 - p0].copy_(pin_in[p0:p1], non_blocking=True)
                    ready[s].record(copy_stream)
                compute.wait_event(ready[s])
                capi.deskew_batch_device(stage[s].data_ptr(), pin_out[p0:p1].data_ptr(), chunk_offs[c].data_ptr(), chunk_pars[c].data_ptr(),
                                         f1 - f0, p1 - p0, 0, compute.cuda_stream)
                free[s].record(compute)
            torch.cuda.synchronize()

        out = {"gpu": rank}
        for name, fn in (("a", variant_a), ("b", variant_b), ("c", variant_c)):
            fn()
            torch.cuda.synchronize()
            barrier.wait(timeout=120)
            t0 = time.perf_counter()
            reps = 0
            while time.perf_counter() - t0 < seconds:
                fn()
                reps += 1
            out[name] = reps * n / (time.perf_counter() - t0) / 1e6
            barrier.wait(timeout=120)
        handle.close()
        results.put(out)
    except Exception as e:  # noqa: BLE001 - report and let the parent stop waiting
        results.put({"gpu": rank, "error": repr(e)})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--scans", type=int, default=512)
    ap.add_argument("--points", type=int, default=130_000)
    ap.add_argument("--chunk-scans", type=int, default=32)
    ap.add_argument("--seconds", type=float, default=1.5)
    args = ap.parse_args()
    mp.set_start_method("spawn", force=True)
    for world in sorted({1, args.gpus}):
        barrier = mp.Barrier(world)
        results = mp.Queue()
        procs = [mp.Process(target=worker, args=(r, world, args.scans, args.points, args.chunk_scans, args.seconds, barrier, results))
                 for r in range(world)]
        for p in procs:
            p.start()
        rows = []
        try:
            for _ in procs:
                rows.append(results.get(timeout=150))
        except Exception as e:  # noqa: BLE001
            print(f"{world} GPU(s): a worker did not answer ({type(e).__name__})", flush=True)
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.terminate()
        bad = [r for r in rows if "error" in r]
        if bad or len(rows) != world:
            print(f"{world} GPU(s): failed: {bad}", flush=True)
            continue
        rows.sort(key=lambda r: r["gpu"])
        for key, label in (("a", "a pipeline (copy engines both ways)"), ("b", "b zero-copy in + out"), ("c", "c copy engine in, kernel writes host")):
            total = sum(r[key] for r in rows)
            print(f"{world} GPU(s)  {label:40s} aggregate {total:8.1f} Mpoints/s = {total * 16 / 1e3:6.1f} GB/s each way   per GPU: "
                  + " ".join(f"{r[key]:.0f}" for r in rows), flush=True)


if __name__ == "__main__":
    main()
