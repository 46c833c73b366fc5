#!/usr/bin/env python
"""Launch-shape sweep of the batched deskew kernel on one GPU (run under gpurun).

Sets KMC_B200_TUNE between launches (the library re-reads it per call), times each shape with CUDA events and prints
achieved algorithmic GB/s (32 B/point).  A torch copy_ of the same buffers is timed as the box's copy ceiling.
"""
import argparse
import itertools
import json
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kitti_motion_compensation_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=4000)
    ap.add_argument("--points", type=int, default=130_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    ap.add_argument("--frame", action="store_true", help="time the single-frame kernel over the whole buffer instead of the batch kernel")
    ap.add_argument("--rounds", type=int, default=1, help="with --tunes: interleave the shapes this many times (A/B/A/B)")
    ap.add_argument("--tunes", default=None, help="semicolon-separated KMC_B200_TUNE strings to time instead of the staged sweep")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    n = args.scans * args.points
    d_in = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    d_out = torch.empty_like(d_in)
    d_off = torch.arange(0, (args.scans + 1) * args.points, args.points, dtype=torch.int64, device="cuda")
    params, _ = capi.synth_frame_params(args.scans, 20110926, 0, 0.5)
    d_par = torch.from_numpy(params.view(np.uint8).copy()).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    capi.synth_scans_device(d_in.data_ptr(), args.points, args.scans, 64, 20110926, 0, stream)
    torch.cuda.synchronize()

    def timed(fn):
        for _ in range(3):
            fn()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        ev[0].record()
        for i in range(args.steps):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
        return statistics.median(ms), min(ms)

    def gbs(ms):
        return 32 * n / (ms * 1e-3) / 1e9

    results = []
    med, best = timed(lambda: d_out.copy_(d_in))
    print(f"torch copy_            median {gbs(med):8.1f} GB/s  best {gbs(best):8.1f} GB/s", flush=True)
    results.append({"shape": "torch.copy_", "median_gbs": gbs(med), "best_gbs": gbs(best)})

    one = capi.FrameParams.from_buffer_copy(params[:1].tobytes())

    def run(tune):
        os.environ["KMC_B200_TUNE"] = tune
        if args.frame:
            med, best = timed(lambda: capi.deskew_frame_device(d_in.data_ptr(), d_out.data_ptr(), n, one, args.mode, stream))
        else:
            med, best = timed(lambda: capi.deskew_batch_device(d_in.data_ptr(), d_out.data_ptr(), d_off.data_ptr(), d_par.data_ptr(),
                                                               args.scans, n, args.mode, stream))
        print(f"{tune:48s} median {gbs(med):8.1f} GB/s  best {gbs(best):8.1f} GB/s", flush=True)
        results.append({"shape": tune, "median_gbs": gbs(med), "best_gbs": gbs(best)})
        return gbs(med)

    if args.tunes and args.rounds > 1:
        # A/B/A/B: the shapes take turns so that box-to-box and thermal drift hit all of them alike
        table = {t.strip(): [] for t in args.tunes.split(";")}
        for _ in range(args.rounds):
            for tune in table:
                table[tune].append(run(tune))
        for tune, vals in table.items():
            print(f"ROUNDS {tune:52s} mean {statistics.fmean(vals):8.1f}  min {min(vals):8.1f}  max {max(vals):8.1f} GB/s", flush=True)
        args.staged = False
    elif args.tunes:
        for tune in args.tunes.split(";"):
            run(tune.strip())
        args.staged = False
    else:
        args.staged = True
    stage1 = {}
    for v, u, h, blk in (itertools.product((1, 2), (1, 2), (0, 1), (128, 256, 512)) if args.staged else ()):
        ctas = 1024 // blk  # same resident threads per SM
        stage1[(v, u, h, blk)] = run(f"vec={v},unroll={u},hint={h},block={blk},ctas={ctas},item_tiles=16")
    v, u, h, blk = max(stage1, key=stage1.get) if stage1 else (2, 1, 0, 256)
    for threads, tiles in (itertools.product((768, 1024, 1280, 1536), (4, 8, 16, 32)) if args.staged else ()):
        if threads % blk:
            continue
        run(f"vec={v},unroll={u},hint={h},block={blk},ctas={threads // blk},item_tiles={tiles}")
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"scans": args.scans, "points": args.points, "mode": args.mode, "results": results}, f, indent=1)
    best = max((r for r in results if r["shape"] != "torch.copy_"), key=lambda r: r["median_gbs"])
    print("BEST", json.dumps(best))


if __name__ == "__main__":
    main()
