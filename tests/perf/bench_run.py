#!/usr/bin/env python
"""Run-level measurement: a whole KITTI raw run folder, files in -> deskewed files out.

  ours       kmc_b200_motion_compensate_run (text files once, overlapped pread -> H2D -> kernel -> D2H -> pwrite)
  reference  handlers.cpp:41-65 compiled from the reference's own sources (oracle/_ref), its execution model: one thread,
             one frame at a time — on a SUBSET of the run (it manages about one frame every 0.1-0.2 s)

The run is synthetic (tests/helpers.make_run_folder: 10 Hz scans of ~123 000 points, OxTS packets on a curve), the size of
2011_09_26_drive_0005 (154 frames).  Prints one JSON line per arm; nothing here is a bench.py number.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from kitti_motion_compensation_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=154)
    ap.add_argument("--points", type=int, default=123_000)
    ap.add_argument("--ref-frames", type=int, default=14, help="frames of the run handed to the reference handler")
    ap.add_argument("--dir", default=None, help="where to put the run (default: /dev/shm if present, else the temp dir)")
    ap.add_argument("--slot-points", type=int, default=2_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--io-threads", default="0,1,2,4,8,16")
    args = ap.parse_args()
    base = args.dir or ("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
    work = tempfile.mkdtemp(prefix="kmc_run_", dir=base)
    try:
        run = os.path.join(work, "2011_09_26_drive_9999_sync")
        t0 = time.perf_counter()
        info = helpers.make_run_folder(run, args.frames, args.points, seed=26, ragged=True)
        total_points = sum(len(s) for s in info["scans"][1:-1])
        print(json.dumps({"generated": run, "frames": args.frames, "points_deskewed_per_run": total_points,
                          "bytes_in": total_points * 16, "seconds": round(time.perf_counter() - t0, 2)}), flush=True)
        out_dir = os.path.join(run, "velodyne_points", "data_motion_compensated")
        a = time.perf_counter()
        with capi.Handle(0, 1024):
            pass
        t_ctx = time.perf_counter() - a
        a = time.perf_counter()
        h = capi.Handle(0, args.slot_points)
        t_handle = time.perf_counter() - a
        print(json.dumps({"cuda_context_plus_small_handle_s": round(t_ctx, 3), "run_handle_create_s": round(t_handle, 3),
                          "run_handle_bytes_pinned": 6 * args.slot_points * 16, "run_handle_bytes_device": 6 * args.slot_points * 16}), flush=True)
        with h:
            a = time.perf_counter()
            h.motion_compensate_run(run)  # first call: cold launches
            print(json.dumps({"arm": "kmc_b200_motion_compensate_run", "first_call_s": round(time.perf_counter() - a, 4)}), flush=True)
            for threads in [int(x) for x in args.io_threads.split(",")]:
                best, stats_best = None, None
                for _ in range(args.reps):
                    shutil.rmtree(out_dir)
                    a = time.perf_counter()
                    stats = h.motion_compensate_run(run, io_threads=threads)
                    sec = time.perf_counter() - a
                    if best is None or sec < best:
                        best, stats_best = sec, stats
                print(json.dumps({"arm": "kmc_b200_motion_compensate_run", "io_threads": threads, "frames": args.frames,
                                  "seconds": round(best, 4), "frames_per_s": round((args.frames - 2) / best, 1),
                                  "mpoints_per_s": round(total_points / best / 1e6, 1),
                                  "file_gb_per_s_each_way": round(total_points * 16 / best / 1e9, 2),
                                  "seconds_prepare": round(stats_best["seconds_prepare"], 4),
                                  "seconds_pipeline": round(stats_best["seconds_pipeline"], 4), "best_of": args.reps,
                                  "slot_points": args.slot_points, "dir": base}), flush=True)
        # the whole process, as a user of the reference's CLI sees it (CUDA context + handle creation included)
        import subprocess
        from kitti_motion_compensation_b200 import build
        cli = build.build_example()
        shutil.rmtree(out_dir)
        a = time.perf_counter()
        r = subprocess.run([cli, work, os.path.basename(run)], capture_output=True, text=True)
        print(json.dumps({"arm": "motion_compensate_runs CLI, one process", "returncode": r.returncode, "frames": args.frames,
                          "seconds_wall": round(time.perf_counter() - a, 3)}), flush=True)
        # the reference's own handler on the first --ref-frames frames of the same run
        from oracle import ref_binding as rb
        if rb.available() and args.ref_frames >= 3:
            sub = os.path.join(work, "ref_subset_sync")
            os.makedirs(os.path.join(sub, "velodyne_points", "data"))
            os.makedirs(os.path.join(sub, "oxts", "data"))
            for i in range(args.ref_frames):
                shutil.copy(os.path.join(run, "velodyne_points", "data", f"{i:010d}.bin"), os.path.join(sub, "velodyne_points", "data"))
                shutil.copy(os.path.join(run, "oxts", "data", f"{i:010d}.txt"), os.path.join(sub, "oxts", "data"))
            for name in ("velodyne_points/timestamps_start.txt", "velodyne_points/timestamps.txt", "velodyne_points/timestamps_end.txt",
                         "oxts/timestamps.txt"):
                shutil.copy(os.path.join(run, name), os.path.join(sub, name))
            sec = rb.motion_compensate_run(sub)
            pts = sum(len(s) for s in info["scans"][1:args.ref_frames - 1])
            print(json.dumps({"arm": "reference handlers.cpp MotionCompensateRun (oracle/_ref, " + rb.eigen_provider() + ")",
                              "frames": args.ref_frames, "seconds": round(sec, 3), "frames_per_s": round((args.ref_frames - 2) / sec, 2),
                              "mpoints_per_s": round(pts / sec / 1e6, 3), "threads": 1}), flush=True)
            # parity of the two arms on the shared frames
            import numpy as np
            worst = 0.0
            for i in range(1, args.ref_frames - 1):
                a = helpers.read_bin(os.path.join(sub, "velodyne_points", "data_motion_compensated", f"{i:010d}.bin"))
                b = helpers.read_bin(os.path.join(out_dir, f"{i:010d}.bin"))
                worst = max(worst, float(np.abs(a[:, :3].astype(np.float64) - b[:, :3]).max()))
            print(json.dumps({"parity_max_abs_m_between_arms": worst, "frames_compared": args.ref_frames - 2}), flush=True)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
