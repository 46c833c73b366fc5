#!/usr/bin/env python
"""Sweeps the knobs of the reference-layout host path (kmc::MotionCompensateFrame -> kmc_b200_deskew_cloud_f64_host) and of
the single-scan float host call on one GPU: chunks per frame, chunk floor, host threads.  Each setting runs in a fresh process
(lib/bench_motion_compensate_frame for the drop-in call) because the handle's thread pool is sized once.
Writes gpurun_out/sweep_dropin.log."""
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BIN = os.path.join(ROOT, "kitti_motion_compensation_b200", "lib", "bench_motion_compensate_frame")
SCAN = os.path.join(ROOT, "tests", "golden", "kitti_2011_09_26_drive_0005_frame0.bin")


def dropin(tune: str, threads: int = 1) -> dict:
    env = dict(os.environ)
    if tune:
        env["KMC_B200_TUNE"] = tune
    r = subprocess.run([BIN, SCAN, "300", str(threads)], capture_output=True, text=True, env=env, timeout=300)
    return json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-300:]}


def single_scan_host(tune: str) -> dict:
    code = r'''
import json, os, statistics, sys, time
import numpy as np, torch
sys.path.insert(0, %r)
from kitti_motion_compensation_b200 import capi
n = 130_000
d = torch.empty((n, 4), dtype=torch.float32, device="cuda")
capi.synth_scans_device(d.data_ptr(), n, 1, 64, 20110926, 0, 0)
torch.cuda.synchronize()
host = d.cpu().numpy()
params, _ = capi.synth_frame_params(1, 20110926, 0, 0.5)
p = capi.FrameParams.from_buffer_copy(params.tobytes())
res = {}
with capi.Handle(0, 250_000) as h:
    out = np.empty_like(host)
    pin_in = torch.from_numpy(host).pin_memory(); pin_out = torch.empty_like(pin_in).pin_memory()
    for name, fn in (("pageable", lambda: h.deskew_frame(host, p, out=out)),
                     ("pinned", lambda: h.deskew_frame_ptr(pin_in.data_ptr(), pin_out.data_ptr(), n, p))):
        for _ in range(20): fn()
        t = []
        for _ in range(300):
            a = time.perf_counter(); fn(); t.append(time.perf_counter() - a)
        res[name + "_us_median"] = round(statistics.median(t) * 1e6, 1); res[name + "_us_min"] = round(min(t) * 1e6, 1)
print(json.dumps(res))
''' % ROOT
    env = dict(os.environ)
    if tune:
        env["KMC_B200_TUNE"] = tune
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    return json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-300:]}


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    lines = []

    def log(s):
        print(s, flush=True)
        lines.append(s)

    log(f"host threads on this box: {os.cpu_count()}")
    log("== kmc::MotionCompensateFrame(Frame const&, Time), real 123 397-point scan, one caller ==")
    for tune in ["", "f64_zc_points=0", "f64_chunk=131072,f64_parts=1", "f64_chunk=131072,f64_parts=1,f64_zc_ctas=8", "f64_chunk=131072,f64_parts=1,f64_zc_ctas=4",
                 "f64_chunk=131072,f64_parts=1,f64_zc_ctas=8,host_threads=12", "f64_chunk=131072,f64_parts=2,f64_zc_ctas=8", "f64_zc_ctas=8,host_threads=12",
                 "f64_chunk=131072,f64_parts=1,f64_zc_ctas=16", "f64_chunk=131072,f64_parts=1,f64_zc_points=0", "", "f64_chunk=131072,f64_parts=1,f64_zc_ctas=8",
                 "host_threads=1", "host_threads=2", "host_threads=4", "host_threads=12"]:
        log(f"{tune or 'default':34s} {json.dumps(dropin(tune))}")
    log("== the same, 2 / 4 / 8 concurrent callers (handle leases) ==")
    for threads in (2, 4, 8):
        log(f"threads={threads:<26d} {json.dumps(dropin('', threads))}")
    log("== kmc_b200_deskew_frame_host, one synthetic 130 000-point scan ==")
    for tune in ["", "frame_parts=1", "frame_parts=2", "frame_parts=3", "frame_parts=6", "frame_parts=8,min_chunk=8192",
                 "frame_parts=4,min_chunk=16384"]:
        log(f"{tune or 'default':34s} {json.dumps(single_scan_host(tune))}")
    with open(os.path.join(ROOT, "gpurun_out", "sweep_dropin.log"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
