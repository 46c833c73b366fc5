#!/usr/bin/env python
"""Latency of kmc_b200_deskew_frame_host for ONE scan (the 10 Hz sensor case, BASELINE configs[0]/[1]) over the knobs of the
zero-copy small-transfer path and against the copy-engine pipeline, pinned and pageable caller memory, several scan sizes.
Each setting runs in a fresh process (KMC_B200_TUNE is read per call, the host pool is sized once).  gpurun_out/sweep_single_scan.log"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r'''
import json, os, statistics, sys, time
import numpy as np, torch
sys.path.insert(0, %r)
from kitti_motion_compensation_b200 import capi
res = {}
params, _ = capi.synth_frame_params(1, 20110926, 0, 0.5)
p = capi.FrameParams.from_buffer_copy(params.tobytes())
with capi.Handle(0, 250_000) as h:
    for n in %s:
        d = torch.empty((n, 4), dtype=torch.float32, device="cuda")
        capi.synth_scans_device(d.data_ptr(), n, 1, 64, 20110926, 0, 0)
        torch.cuda.synchronize()
        host = d.cpu().numpy()
        out = np.empty_like(host)
        pin_in = torch.from_numpy(host).pin_memory(); pin_out = torch.empty_like(pin_in).pin_memory()
        want = torch.empty_like(d)
        capi.deskew_frame_device(d.data_ptr(), want.data_ptr(), n, p, 0, 0)
        torch.cuda.synchronize()
        for name, fn in (("pinned", lambda: h.deskew_frame_ptr(pin_in.data_ptr(), pin_out.data_ptr(), n, p)),
                         ("pageable", lambda: h.deskew_frame(host, p, out=out))):
            for _ in range(20): fn()
            t = []
            for _ in range(300):
                a = time.perf_counter(); fn(); t.append(time.perf_counter() - a)
            res[f"{n}_{name}_us"] = round(statistics.median(t) * 1e6, 1)
        assert torch.equal(pin_out.cuda(), want) and np.array_equal(out, want.cpu().numpy())
print(json.dumps(res))
'''


def run(tune, sizes):
    env = dict(os.environ)
    env.pop("KMC_B200_TUNE", None)
    if tune:
        env["KMC_B200_TUNE"] = tune
    r = subprocess.run([sys.executable, "-c", CODE % (ROOT, repr(sizes))], capture_output=True, text=True, env=env, timeout=900)
    return json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-400:]}


def main():
    lines = []

    def log(s):
        print(s, flush=True)
        lines.append(s)

    if len(sys.argv) > 1:  # python tools/sweep_single_scan.py "tune a;tune b" [sizes]
        sizes = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [130_000]
        for tune in sys.argv[1].split(";"):
            log(f"{tune or 'default':52s} {json.dumps(run(tune, sizes))}")
        return
    log("== one 130 000-point scan, zero-copy shape sweep (us, median of 300) ==")
    for tune in ["", "zc_tiles=4", "zc_tiles=2", "zc_unroll=2", "zc_unroll=2,zc_tiles=2", "zc_hint=1", "zc_hint=1,zc_tiles=4", "zc_block=512", "zc_block=512,zc_tiles=2",
                 "zc_block=128,zc_ctas=2,zc_tiles=4", "zc_ctas=2,zc_tiles=4", "", "zc_tiles=4", "zc_tiles=8", "zc_unroll=2,zc_tiles=4", "zc_points=0"]:
        log(f"{tune or 'default (zero copy, 1 x 256 per SM, 128-bit)':52s} {json.dumps(run(tune, [130_000]))}")
    log("== scan size: zero copy vs copy-engine pipeline (zc_points=0) ==")
    sizes = [32_768, 65_536, 130_000, 250_000, 500_000, 1_000_000, 2_000_000]
    log(f"{'zero copy (zc_points=100000000)':52s} {json.dumps(run('zc_points=100000000', sizes))}")
    log(f"{'zero copy, 2 parts':52s} {json.dumps(run('zc_points=100000000,zc_parts=2', sizes))}")
    log(f"{'copy engines (zc_points=0)':52s} {json.dumps(run('zc_points=0', sizes))}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep_single_scan.log"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
