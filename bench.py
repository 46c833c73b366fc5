#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native LiDAR deskew path.

Metric (BASELINE.json): Mpoints/s deskewed at 1/2/4/8 B200 and achieved HBM GB/s against the roofline (32 B/point).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm on the host cores (oracle/_ref, else the oracle port)

A "step" is one pass of the fused deskew kernel over one batch of synthetic 130 000-point HDL-64E scans that is
already resident in HBM (BASELINE config 3: 10 000 scans = 1.3e9 points = 41.6 GB of traffic per step, far larger than
the 126 MB L2, so no L2 flush is needed between steps).  With N GPUs every rank owns its own 10 000 independent scans
(weak scaling, no collective on the data path); `--scaling strong` shards one 10 000-scan batch instead (config 4).

One JSON line is printed by rank 0.  Besides the contract keys it carries
  roofline      achieved algorithmic GB/s of the deskew kernel (32 B/point x points per launch / mean launch time,
                CUDA events on the launching stream) against the measured HBM copy peak of MEASURED_PEAKS.json
  cpu_baseline  the reference algorithm (oracle/_ref when built, else the oracle port) timed on this box's host cores on a bounded sample of the
                same scans (rank 0, N=1 only) — a reported baseline, plus the max |dxyz| of the GPU result on them
  e2e           the same metric through the C ABI's host entry point (kmc_b200_deskew_batch_host): pinned host buffers,
                H2D + kernel + D2H inside the timed region
                + e2e.roofline: the same bytes as plain concurrent H2D + D2H copies on every rank at the same time (the
                box's ceiling for any host-buffer path) and the fraction of it the pipeline reaches
  e2e_dropin    (N = 1) kmc::MotionCompensateFrame(Frame const&, Time) — the reference's own signature, double
                column-major cloud in pageable memory — on the real KITTI scan, with the fraction of the link it uses
  strong        (N > 1) BASELINE configs[3]: ONE batch of --scans scans sharded over the N GPUs with
                kmc_b200_shard_range, timed the same way, plus shards_bit_equal: every frame's checksum on its shard GPU
                equals the checksum of the same frame computed by rank 0 alone over the whole batch
  e2e_inprocess (N > 1) the single-process API kmc_b200_deskew_batch_multi_gpu on all N devices (rank 0, others idle)
  clocks        SM clock / throttle reasons sampled through NVML during the timed region
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POINTS_PER_SCAN = 130_000
SCANS = 10_000
SEED = 20110926
BYTES_PER_POINT = 32  # 16 B float4 xyzi read + 16 B written
METRIC = "Mpoints/s deskewed"
UNIT = "Mpoints/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--scans", type=int, default=SCANS, help="scans per GPU (weak) or in total (strong)")
    ap.add_argument("--points", type=int, default=POINTS_PER_SCAN, help="points per scan")
    ap.add_argument("--rings", type=int, default=64)
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    ap.add_argument("--e2e-scans", type=int, default=1000, help="scans per end-to-end (host buffer) step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="target wall time of the all-core CPU baseline sample")
    return ap.parse_args()


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy_ read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s; MEASURED_PEAKS.json absent)"


def ncu_traffic_bytes_per_point():
    """dram bytes per point of the deskew kernel from the committed ncu capture (profiles/roofline_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_point"])
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
    NOTED = {"sw_power_cap": 0x4}

    def __init__(self, device_index: int, period_s: float = 0.01):
        self.period = period_s
        self.samples = []
        self.reason_bits = 0
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = device_index
            if visible:
                ids = [v for v in visible.split(",") if v.strip() != ""]
                if device_index < len(ids) and ids[device_index].strip().isdigit():
                    phys = int(ids[device_index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception as exc:  # pragma: no cover - depends on the box
            self.nv = None
            self.error = repr(exc)

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append(mhz)
                self.reason_bits |= int(reasons)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        reasons = [name for name, bit in {**self.BAD, **self.NOTED}.items() if self.reason_bits & bit]
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------
def cpu_baseline_leg(pts_host: np.ndarray, xi: np.ndarray, points: int, gpu_out_host: np.ndarray | None, target_seconds: float,
                     steps: int | None = None, warmup: int = 0):
    """Times the reference's CPU algorithm for the path (per-point Log/inverse recomputation included) on the host cores:
    oracle/_ref (the reference's own four source files, compiled by `make -C oracle ref`) when that library is present,
    else the oracle port.  This is the one place bench.py executes oracle/.  pts_host: (frames, points, 4) float32."""
    from oracle import binding as ob
    from oracle import ref_binding as rb
    ob.build()
    use_ref = rb.available()
    engine = rb if use_ref else ob
    if use_ref:
        kind = "reference"
        what = ("oracle/_ref/libkmc_ref.so = the reference's own motion_compensation / trajectory_interpolation / lie_algebra / "
                "timestamp_mocking .cpp compiled unmodified (-O3), Eigen supplied by "
                + ("Eigen 3" if rb.eigen_provider() == "eigen3" else "this repo's eigen_shim.hpp (no Eigen 3 in the image)"))
    else:
        kind = "port"
        what = "oracle/kmc_oracle.cpp (-O3, double, reference's per-point Log/SVD/inverse kept)"
    frames_avail = pts_host.shape[0]
    cores = max(1, ob.hardware_threads())
    eye = np.eye(4)
    T_end = [ob.se3_exp(x) for x in xi[:frames_avail]]
    stamps = [[0.0, 0.1, 0.05]] * frames_avail

    def prepare(n_frames):
        idx = [i % frames_avail for i in range(n_frames)]
        return np.ascontiguousarray(pts_host[idx]), [eye] * n_frames, [T_end[i] for i in idx], [stamps[i] for i in idx]

    def run(prepared, threads, eng=None):
        block, ts, te, st = prepared
        sec, _ = (eng or engine).timed_frames(block, points, ts, te, st, threads)  # seconds inside the C++ thread pool only
        return len(ts) * points / sec / 1e6, sec

    one = prepare(1)
    single_mpts, t1 = run(one, 1)
    if steps is None:  # cpu_baseline object of the b200 arm: one bounded all-core sample
        n_frames = int(min(max(cores, cores * target_seconds / max(t1, 1e-3)), 4096))
        all_mpts, sec = run(prepare(n_frames), cores)
        out = {"value": round(all_mpts, 4), "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{n_frames} scans x {points} pts of the same synthetic workload, one scan per task over {cores} "
                         f"std::threads, {sec:.1f} s wall; {what}",
               "single_thread_value": round(single_mpts, 4)}
        if use_ref:  # the restatement beside it, same single-thread sample
            out["oracle_port_single_thread_value"] = round(run(one, 1, ob)[0], 4)
        if gpu_out_host is not None:
            # SURVEY 8d config 3 spot check, widened: 64 WHOLE frames of the benchmark data against the oracle port and the
            # compiled reference sources, one frame per host thread (ctypes releases the GIL inside the C++ call)
            from concurrent.futures import ThreadPoolExecutor
            k = min(64, frames_avail, gpu_out_host.shape[0])

            def frame_err(f):
                got = gpu_out_host[f][:, :3].astype(np.float64)
                e1 = float(np.abs(got - ob.deskew_xyzi_scan(pts_host[f], eye, T_end[f], 0.0, 0.1, 0.05)[:, :3]).max())
                e2 = float(np.abs(got - rb.deskew_xyzi_scan(pts_host[f], eye, T_end[f], 0.0, 0.1, 0.05)[:, :3]).max()) if use_ref else 0.0
                same_w = bool(np.array_equal(gpu_out_host[f][:, 3], pts_host[f][:, 3]))
                return e1, e2, same_w

            with ThreadPoolExecutor(max_workers=cores) as pool:
                errs = list(pool.map(frame_err, range(k)))
            out["gpu_vs_oracle_max_abs_err_m"] = max(e[0] for e in errs)
            if use_ref:
                out["gpu_vs_reference_sources_max_abs_err_m"] = max(e[1] for e in errs)
            out["gpu_vs_oracle_frames"] = k
            out["gpu_vs_oracle_points_per_frame"] = points
            out["intensity_bit_exact"] = all(e[2] for e in errs)
        return out
    # reference arm: `steps` timed steps, each a bounded sample of `cores` scans spread over all host threads
    per_step = cores
    prepared = prepare(per_step)
    for _ in range(warmup):
        run(prepared, cores)
    elapsed = 0.0
    for _ in range(steps):
        elapsed += run(prepared, cores)[1]
    res = {"value": per_step * points * steps / elapsed / 1e6, "elapsed": elapsed, "cores": cores, "per_step": per_step,
           "single_thread_value": single_mpts, "kind": kind, "what": what}
    if use_ref:
        res["oracle_port_value"] = run(prepared, cores, ob)[0]
    return res


def numpy_scans(n_scans: int, points: int, rings: int, seed: int) -> np.ndarray:
    """Host-side generator with the bench workload's distribution (used by the reference arm when no GPU is visible)."""
    out = np.empty((n_scans, points, 4), dtype=np.float32)
    steps = -(-points // rings)
    i = np.arange(points)
    ring, step = i // steps, i % steps
    el_top, el_bot = (2.0, -24.8) if rings == 64 else (15.0, -25.0)
    el = np.deg2rad(el_top + (el_bot - el_top) * ring / (rings - 1))
    for k in range(n_scans):
        rng = np.random.default_rng(seed + k)
        az = 2 * np.pi * (step + rng.uniform(0, 1, points)) / steps
        r = 2.0 * np.exp(rng.uniform(0, 1, points) * np.log(60.0))
        out[k, :, 0] = r * np.cos(el) * np.cos(az)
        out[k, :, 1] = r * np.cos(el) * np.sin(az)
        out[k, :, 2] = r * np.sin(el)
        out[k, :, 3] = rng.integers(0, 100, points) * 0.01
    return out


def numpy_twists(n_frames: int, seed: int, first_scan_index: int = 0) -> np.ndarray:
    """The twists of kmc_b200_synth_frame_params (csrc/kmc_capi.cu: SplitMix64 counter generator, Box-Muller) restated in
    Python, so that the reference arm draws the same motions without loading this repository's CUDA library."""
    mask = (1 << 64) - 1

    def mix(z):
        z = (z + 0x9E3779B97F4A7C15) & mask
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        return z ^ (z >> 31)

    out = np.zeros((n_frames, 6))
    for k in range(n_frames):
        state = [mix((seed + first_scan_index + k) & mask) ^ 0x5DEECE66D]

        def uniform():
            state[0] = mix(state[0])
            return (state[0] >> 11) * (1.0 / 9007199254740992.0)

        def normal():
            u1, u2 = uniform(), uniform()
            return math.sqrt(-2.0 * math.log(max(u1, 1e-300))) * math.cos(2.0 * math.pi * u2)

        out[k] = [3.0 * uniform(), 0.05 * normal(), 0.02 * normal(), 0.003 * normal(), 0.004 * normal(), 0.05 * normal()]
    return out


def workload_name(args, world):
    if args.scaling == "weak":
        return f"{args.scans} synthetic {args.points}-pt HDL-64E scans per GPU, resident in HBM (BASELINE configs[2]; x{world} GPUs, independent shards)"
    return f"{args.scans} synthetic {args.points}-pt HDL-64E scans sharded over {world} GPU(s), resident in HBM (BASELINE configs[2]/[3])"


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference arm is hermetic: numpy generates the scans and the twists, oracle/_ref (or the oracle port) does the
    timed work; this repository's CUDA library is never loaded into the process."""
    rank, _, world = env_rank()
    if rank != 0:
        return 0
    n_sample = 64
    xi = numpy_twists(n_sample, SEED, 0)
    pts = numpy_scans(n_sample, args.points, args.rings, SEED)
    res = cpu_baseline_leg(pts, xi, args.points, None, args.cpu_seconds, steps=args.steps, warmup=args.warmup)
    sample = (f"each step = {res['per_step']} scans x {args.points} pts (one per host thread) of the same synthetic workload "
              f"(numpy generator, same distribution and twists as the CUDA generator), timed inside the C++ thread pool; {res['what']}")
    cpu = {"value": round(res["value"], 4), "unit": UNIT, "cores": res["cores"], "kind": res["kind"], "sample": sample,
           "single_thread_value": round(res["single_thread_value"], 4)}
    if "oracle_port_value" in res:
        cpu["oracle_port_value"] = round(res["oracle_port_value"], 4)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(res["value"], 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * res["elapsed"] / max(args.steps, 1), 3),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "points_per_scan": args.points, "sample": sample},
        "cpu_baseline": cpu,
        "e2e": {"value": round(res["value"], 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def shards_bit_equal(gathered, whole, n_frames, world, shard_range):
    """The cross-GPU bit-equality check of the strong-scaling leg: gathered[r] holds rank r's per-frame checksums of ITS shard
    (padded to a common length for all_gather), `whole` the checksums of the single-GPU result over the whole batch; shard r
    owns frames shard_range(n_frames, world, r).  Returns (all equal, frames compared)."""
    import torch
    equal, compared = True, 0
    for r in range(world):
        b, e = shard_range(n_frames, world, r)
        equal = equal and bool(torch.equal(gathered[r][: e - b], whole[b:e]))
        compared += e - b
    return equal, compared


def copy_ceiling(torch, device, pin_in, pin_out, barrier, reps=6):
    """Plain concurrent H2D + D2H of the e2e step's bytes (one cudaMemcpyAsync each way per rep on two streams): what the
    box gives any host-buffer path when every rank uses its link at the same time.  Returns seconds for `reps` reps."""
    d_a = torch.empty(pin_in.shape, dtype=pin_in.dtype, device=device)
    d_b = torch.empty(pin_out.shape, dtype=pin_out.dtype, device=device)
    up, down = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def rep():
        with torch.cuda.stream(up):
            d_a.copy_(pin_in, non_blocking=True)
        with torch.cuda.stream(down):
            pin_out.copy_(d_b, non_blocking=True)

    for _ in range(2):
        rep()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        rep()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    barrier()
    del d_a, d_b
    return sec, reps


def dropin_leg(capi, torch, ceiling_gbs_each_way):
    """kmc::MotionCompensateFrame(Frame const&, Time) on the real KITTI scan (BASELINE configs[0]): the C++ call through
    libkitti_motion_compensation_lib.so (separate process: lib/bench_motion_compensate_frame) and the C ABI call beneath it
    (kmc_b200_deskew_cloud_f64_host) from numpy's pageable memory, checked against the double-precision closed form."""
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    # inputs and the check are built WITHOUT the oracle (which bench.py runs only inside cpu_baseline_leg): poses from the
    # product's own host math, stamps from numpy's atan2, parity against the double-precision numpy closed form of tests/helpers.py
    k = helpers.kats()
    pts = helpers.real_scan()
    n = len(pts)
    r0, o = k["real_scan_frame0"], k["oxts_to_pose"]["oxts"]
    t0, t1, t2 = r0["stamp_start"], r0["stamp_middle"], r0["stamp_end"]
    T_start = capi.oxts_to_pose(o["lat"], o["lon"], o["alt"], o["roll"], o["pitch"], o["yaw"])
    T_end = T_start @ capi.se3_exp(helpers.CONFIG1_TWIST)
    p = capi.frame_params_from_poses(T_start, T_end, t0, t2, t1)
    cloud = np.concatenate([pts[:, :3].astype(np.float64), np.ones((n, 1))], axis=1)
    frac = (np.pi - np.arctan2(cloud[:, 1], cloud[:, 0])) / (2 * np.pi)  # timestamp_mocking.cpp:46
    stamps = t0 + frac * (t2 - t0)
    cm = np.ascontiguousarray(cloud.T)
    out64 = np.empty_like(cm)
    dp = C.POINTER(C.c_double)
    out = {"api": "kmc::MotionCompensateFrame(Frame const&, Time) -> kmc_b200_deskew_cloud_f64_host", "points": n,
           "workload": "real KITTI scan 2011_09_26_drive_0005 frame 0, Mercator-magnitude start pose, 13 m/s + 0.5 rad/s (BASELINE configs[0])",
           "link_bytes_per_point": 28, "h2d_bytes_per_call": 16 * n, "d2h_bytes_per_call": 12 * n,
           "host_memory": "pageable (numpy / Eigen buffers), column-major double in and out"}
    with capi.Handle(0, 250_000) as h:
        def call():
            return capi.lib().kmc_b200_deskew_cloud_f64_host(h.raw, cm.ctypes.data_as(dp), stamps.ctypes.data_as(dp), out64.ctypes.data_as(dp),
                                                             n, t0, t2, t1, C.byref(p), None)
        for _ in range(10):
            assert call() == 0
        t = []
        for _ in range(200):
            a = time.perf_counter()
            call()
            t.append(time.perf_counter() - a)
    med = statistics.median(t)
    closed = helpers.closed_form_deskew(pts, np.array(helpers.CONFIG1_TWIST), (t1 - t0) / (t2 - t0), frac=(stamps - t0) / (t2 - t0))
    out["c_abi_us_median"] = round(med * 1e6, 2)
    out["c_abi_us_min"] = round(min(t) * 1e6, 2)
    out["c_abi_mpoints_per_s"] = round(n / med / 1e6, 1)
    out["max_abs_err_m_vs_closed_form_f64"] = float(np.abs(out64.T[:, :3] - closed).max())
    out["check"] = ("every point against the double-precision closed form Exp((x_i - x_req) xi) p (numpy, tests/helpers.py); the oracle and the "
                    "compiled reference sources check this entry point in tests/ (2e-7 m)")
    binary = os.path.join(ROOT, "kitti_motion_compensation_b200", "lib", "bench_motion_compensate_frame")
    scan = os.path.join(ROOT, "tests", "golden", helpers.kats()["real_scan_frame0"]["file"])
    if os.path.exists(binary):
        for threads, key in ((1, "cpp"), (4, "cpp_4_concurrent_callers")):
            try:
                r = subprocess.run([binary, scan, "200", str(threads)], capture_output=True, text=True, timeout=120, check=True)
                out[key] = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception as exc:  # pragma: no cover - depends on the box
                out[key] = {"error": repr(exc)[:200]}
    cpp = out.get("cpp", {})
    us = cpp.get("us_median", out["c_abi_us_median"])
    out["value"] = round(n / us, 1)  # Mpoints/s of the C++ call (the C ABI call when the binary is absent)
    out["unit"] = UNIT
    if ceiling_gbs_each_way:
        # the link moves 16 B/point up and 12 B/point down concurrently; at the box's concurrent ceiling the call could not
        # take less than 16 n / ceiling
        out["pcie"] = {"ceiling_gbs_each_way": round(ceiling_gbs_each_way, 2), "link_gbs_both_ways": round(28 * n / (us * 1e-6) / 1e9, 2),
                       "frac": round((16 * n / (ceiling_gbs_each_way * 1e9)) / (us * 1e-6), 4)}
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    from kitti_motion_compensation_b200 import build, capi

    rank, local_rank, world = env_rank()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the deskew path has no CPU fallback (use --impl reference for the CPU baseline)")
    if not os.path.exists(capi.LIB_PATH):
        build.build()
    capi.lib()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- this rank's shard: independent scans, no data-path collective ------------------------------------------------
    if args.scaling == "weak":
        my_scans, first_scan = args.scans, rank * args.scans
    else:
        b, e = capi.shard_range(args.scans, world, rank)
        my_scans, first_scan = e - b, b
    points = args.points
    free_b, _ = torch.cuda.mem_get_info()
    need = 2 * my_scans * points * 16
    shrunk = False
    if need > 0.9 * free_b:
        my_scans = int(0.9 * free_b // (2 * points * 16))
        shrunk = True
    n_pts = my_scans * points
    d_in = torch.empty((n_pts, 4), dtype=torch.float32, device=device)
    d_out = torch.empty_like(d_in)
    d_off = torch.arange(0, (my_scans + 1) * points, points, dtype=torch.int64, device=device)
    params, xi = capi.synth_frame_params(my_scans, SEED, first_scan, 0.5)
    d_par = torch.from_numpy(params.view(np.uint8).copy()).to(device)
    stream = torch.cuda.current_stream().cuda_stream
    capi.synth_scans_device(d_in.data_ptr(), points, my_scans, args.rings, SEED, first_scan, stream)
    torch.cuda.synchronize()

    def timed_steps(step, steps, sampler=None):
        """W warm-up steps, then `steps` timed ones: CUDA events on the launching stream, barrier + synchronize both sides."""
        for _ in range(max(args.warmup, 3)):
            step()
        events = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        if sampler:
            sampler.start()
        events[0].record()
        for i in range(steps):
            step()
            events[i + 1].record()
        torch.cuda.synchronize()
        if sampler:
            sampler.stop()
        barrier()
        return events[0].elapsed_time(events[-1]), [events[i].elapsed_time(events[i + 1]) for i in range(steps)]

    def step():
        capi.deskew_batch_device(d_in.data_ptr(), d_out.data_ptr(), d_off.data_ptr(), d_par.data_ptr(), my_scans, n_pts,
                                 capi.TIME_FROM_AZIMUTH, stream)

    sampler = ClockSampler(local_rank)
    launches_before = capi.launch_count()
    elapsed_ms, step_ms = timed_steps(step, args.steps, sampler)
    launches = capi.launch_count() - launches_before - max(args.warmup, 3)
    total_pts = torch.tensor([float(n_pts)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_pts, op=dist.ReduceOp.SUM)
    elapsed_ms_max = max_over_ranks(elapsed_ms)
    all_pts = float(total_pts.item())
    value = all_pts * args.steps / (elapsed_ms_max * 1e-3) / 1e6

    # ---- roofline of the dominant (only) kernel: one launch per step ---------------------------------------------------
    peak, peak_src = measured_peak()
    mean_launch_ms = statistics.fmean(step_ms)
    achieved = BYTES_PER_POINT * n_pts / (mean_launch_ms * 1e-3) / 1e9
    bpp = ncu_traffic_bytes_per_point()
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": (round(bpp * n_pts) if bpp is not None else None),
                "traffic_source": "ncu capture x points: dram__bytes_read+write per point of DeskewBatchKernel from the committed "
                                  "ncu --set full capture (profiles/roofline_traffic.json) times the points of one launch; not "
                                  "re-measured in this run",
                "kernel": "DeskewBatchKernel",
                "algorithmic_bytes_per_launch": BYTES_PER_POINT * n_pts, "mean_launch_ms": round(mean_launch_ms, 4),
                "best_launch_ms": round(min(step_ms), 4), "peak_source": peak_src,
                "frac_of_8TBps_spec": round(achieved / 8000.0, 4)}

    # ---- strong scaling (BASELINE configs[3]): ONE batch of args.scans scans sharded over the ranks --------------------
    strong = None
    if world > 1 and args.scaling == "weak" and not shrunk:
        # rank 0's weak shard is scans [0, args.scans) = the whole strong batch, deskewed by ONE GPU in one launch: its
        # per-frame checksums are the G = 1 result every shard is compared with
        d_sums = torch.zeros(my_scans, dtype=torch.int64, device=device)
        if rank == 0:
            capi.frame_checksums_device(d_out.data_ptr(), d_off.data_ptr(), my_scans, n_pts, d_sums.data_ptr(), stream)
            torch.cuda.synchronize()
            sums_g1 = d_sums.clone()
        sb, se = capi.shard_range(args.scans, world, rank)
        s_scans, s_pts = se - sb, (se - sb) * points
        if rank != 0:  # rank 0's shard [0, se) is already in place (same seeds, same records)
            capi.synth_scans_device(d_in.data_ptr(), points, s_scans, args.rings, SEED, sb, stream)
            s_params, _ = capi.synth_frame_params(s_scans, SEED, sb, 0.5)
            d_par[: s_scans * 64].copy_(torch.from_numpy(s_params.view(np.uint8).copy()).to(device))
        d_out[:s_pts].zero_()
        torch.cuda.synchronize()

        def strong_step():
            capi.deskew_batch_device(d_in.data_ptr(), d_out.data_ptr(), d_off.data_ptr(), d_par.data_ptr(), s_scans, s_pts,
                                     capi.TIME_FROM_AZIMUTH, stream)

        s_elapsed_ms, s_step_ms = timed_steps(strong_step, args.steps)
        s_elapsed_max = max_over_ranks(s_elapsed_ms)
        # every rank's per-frame checksums travel to rank 0 (64 bits per frame: verification, not data path)
        cap = -(-args.scans // world)
        mine = torch.zeros(cap, dtype=torch.int64, device=device)
        capi.frame_checksums_device(d_out.data_ptr(), d_off.data_ptr(), s_scans, s_pts, mine.data_ptr(), stream)
        torch.cuda.synchronize()
        gathered = [torch.zeros(cap, dtype=torch.int64, device=device) for _ in range(world)]
        dist.all_gather(gathered, mine)
        strong_value = args.scans * points * args.steps / (s_elapsed_max * 1e-3) / 1e6
        if rank == 0:
            equal, compared = shards_bit_equal(gathered, sums_g1, args.scans, world, capi.shard_range)
            nonzero = bool((sums_g1 != 0).all().item())
            one_gpu_ms = elapsed_ms_max / args.steps  # one GPU over the whole batch (the weak leg's step, max over ranks)
            strong = {"value": round(strong_value, 1), "unit": UNIT, "ms_per_step": round(s_elapsed_max / args.steps, 4),
                      "speedup_vs_n1": round(one_gpu_ms / (s_elapsed_max / args.steps), 3),
                      "n1_ms_per_step": round(one_gpu_ms, 4),
                      "n1_source": "this run's weak leg: one GPU deskewing the same 10 000-scan batch in one launch (max over ranks)",
                      "scans_total": args.scans, "scans_per_gpu": s_scans, "shards_bit_equal": bool(equal and nonzero),
                      "frames_compared": compared,
                      "check": "per-frame 64-bit position-weighted checksums (kmc_b200_frame_checksums_device) of every shard on its "
                               "own GPU vs the checksums of rank 0's single-GPU result over the whole batch",
                      "per_gpu_gbs": round(BYTES_PER_POINT * s_pts / (statistics.fmean(s_step_ms) * 1e-3) / 1e9, 1),
                      "workload": f"{args.scans} synthetic {points}-pt HDL-64E scans sharded over {world} GPUs (BASELINE configs[3])"}
        # put the weak shard's output back so that the legs below see the weak result
        if rank != 0:
            capi.synth_scans_device(d_in.data_ptr(), points, my_scans, args.rings, SEED, first_scan, stream)
            d_par.copy_(torch.from_numpy(params.view(np.uint8).copy()).to(device))
        step()
        torch.cuda.synchronize()

    # ---- end to end through the C ABI host entry point ---------------------------------------------------------------
    e2e = None
    e2e_inprocess = None
    ceiling_each_way = None
    if not args.no_e2e:
        e_scans = min(args.e2e_scans, my_scans)
        e_pts = e_scans * points
        pin_in = torch.empty((e_pts, 4), dtype=torch.float32, pin_memory=True)
        pin_out = torch.empty((e_pts, 4), dtype=torch.float32, pin_memory=True)
        pin_in.copy_(d_in[:e_pts])
        torch.cuda.synchronize()
        # the ceiling first: plain copies of the same bytes, all ranks at once
        c_sec, c_reps = copy_ceiling(torch, device, pin_in, pin_out, barrier)
        c_sec = max_over_ranks(c_sec)
        ceiling_each_way = world * e_pts * 16 * c_reps / c_sec / 1e9
        offs = np.arange(0, (e_scans + 1) * points, points, dtype=np.int64)
        with capi.Handle(local_rank, 32 * points) as h:
            for _ in range(max(2, min(args.warmup, 3))):
                h.deskew_batch_ptr(pin_in.data_ptr(), pin_out.data_ptr(), offs, params[:e_scans])
            e_steps = max(3, min(args.steps, 20))
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                h.deskew_batch_ptr(pin_in.data_ptr(), pin_out.data_ptr(), offs, params[:e_scans])
            e_sec = time.perf_counter() - t0
            barrier()
        ok = bool(torch.equal(pin_out[: 4 * points].to(device), d_out[: 4 * points]))
        e_sec = max_over_ranks(e_sec)
        achieved_each_way = world * e_pts * 16 * e_steps / e_sec / 1e9
        e2e = {"value": round(world * e_pts * e_steps / e_sec / 1e6, 2), "unit": UNIT,
               "h2d_bytes_per_step": e_pts * 16 + e_scans * 72 + 8, "d2h_bytes_per_step": e_pts * 16,
               "steps": e_steps, "scans_per_step_per_gpu": e_scans, "matches_resident_result": ok,
               "api": "kmc_b200_deskew_batch_host (pinned host in/out, 3-slot H2D/kernel/D2H pipeline)",
               "roofline": {"bound": "pcie / host memory", "achieved_gbs_each_way": round(achieved_each_way, 2),
                            "ceiling_gbs": round(ceiling_each_way, 2), "frac": round(achieved_each_way / ceiling_each_way, 4),
                            "ceiling_source": f"plain cudaMemcpyAsync of the same {e_pts * 16 / 1e9:.2f} GB per direction per rank, H2D and D2H "
                                              f"concurrently on two streams, all {world} rank(s) at once, {c_reps} reps, wall clock, max over ranks"}}
        # ---- the single-process API on all N devices: kmc_b200_deskew_batch_multi_gpu (rank 0; the other ranks idle) ----
        if world > 1:
            barrier()
            if rank == 0 and torch.cuda.device_count() >= world:
                handles = [capi.Handle(d, 32 * points) for d in range(world)]
                try:
                    arr = (capi._vp * world)(*[hh.raw for hh in handles])

                    def multi():
                        capi.check(capi.lib().kmc_b200_deskew_batch_multi_gpu(arr, world, pin_in.data_ptr(), pin_out.data_ptr(), offs.ctypes.data,
                                                                              params[:e_scans].ctypes.data, e_scans, capi.TIME_FROM_AZIMUTH))
                    pin_out.zero_()
                    for _ in range(2):
                        multi()
                    i_steps = max(3, min(args.steps, 10))
                    t0 = time.perf_counter()
                    for _ in range(i_steps):
                        multi()
                    i_sec = time.perf_counter() - t0
                    torch.cuda.set_device(local_rank)
                    same = bool(torch.equal(pin_out.to(device), d_out[:e_pts]))
                    e2e_inprocess = {"value": round(e_pts * i_steps / i_sec / 1e6, 2), "unit": UNIT, "devices": world,
                                     "scans_per_step_total": e_scans, "steps": i_steps, "gbs_each_way": round(e_pts * 16 * i_steps / i_sec / 1e9, 2),
                                     "matches_single_gpu_resident_result": same,
                                     "api": "kmc_b200_deskew_batch_multi_gpu: one process, one handle + host thread per device, one pinned host batch"}
                finally:
                    for hh in handles:
                        hh.close()
                    torch.cuda.set_device(local_rank)
            barrier()
        del pin_in, pin_out

    # ---- CPU baseline + the drop-in call (rank 0, N = 1 only) -------------------------------------------------------------
    cpu = None
    e2e_dropin = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        k = min(64, my_scans)
        pts_host = d_in[: k * points].cpu().numpy().reshape(k, points, 4)
        out_host = d_out[: k * points].cpu().numpy().reshape(k, points, 4)
        cpu = cpu_baseline_leg(pts_host, xi, points, out_host, args.cpu_seconds)
    if world == 1 and rank == 0 and not args.no_e2e:
        try:
            e2e_dropin = dropin_leg(capi, torch, ceiling_each_way)
        except Exception as exc:  # the headline line must still print
            e2e_dropin = {"error": repr(exc)[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(elapsed_ms_max / args.steps, 4), "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "scans_per_gpu": my_scans, "points_per_scan": points,
                       "points_per_step_all_gpus": int(all_pts), "bytes_per_point": BYTES_PER_POINT, "seed": SEED,
                       "l2": f"inputs ({n_pts * 16 / 1e9:.1f} GB in + {n_pts * 16 / 1e9:.1f} GB out per GPU) far exceed the 126 MB L2; no flush between steps needed",
                       "timing": "CUDA events on the launching stream, barrier+synchronize both sides, max over ranks",
                       "parallelism": f"frame-sharded x{world}, no collective", "shrunk_to_fit": shrunk,
                       "tune": os.environ.get("KMC_B200_TUNE", "default")},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": sampler.summary(), "gpu_launches": int(launches),
        }
        if strong is not None:
            line["strong"] = strong
        if e2e_inprocess is not None:
            line["e2e_inprocess"] = e2e_inprocess
        if e2e_dropin is not None:
            line["e2e_dropin"] = e2e_dropin
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
