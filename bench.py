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
  clocks        SM clock / throttle reasons sampled through NVML during the timed region
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
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
            worst = worst_ref = 0.0
            k = min(16, frames_avail)  # spot-check 16 frames of the benchmark data
            for f in range(k):
                got = gpu_out_host[f][::8, :3].astype(np.float64)
                ref = ob.deskew_xyzi_scan(pts_host[f][::8], eye, T_end[f], 0.0, 0.1, 0.05)
                worst = max(worst, float(np.abs(got - ref[:, :3]).max()))
                if use_ref:
                    ref2 = rb.deskew_xyzi_scan(pts_host[f][::8], eye, T_end[f], 0.0, 0.1, 0.05)
                    worst_ref = max(worst_ref, float(np.abs(got - ref2[:, :3]).max()))
            out["gpu_vs_oracle_max_abs_err_m"] = worst
            if use_ref:
                out["gpu_vs_reference_sources_max_abs_err_m"] = worst_ref
            out["gpu_vs_oracle_frames"] = k
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


def workload_name(args, world):
    if args.scaling == "weak":
        return f"{args.scans} synthetic {args.points}-pt HDL-64E scans per GPU, resident in HBM (BASELINE configs[2]; x{world} GPUs, independent shards)"
    return f"{args.scans} synthetic {args.points}-pt HDL-64E scans sharded over {world} GPU(s), resident in HBM (BASELINE configs[2]/[3])"


# ---------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, _, world = env_rank()
    if rank != 0:
        return 0
    from kitti_motion_compensation_b200 import capi
    n_sample = 64
    _, xi = capi.synth_frame_params(n_sample, SEED, 0, 0.5)
    pts = None
    try:
        import torch
        if torch.cuda.is_available():
            buf = torch.empty((n_sample * args.points, 4), dtype=torch.float32, device="cuda:0")
            capi.synth_scans_device(buf.data_ptr(), args.points, n_sample, args.rings, SEED, 0)
            torch.cuda.synchronize()
            pts = buf.cpu().numpy().reshape(n_sample, args.points, 4)
    except Exception:
        pts = None
    if pts is None:
        pts = numpy_scans(n_sample, args.points, args.rings, SEED)
    res = cpu_baseline_leg(pts, xi, args.points, None, args.cpu_seconds, steps=args.steps, warmup=args.warmup)
    sample = (f"each step = {res['per_step']} scans x {args.points} pts (one per host thread) of the same synthetic workload, "
              f"timed inside the C++ thread pool; {res['what']}")
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

    def step():
        capi.deskew_batch_device(d_in.data_ptr(), d_out.data_ptr(), d_off.data_ptr(), d_par.data_ptr(), my_scans, n_pts,
                                 capi.TIME_FROM_AZIMUTH, stream)

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    events = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches_before = capi.launch_count()
    barrier()
    sampler.start()
    events[0].record()
    for i in range(args.steps):
        step()
        events[i + 1].record()
    torch.cuda.synchronize()
    sampler.stop()
    barrier()
    launches = capi.launch_count() - launches_before
    elapsed_ms = events[0].elapsed_time(events[-1])
    step_ms = [events[i].elapsed_time(events[i + 1]) for i in range(args.steps)]
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    total_pts = torch.tensor([float(n_pts)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(total_pts, op=dist.ReduceOp.SUM)
    elapsed_ms_max = float(t.item())
    all_pts = float(total_pts.item())
    value = all_pts * args.steps / (elapsed_ms_max * 1e-3) / 1e6

    # ---- roofline of the dominant (only) kernel: one launch per step ---------------------------------------------------
    peak, peak_src = measured_peak()
    mean_launch_ms = statistics.fmean(step_ms)
    achieved = BYTES_PER_POINT * n_pts / (mean_launch_ms * 1e-3) / 1e9
    bpp = ncu_traffic_bytes_per_point()
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": (round(bpp * n_pts) if bpp is not None else None), "kernel": "DeskewBatchKernel",
                "algorithmic_bytes_per_launch": BYTES_PER_POINT * n_pts, "mean_launch_ms": round(mean_launch_ms, 4),
                "best_launch_ms": round(min(step_ms), 4), "peak_source": peak_src,
                "frac_of_8TBps_spec": round(achieved / 8000.0, 4)}

    # ---- end to end through the C ABI host entry point ---------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e_scans = min(args.e2e_scans, my_scans)
        e_pts = e_scans * points
        pin_in = torch.empty((e_pts, 4), dtype=torch.float32, pin_memory=True)
        pin_out = torch.empty((e_pts, 4), dtype=torch.float32, pin_memory=True)
        pin_in.copy_(d_in[:e_pts])
        torch.cuda.synchronize()
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
        te = torch.tensor([e_sec], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": round(world * e_pts * e_steps / float(te.item()) / 1e6, 2), "unit": UNIT,
               "h2d_bytes_per_step": e_pts * 16 + e_scans * 72 + 8, "d2h_bytes_per_step": e_pts * 16,
               "steps": e_steps, "scans_per_step_per_gpu": e_scans, "matches_resident_result": ok,
               "api": "kmc_b200_deskew_batch_host (pinned host in/out, 3-slot H2D/kernel/D2H pipeline)"}
        del pin_in, pin_out

    # ---- CPU baseline (rank 0, N = 1 only) --------------------------------------------------------------------------------
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        k = min(64, my_scans)
        pts_host = d_in[: k * points].cpu().numpy().reshape(k, points, 4)
        out_host = d_out[: k * points].cpu().numpy().reshape(k, points, 4)
        cpu = cpu_baseline_leg(pts_host, xi, points, out_host, args.cpu_seconds)

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
