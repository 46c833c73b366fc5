"""Shared builders for the tests and bench.py's parity spot-checks: real scan, synthetic scans, frames, closed form."""
from __future__ import annotations

import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

TOL_M = 1e-5  # BASELINE.json north_star: max |dxyz| < 1e-5 m per coordinate vs the reference's double result

# T_end = T_start * Exp(CONFIG1_TWIST) — 13 m/s, 0.5 rad/s, deliberately aggressive.
CONFIG1_TWIST = [1.34, 0.03, -0.01, -0.003, 0.004, 0.05]


def kats() -> dict:
    with open(os.path.join(GOLDEN, "reference_kats.json")) as f:
        return json.load(f)


def real_scan() -> np.ndarray:
    """The reference's shipped KITTI scan (drive_0005 frame 0): (123397, 4) float32 xyzi."""
    k = kats()["real_scan_frame0"]
    pts = np.fromfile(os.path.join(GOLDEN, k["file"]), dtype=np.float32).reshape(-1, 4)
    assert pts.shape[0] == k["num_points"]
    return pts


def oxts7(d: dict, stamp: float | None = None):
    return [d.get("stamp", 0.0) if stamp is None else stamp, d["lat"], d["lon"], d["alt"], d["roll"], d["pitch"], d["yaw"]]


def config1_frame():
    """(T_start, T_end, stamp_start, stamp_middle, stamp_end) for BASELINE config 1 (real scan, Mercator-magnitude pose)."""
    from oracle import binding as ob
    k = kats()
    r = k["real_scan_frame0"]
    T_start = ob.oxts_to_pose(oxts7(k["oxts_to_pose"]["oxts"], r["oxts_stamp"]))
    T_end = T_start @ ob.se3_exp(CONFIG1_TWIST)
    return T_start, T_end, r["stamp_start"], r["stamp_middle"], r["stamp_end"]


def synthetic_scan(n_points: int = 130_000, n_rings: int = 64, seed: int = 20110926, max_range: float = 120.0) -> np.ndarray:
    """HDL-64E style scan: ring-major, azimuth increasing within a ring, log-uniform range in
    [2, max_range) m, coordinates rounded to float32.  numpy generator — independent of the CUDA generator."""
    rng = np.random.default_rng(seed)
    steps = -(-n_points // n_rings)
    i = np.arange(n_points)
    ring, step = i // steps, i % steps
    el_top, el_bot = (2.0, -24.8) if n_rings == 64 else (15.0, -25.0)
    el = np.deg2rad(el_top + (el_bot - el_top) * ring / (n_rings - 1))
    az = 2 * np.pi * (step + rng.uniform(0, 1, n_points)) / steps
    r = 2.0 * np.exp(rng.uniform(0, 1, n_points) * np.log(max_range / 2.0))
    pts = np.empty((n_points, 4), dtype=np.float32)
    pts[:, 0] = r * np.cos(el) * np.cos(az)
    pts[:, 1] = r * np.cos(el) * np.sin(az)
    pts[:, 2] = r * np.sin(el)
    pts[:, 3] = rng.integers(0, 100, n_points) * 0.01
    return pts


def random_twist(rng) -> np.ndarray:
    """SURVEY 8d config 2 twist distribution: rho = (U(0,3), N(0,.05), N(0,.02)) m, phi = (N(0,.003), N(0,.004), N(0,.05)) rad."""
    return np.array([rng.uniform(0, 3), rng.normal(0, 0.05), rng.normal(0, 0.02),
                     rng.normal(0, 0.003), rng.normal(0, 0.004), rng.normal(0, 0.05)])


def random_pose(rng, mercator: bool = False) -> np.ndarray:
    from oracle import binding as ob
    T = ob.se3_exp(np.concatenate([rng.normal(0, 10, 3), rng.normal(0, 0.7, 3)]))
    if mercator:
        T[:3, 3] += np.array([937631.25, 6276764.0, 112.8])
    return T


def edge_points() -> np.ndarray:
    """SURVEY 8c edge cases: y = -0 / +0 with x < 0 (frac 1 / 0), x = y = 0, axis points, far points."""
    return np.array([
        [-17.173, -0.0, -1.81, 0.11],   # the real scan's own y == -0.0 point: frac == 1.0 exactly
        [-17.173, 0.0, -1.81, 0.12],    # frac == 0.0
        [0.0, 0.0, 1.5, 0.13],          # atan2(0, 0) = 0 -> frac 0.5, no NaN
        [-0.0, 0.0, 1.5, 0.14],         # atan2(+0, -0) = +pi -> frac 0
        [-0.0, -0.0, 1.5, 0.15],        # atan2(-0, -0) = -pi -> frac 1
        [0.0, -0.0, 1.5, 0.16],
        [5.0, 0.0, 0.0, 0.17], [0.0, 5.0, 0.0, 0.18], [0.0, -5.0, 0.0, 0.19], [-5.0, 1e-6, 0.0, 0.2], [-5.0, -1e-6, 0.0, 0.21],
        [3.0, 3.0, -1.0, 0.22], [-3.0, 3.0, -1.0, 0.23], [-3.0, -3.0, -1.0, 0.24], [3.0, -3.0, -1.0, 0.25],
        [119.9, 0.5, -20.0, 0.26], [-80.0, -90.0, 2.0, 0.27], [1e-3, -2e-3, 0.0, 0.28],
    ], dtype=np.float32)


def closed_form_deskew(xyzi: np.ndarray, xi: np.ndarray, x_req: float, frac: np.ndarray | None = None) -> np.ndarray:
    """Double-precision numpy evaluation of p' = Exp((frac - x_req) xi) p — an independent cross-check of the oracle
    (which follows the reference's literal GetPoseAtTime(t_req)^-1 GetPoseAtTime(t_i) product)."""
    p = xyzi[:, :3].astype(np.float64)
    if frac is None:
        frac = (np.pi - np.arctan2(xyzi[:, 1].astype(np.float64), xyzi[:, 0].astype(np.float64))) / (2 * np.pi)
    s = (frac - x_req)[:, None]
    rho, phi = np.asarray(xi[:3], float), np.asarray(xi[3:], float)
    th = np.linalg.norm(phi)
    if th < 1e-12:
        return p + s * rho + np.cross(np.broadcast_to(s * phi, p.shape), p)
    a = phi / th
    ang = s * th
    sn, cs = np.sin(ang), np.cos(ang)
    rot = cs * p + (1 - cs) * a * (p @ a)[:, None] + sn * np.cross(np.broadcast_to(a, p.shape), p)
    with np.errstate(divide="ignore", invalid="ignore"):
        k1 = np.where(np.abs(ang) > 1e-12, sn / ang, 1.0)
        k2 = np.where(np.abs(ang) > 1e-12, (1 - cs) / ang, 0.5 * ang)
    trans = s * (k1 * rho + (1 - k1) * a * (a @ rho) + k2 * np.cross(a, rho))
    return rot + trans


def max_abs_err(out_xyzi: np.ndarray, ref_xyz1: np.ndarray) -> float:
    return float(np.max(np.abs(out_xyzi[:, :3].astype(np.float64) - ref_xyz1[:, :3])))


# ---- KITTI run folders (the layout handlers.cpp:41-65 / data_io.cpp:18-66,140-166 read) -------------------------------------
def _clock(t: float) -> str:
    """seconds since midnight -> '2011-09-26 HH:MM:SS.nnnnnnnnn' (parsed by utils.cpp:34-41)."""
    h, rem = divmod(t, 3600.0)
    m, s = divmod(rem, 60.0)
    return f"2011-09-26 {int(h):02d}:{int(m):02d}:{s:012.9f}"


def make_run_folder(root: str, n_frames: int = 6, points: int = 20_000, seed: int = 1, yaw_rate: float = 0.3,
                    speed: float = 11.0, ragged: bool = True) -> dict:
    """Writes a synthetic KITTI raw run (oxts/, velodyne_points/{data,timestamps*.txt}) under `root` and returns what was
    written.  10 Hz scans, OxTS packets at the middle of each scan, a vehicle driving a curve (speed m/s, yaw_rate rad/s).
    Scans are chosen so that the reference's own range assert cannot fire (every pseudo stamp stays inside its scan)."""
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "oxts", "data"), exist_ok=True)
    os.makedirs(os.path.join(root, "velodyne_points", "data"), exist_ok=True)
    t0 = 13 * 3600 + 4 * 60 + 32.0
    starts, middles, ends, scans, oxts = [], [], [], [], []
    lat, lon, yaw = 49.011212804408, 8.4228850417969, -1.2219096732051
    for i in range(n_frames):
        starts.append(t0 + 0.1 * i + rng.uniform(0, 1e-3))
        ends.append(starts[-1] + 0.1033 + rng.uniform(-1e-3, 1e-3))
        middles.append(0.5 * (starts[-1] + ends[-1]) + rng.uniform(-1e-4, 1e-4))
        n = points + (int(rng.integers(-points // 10, points // 10)) if ragged else 0)
        pts = synthetic_scan(n, 64, seed * 1000 + i)
        # keep the reference away from its own abort: drop points whose pseudo stamp would round outside the scan
        frac = (np.pi - np.arctan2(pts[:, 1].astype(np.float64), pts[:, 0].astype(np.float64))) / (2 * np.pi)
        st = starts[-1] + frac * (ends[-1] - starts[-1])
        pts = pts[(st >= starts[-1]) & (st <= ends[-1])]
        scans.append(pts)
        pts.tofile(os.path.join(root, "velodyne_points", "data", f"{i:010d}.bin"))
        o = [lat, lon, 112.8 + 0.01 * i, 0.02 + 0.001 * i, 1e-3 * i, yaw]
        oxts.append(o)
        with open(os.path.join(root, "oxts", "data", f"{i:010d}.txt"), "w") as f:
            f.write(" ".join(f"{v:.13g}" for v in o) + " 0 0 " + f"{speed} 0 0 " + " ".join(["0"] * 14) + " 4 10 4 4 0\n")
        lat += speed * 0.1 * np.sin(yaw + np.pi / 2) / 111_000.0 * 0.5
        lon += speed * 0.1 * np.cos(yaw) / 73_000.0
        yaw += yaw_rate * 0.1
    for name, vals in (("velodyne_points/timestamps_start.txt", starts), ("velodyne_points/timestamps.txt", middles),
                       ("velodyne_points/timestamps_end.txt", ends), ("oxts/timestamps.txt", middles)):
        with open(os.path.join(root, name), "w") as f:
            f.write("\n".join(_clock(t) for t in vals) + "\n")
    return {"starts": starts, "middles": middles, "ends": ends, "scans": scans, "oxts": oxts}


def read_bin(path: str) -> np.ndarray:
    return np.fromfile(path, dtype=np.float32).reshape(-1, 4)


def write_calibration_folder(root: str, calib: dict) -> None:
    """calib_velo_to_cam.txt + calib_cam_to_cam.txt in the KITTI raw format (data_io.cpp:168-210, 321-406 read them):
    a header line (two for cam_to_cam), then `label: numbers` lines; each camera block is S K D R T S_rect R_rect P_rect."""
    def line(label, values):
        return f"{label}: " + " ".join(f"{v:.6e}" for v in values)
    os.makedirs(root, exist_ok=True)
    with open(os.path.join(root, "calib_velo_to_cam.txt"), "w") as f:
        f.write("calib_time: 15-Mar-2012 11:37:16\n" + line("R", calib["velo_to_cam"]["R"]) + "\n" + line("T", calib["velo_to_cam"]["T"])
                + "\ndelta_f: 0.000000e+00 0.000000e+00\ndelta_c: 0.000000e+00 0.000000e+00\n")
    with open(os.path.join(root, "calib_cam_to_cam.txt"), "w") as f:
        f.write("calib_time: 09-Jan-2012 13:57:47\ncorner_dist: 9.950000e-02\n")
        for k in ("00", "01", "02", "03"):
            f.write(line(f"S_{k}", [1392.0, 512.0]) + "\n" + line(f"K_{k}", [984.2439, 0, 690.0, 0, 980.8141, 233.1966, 0, 0, 1]) + "\n"
                    + line(f"D_{k}", [-0.3728755, 0.2037299, 0.002219027, 0.001383707, -0.07233722]) + "\n"
                    + line(f"R_{k}", [1, 0, 0, 0, 1, 0, 0, 0, 1]) + "\n" + line(f"T_{k}", [-0.5 * int(k), 0, 0]) + "\n"
                    + line(f"S_rect_{k}", calib["S_rect_00"]) + "\n" + line(f"R_rect_{k}", calib["R_rect_00"]) + "\n"
                    + line(f"P_rect_{k}", calib["P_rect"][k]) + "\n")
