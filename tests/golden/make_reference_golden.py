"""Writes tests/golden/reference_sources_golden.npz from oracle/_ref — the reference's OWN sources compiled unmodified
(`make -C oracle ref`; needs /root/reference, so it runs in the development container only).  The fixture lets the oracle
and the CUDA path be checked against outputs of the reference's code even where oracle/_ref is not available.

    python tests/golden/make_reference_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import binding as ob  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402
import helpers  # noqa: E402

STRIDE = 61  # every 61st point of the real scan: 2 023 points


def main():
    if not rb.available():
        raise SystemExit("oracle/_ref/libkmc_ref.so is not built: run `make -C oracle ref` where /root/reference exists")
    out = {"stride": np.array(STRIDE), "eigen_provider": np.array(rb.eigen_provider())}
    # BASELINE config 1: the shipped scan, Mercator-magnitude pose, 13 m/s + 0.5 rad/s, three requested times
    pts = helpers.real_scan()
    T_start, T_end, t0, t1, t2 = helpers.config1_frame()
    for name, t_req in (("middle", t1), ("start", t0), ("end", t2)):
        out[f"config1_{name}"] = rb.deskew_xyzi_scan(pts, T_start, T_end, t0, t2, t_req)[::STRIDE, :3]
    # a synthetic scan under a fast rotating motion from a random (non-Mercator) pose, requested at 30 % of the scan
    rng = np.random.default_rng(2011)
    syn = helpers.synthetic_scan(20_000, 64, 926)
    P1 = helpers.random_pose(rng)
    xi = np.array([2.1, -0.07, 0.03, 0.006, -0.009, 0.11])
    out["synthetic_pose_start"] = P1
    out["synthetic_twist"] = xi
    out["synthetic_out"] = rb.deskew_xyzi_scan(syn, P1, P1 @ ob.se3_exp(xi), 0.0, 0.1, 0.03)[::5, :3]
    # the four cameras' draw lists of the raw real scan (camera_model.cpp through the recording cv::circle)
    with open(os.path.join(HERE, "kitti_calibration_2011_09_26.json")) as f:
        c = json.load(f)
    T = np.eye(4)
    T[:3, :3] = np.array(c["velo_to_cam"]["R"]).reshape(3, 3)
    T[:3, 3] = c["velo_to_cam"]["T"]
    R_rect = np.array(c["R_rect_00"]).reshape(3, 3)
    P = [np.array(c["P_rect"][k]).reshape(3, 4) for k in ("00", "01", "02", "03")]
    cloud = np.concatenate([pts[:, :3].astype(np.float64), np.ones((len(pts), 1))], axis=1)
    for k, (uv, col) in enumerate(rb.project_pointcloud_on_frame(cloud, T, R_rect, P)):
        out[f"draw_count_{k}"] = np.array(len(uv))
        out[f"draw_uv_sum_{k}"] = uv.astype(np.int64).sum(axis=0)
        if k in (0, 2):  # two cameras in full, the other two as counts and checksums
            out[f"draw_uv_{k}"] = uv.astype(np.int32)
            out[f"draw_green_{k}"] = col[:, 1].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "reference_sources_golden.npz"), **out)
    print({k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
