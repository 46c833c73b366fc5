"""tests/golden/reference_sources_golden.npz — outputs of the reference's OWN compiled sources (oracle/_ref), generated
by tests/golden/make_reference_golden.py in the development container and committed, so that the oracle (CPU) and the
CUDA path (GPU) are checked against the reference's code even on a box where oracle/_ref itself is absent."""
import json
import os

import numpy as np
import pytest

import helpers
from helpers import TOL_M


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(helpers.GOLDEN, "reference_sources_golden.npz"))


def _synthetic_inputs(golden, oracle):
    syn = helpers.synthetic_scan(20_000, 64, 926)
    P1 = golden["synthetic_pose_start"]
    return syn, P1, P1 @ oracle.se3_exp(golden["synthetic_twist"])


def _calibration():
    with open(os.path.join(helpers.GOLDEN, "kitti_calibration_2011_09_26.json")) as f:
        c = json.load(f)
    T = np.eye(4)
    T[:3, :3] = np.array(c["velo_to_cam"]["R"]).reshape(3, 3)
    T[:3, 3] = c["velo_to_cam"]["T"]
    return T, np.array(c["R_rect_00"]).reshape(3, 3), [np.array(c["P_rect"][k]).reshape(3, 4) for k in ("00", "01", "02", "03")]


def test_fixture_is_current_when_the_reference_library_is_built(golden):
    """Regenerating the fixture must not change it (skipped where oracle/_ref is absent)."""
    from oracle import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref not built")
    pts = helpers.real_scan()
    T_start, T_end, t0, t1, t2 = helpers.config1_frame()
    stride = int(golden["stride"])
    assert np.array_equal(rb.deskew_xyzi_scan(pts, T_start, T_end, t0, t2, t1)[::stride, :3], golden["config1_middle"])
    assert str(golden["eigen_provider"]) == rb.eigen_provider()


def test_oracle_against_the_fixture(golden, oracle):
    pts = helpers.real_scan()
    T_start, T_end, t0, t1, t2 = helpers.config1_frame()
    stride = int(golden["stride"])
    for name, t_req in (("middle", t1), ("start", t0), ("end", t2)):
        got = oracle.deskew_xyzi_scan(pts[::stride], T_start, T_end, t0, t2, t_req)[:, :3]
        assert np.abs(got - golden[f"config1_{name}"]).max() < 1e-8  # both lose ~1e-9 m to the Mercator-magnitude poses
    syn, P1, P2 = _synthetic_inputs(golden, oracle)
    got = oracle.deskew_xyzi_scan(syn[::5], P1, P2, 0.0, 0.1, 0.03)[:, :3]
    assert np.abs(got - golden["synthetic_out"]).max() < 1e-11
    T, R_rect, P = _calibration()
    cloud = np.concatenate([pts[:, :3].astype(np.float64), np.ones((len(pts), 1))], axis=1)
    for k in range(4):
        uv, valid, color, _ = oracle.project_pointcloud(cloud, T, R_rect, P[k], 15.0)
        assert int(valid.sum()) == int(golden[f"draw_count_{k}"])
        assert np.array_equal(uv[valid].astype(np.int64).sum(axis=0), golden[f"draw_uv_sum_{k}"])
        if k in (0, 2):
            assert np.array_equal(uv[valid].astype(np.int32), golden[f"draw_uv_{k}"])
            assert np.abs(color[valid] - golden[f"draw_green_{k}"]).max() < 1e-4


@pytest.mark.gpu
def test_cuda_against_the_fixture(golden, capi, oracle, cuda):
    torch = cuda
    from test_deskew_gpu import run_frame
    pts = helpers.real_scan()
    T_start, T_end, t0, t1, t2 = helpers.config1_frame()
    stride = int(golden["stride"])
    for name, t_req in (("middle", t1), ("start", t0), ("end", t2)):
        out = run_frame(torch, capi, pts, capi.frame_params_from_poses(T_start, T_end, t0, t2, t_req))
        err = float(np.abs(out[::stride, :3].astype(np.float64) - golden[f"config1_{name}"]).max())
        assert err < TOL_M, (name, err)
    syn, P1, P2 = _synthetic_inputs(golden, oracle)
    out = run_frame(torch, capi, syn, capi.frame_params_from_poses(P1, P2, 0.0, 0.1, 0.03))
    assert float(np.abs(out[::5, :3].astype(np.float64) - golden["synthetic_out"]).max()) < TOL_M
    # projection: the same points drawn; integer pixels equal wherever the kernel's pixel is not within 0.02 px of an edge
    T, R_rect, P = _calibration()
    d_in = torch.from_numpy(pts).cuda()
    for k in (0, 2):
        cam = capi.camera_params_from_calibration(P[k], R_rect, T, 15.0)
        d_pix = torch.empty_like(d_in)
        capi.project_frame_device(d_in.data_ptr(), d_pix.data_ptr(), len(pts), cam, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        pix = d_pix.cpu().numpy()
        keep = pix[:, 3] >= 0
        assert abs(int(keep.sum()) - int(golden[f"draw_count_{k}"])) <= 2
        if int(keep.sum()) == int(golden[f"draw_count_{k}"]):
            got, ref_uv = pix[keep], golden[f"draw_uv_{k}"]
            on_image = (ref_uv[:, 0] >= 0) & (ref_uv[:, 0] < 1242) & (ref_uv[:, 1] >= 0) & (ref_uv[:, 1] < 375) & (got[:, 2] > 0.5)
            edge = np.minimum(np.abs(got[:, :2]) % 1.0, 1.0 - np.abs(got[:, :2]) % 1.0).min(axis=1) < 0.02
            same = (np.trunc(got[:, 0]).astype(np.int64) == ref_uv[:, 0]) & (np.trunc(got[:, 1]).astype(np.int64) == ref_uv[:, 1])
            assert (on_image & ~edge).sum() > 1000 and np.all(same[on_image & ~edge])
            assert np.abs(got[:, 3] - golden[f"draw_green_{k}"]).max() < 1e-3
