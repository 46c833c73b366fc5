"""the per-point part of the camera projection (reference camera_model.cpp:5-36,38-95), standalone
and fused behind the deskew.  The reference has no test for camera_model.cpp, so this row is oracle-vs-CUDA only."""
import json
import os

import numpy as np
import pytest

import helpers


def calibration():
    with open(os.path.join(helpers.GOLDEN, "kitti_calibration_2011_09_26.json")) as f:
        c = json.load(f)
    T = np.eye(4)
    T[:3, :3] = np.array(c["velo_to_cam"]["R"]).reshape(3, 3)  # row-major in the file (data_io.cpp:191-194)
    T[:3, 3] = c["velo_to_cam"]["T"]
    R_rect = np.array(c["R_rect_00"]).reshape(3, 3)
    P = {k: np.array(v).reshape(3, 4) for k, v in c["P_rect"].items()}
    return T, R_rect, P


def test_oracle_projection_is_the_devkit_formula(oracle):
    """Y = P_rect_xx * R_rect_00 * (R|T)_velo_to_cam * X (camera_model.cpp:47), filters :21-23, colour :27-28."""
    T, R_rect, P = calibration()
    pts = helpers.real_scan()[::50]
    cloud = np.concatenate([pts[:, :3].astype(np.float64), np.ones((len(pts), 1))], axis=1)
    uv, valid, color, rect = oracle.project_pointcloud(cloud, T, R_rect, P["02"], 15.0)
    R4 = np.eye(4)
    R4[:3, :3] = R_rect
    want_rect = (R4 @ T @ cloud.T).T
    pix = (P["02"] @ want_rect.T).T
    assert np.abs(rect - want_rect[:, :3]).max() < 1e-12
    assert np.allclose(uv, pix[:, :2] / pix[:, 2:3], rtol=1e-12, atol=1e-9)  # points at ~0 depth have huge |u|
    want_valid = ~((want_rect[:, 2] < 0.01) | (want_rect[:, 2] > 15.0) | (want_rect[:, 1] > 1.25))
    assert np.array_equal(valid, want_valid) and 0 < valid.sum() < len(valid)
    assert np.abs(color - 255.0 * want_rect[:, 2] / 14.99).max() < 1e-9


def test_camera_params_compose_the_calibration(capi):
    T, R_rect, P = calibration()
    for name, Pk in P.items():
        cam = capi.camera_params_from_calibration(Pk, R_rect, T, 15.0)
        rect = (R_rect @ T[:3, :])
        R4 = np.eye(4)
        R4[:3, :] = rect
        pix = Pk @ R4
        assert np.abs(np.array(cam.rect).reshape(3, 4) - rect).max() < 1e-6
        assert np.abs(np.array(cam.pix).reshape(3, 4) - pix).max() < 1e-3 * 1e-1  # entries up to ~720, float32
        assert cam.min_depth == np.float32(0.01) and cam.max_range == 15.0 and cam.max_below == 1.25
        assert abs(cam.color_gain - 255.0 / 14.99) < 1e-5
    with pytest.raises(capi.KmcError):
        capi.camera_params_from_calibration(P["00"], R_rect, T, 0.0)


def _compare(out, uv, valid, color, rect):
    """Pixel parity for points the reference keeps, culling parity away from the thresholds."""
    near_threshold = (np.abs(rect[:, 2] - 0.01) < 1e-4) | (np.abs(rect[:, 2] - 15.0) < 1e-4) | (np.abs(rect[:, 1] - 1.25) < 1e-4)
    got_valid = out[:, 3] >= 0
    assert np.array_equal(got_valid[~near_threshold], valid[~near_threshold])
    # pixels that can land on (or near) the 1242 x 375 image; far off-image points have |u| of 1e4 px and more, where
    # 1e-7 relative is no longer 1e-2 px (OpenCV discards them anyway, camera_model.cpp:18-20)
    on_image = (uv[:, 0] > -200) & (uv[:, 0] < 1442) & (uv[:, 1] > -200) & (uv[:, 1] < 575)
    keep = valid & got_valid & (rect[:, 2] > 0.5) & on_image
    assert keep.sum() > 100
    assert np.abs(out[keep, 0] - uv[keep, 0]).max() < 2e-2, "u differs by more than 0.02 px"
    assert np.abs(out[keep, 1] - uv[keep, 1]).max() < 2e-2, "v differs by more than 0.02 px"
    assert np.abs(out[keep, 2] - rect[keep, 2]).max() < 1e-5
    assert np.abs(out[keep, 3] - color[keep]).max() < 1e-3
    # the reference truncates to integer pixels (cv::Point); away from pixel edges the integers agree
    frac = np.minimum(uv[keep] % 1.0, 1.0 - uv[keep] % 1.0).min(axis=1)
    safe = frac > 0.05
    assert np.array_equal(out[keep][safe, :2].astype(np.int64), uv[keep][safe].astype(np.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("camera", ["00", "01", "02", "03"])
def test_projection_matches_oracle(capi, oracle, cuda, camera):
    torch = cuda
    T, R_rect, P = calibration()
    pts = helpers.real_scan()
    cam = capi.camera_params_from_calibration(P[camera], R_rect, T, 15.0)
    d_in = torch.from_numpy(pts).cuda()
    d_pix = torch.empty_like(d_in)
    capi.project_frame_device(d_in.data_ptr(), d_pix.data_ptr(), len(pts), cam, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    cloud = np.concatenate([pts[:, :3].astype(np.float64), np.ones((len(pts), 1))], axis=1)
    uv, valid, color, rect = oracle.project_pointcloud(cloud, T, R_rect, P[camera], 15.0)
    _compare(d_pix.cpu().numpy(), uv, valid, color, rect)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3, 255, 257, 100_001])
def test_projection_ragged_and_unaligned(capi, cuda, n):
    torch = cuda
    T, R_rect, P = calibration()
    cam = capi.camera_params_from_calibration(P["02"], R_rect, T, 15.0)
    pts = helpers.real_scan()[:n]
    buf = torch.zeros((n + 2, 4), dtype=torch.float32, device="cuda")
    buf[1:n + 1] = torch.from_numpy(pts).cuda()
    out_a = torch.full((n + 2, 4), 7.0, dtype=torch.float32, device="cuda")
    out_b = torch.full((n + 2, 4), 7.0, dtype=torch.float32, device="cuda")
    capi.project_frame_device(buf[1:].data_ptr(), out_a[1:].data_ptr(), n, cam)       # 16-byte aligned only: 128-bit path
    aligned = buf[1:n + 1].clone()
    capi.project_frame_device(aligned.data_ptr(), out_b.data_ptr(), n, cam)             # 256-bit path
    torch.cuda.synchronize()
    assert torch.equal(out_a[1:n + 1], out_b[:n])
    assert float(out_a[0].min()) == 7.0 and float(out_a[n + 1].min()) == 7.0 and float(out_b[n:].min()) == 7.0


@pytest.mark.gpu
def test_fused_deskew_project_equals_two_passes(capi, oracle, cuda):
    """handlers.cpp:83-87 (deskew, then project the deskewed cloud) as ONE kernel: bit-identical to the two-pass result."""
    torch = cuda
    T, R_rect, P = calibration()
    cam = capi.camera_params_from_calibration(P["02"], R_rect, T, 15.0)
    pts = helpers.real_scan()
    T_start, T_end, t0, t1, t2 = helpers.config1_frame()
    p = capi.frame_params_from_poses(T_start, T_end, t0, t2, t1)
    n = len(pts)
    s = torch.cuda.current_stream().cuda_stream
    d_in = torch.from_numpy(pts).cuda()
    desk, pix2 = torch.empty_like(d_in), torch.empty_like(d_in)
    capi.deskew_frame_device(d_in.data_ptr(), desk.data_ptr(), n, p, 0, s)
    capi.project_frame_device(desk.data_ptr(), pix2.data_ptr(), n, cam, s)
    fused_cloud, fused_pix, pix_only = torch.empty_like(d_in), torch.empty_like(d_in), torch.empty_like(d_in)
    capi.deskew_project_frame_device(d_in.data_ptr(), fused_cloud.data_ptr(), fused_pix.data_ptr(), n, p, cam, 0, s)
    capi.deskew_project_frame_device(d_in.data_ptr(), 0, pix_only.data_ptr(), n, p, cam, 0, s)
    torch.cuda.synchronize()
    assert torch.equal(fused_cloud, desk) and torch.equal(fused_pix, pix2) and torch.equal(pix_only, pix2)
    # and against the oracle: reference deskew (double) followed by reference projection
    ref = oracle.deskew_xyzi_scan(pts, T_start, T_end, t0, t2, t1)
    uv, valid, color, rect = oracle.project_pointcloud(ref, T, R_rect, P["02"], 15.0)
    _compare(fused_pix.cpu().numpy(), uv, valid, color, rect)
    with pytest.raises(capi.KmcError):
        capi.deskew_project_frame_device(d_in.data_ptr(), fused_cloud.data_ptr(), fused_cloud.data_ptr(), n, p, cam, 0, s)


@pytest.mark.gpu
def test_four_cameras_in_one_pass_equal_four_launches(capi, cuda):
    """camera_model.cpp:85-92 projects the same cloud onto image_00..03; here one kernel reads the cloud once and writes
    the four draw lists — bit-identical to four single-camera launches, with and without the fused deskew."""
    torch = cuda
    T, R_rect, P = calibration()
    cams = [capi.camera_params_from_calibration(P[k], R_rect, T, 15.0) for k in ("00", "01", "02", "03")]
    pts = helpers.real_scan()
    n = len(pts)
    s = torch.cuda.current_stream().cuda_stream
    d_in = torch.from_numpy(pts).cuda()
    singles = [torch.empty_like(d_in) for _ in range(4)]
    for c in range(4):
        capi.project_frame_device(d_in.data_ptr(), singles[c].data_ptr(), n, cams[c], s)
    planes = [torch.empty_like(d_in) for _ in range(4)]
    capi.deskew_project_frame4_device(d_in.data_ptr(), 0, [p.data_ptr() for p in planes], n, None, cams, 0, s)
    torch.cuda.synchronize()
    for c in range(4):
        assert torch.equal(planes[c], singles[c]), f"camera {c}"
    # fused behind the deskew, cloud written as well
    T_start, T_end, t0, t1, t2 = helpers.config1_frame()
    p = capi.frame_params_from_poses(T_start, T_end, t0, t2, t1)
    desk = torch.empty_like(d_in)
    capi.deskew_frame_device(d_in.data_ptr(), desk.data_ptr(), n, p, 0, s)
    for c in range(4):
        capi.project_frame_device(desk.data_ptr(), singles[c].data_ptr(), n, cams[c], s)
    cloud = torch.empty_like(d_in)
    capi.deskew_project_frame4_device(d_in.data_ptr(), cloud.data_ptr(), [q.data_ptr() for q in planes], n, p, cams, 0, s)
    torch.cuda.synchronize()
    assert torch.equal(cloud, desk)
    for c in range(4):
        assert torch.equal(planes[c], singles[c]), f"camera {c} (fused)"
    with pytest.raises(capi.KmcError):  # two planes aliasing each other
        capi.deskew_project_frame4_device(d_in.data_ptr(), 0, [planes[0].data_ptr()] * 4, n, None, cams, 0, s)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3, 255, 257, 100_001])
def test_four_camera_kernel_ragged_and_unaligned(capi, cuda, n):
    """The four-camera kernel's 256-bit path (with its odd tail) and its 128-bit path (buffers only 16-byte aligned) both
    equal four single-camera launches bit for bit, with and without the fused deskew, and stay inside their buffers."""
    torch = cuda
    T, R_rect, P = calibration()
    cams = [capi.camera_params_from_calibration(P[k], R_rect, T, 15.0) for k in ("00", "01", "02", "03")]
    pts = helpers.real_scan()[:n]
    T_start, T_end, t0, t1, t2 = helpers.config1_frame()
    p = capi.frame_params_from_poses(T_start, T_end, t0, t2, t1)
    s = torch.cuda.current_stream().cuda_stream
    aligned = torch.from_numpy(pts).cuda()
    shifted = torch.zeros((n + 2, 4), dtype=torch.float32, device="cuda")
    shifted[1:n + 1] = aligned
    for params in (None, p):
        want = [torch.empty_like(aligned) for _ in range(4)]
        src = aligned
        if params is not None:
            src = torch.empty_like(aligned)
            capi.deskew_frame_device(aligned.data_ptr(), src.data_ptr(), n, params, 0, s)
        for c in range(4):
            capi.project_frame_device(src.data_ptr(), want[c].data_ptr(), n, cams[c], s)
        for offset in (0, 1):  # 0: 32-byte aligned buffers (256-bit path), 1: shifted by one point (128-bit path)
            planes = [torch.full((n + 2, 4), 7.0, dtype=torch.float32, device="cuda") for _ in range(4)]
            cloud = torch.full((n + 2, 4), 7.0, dtype=torch.float32, device="cuda")
            d_in = shifted[1:] if offset else aligned
            capi.deskew_project_frame4_device(d_in.data_ptr(), cloud[offset:].data_ptr() if params is not None else 0,
                                              [q[offset:].data_ptr() for q in planes], n, params, cams, 0, s)
            torch.cuda.synchronize()
            for c in range(4):
                assert torch.equal(planes[c][offset:offset + n], want[c]), (n, offset, c)
                assert float(planes[c][offset + n:].min()) == 7.0 and (offset == 0 or float(planes[c][0].min()) == 7.0)
            if params is not None:
                assert torch.equal(cloud[offset:offset + n], src)
                assert float(cloud[offset + n:].min()) == 7.0
