"""Batch forms of the secondary entry points (reference: the viz handler's frame loop, handlers.cpp:67-92 —
per frame one MotionCompensateFrame and four projections): kmc_b200_deskew_project_batch_device,
kmc_b200_deskew_cloud_f64_batch_device, kmc_b200_deskew_cloud_f64_batch_host.  Each is bit-identical to the per-frame call
it batches, and every frame is checked against the oracle / the reference's compiled sources."""
import ctypes as C

import numpy as np
import pytest

import helpers
from oracle import ref_binding as rb
from test_deskew_gpu import batch_params, dev, make_batch
from test_projection import calibration

pytestmark = pytest.mark.gpu

SIZES = [30_001, 0, 1, 4_096, 77_777, 3, 0, 130_000, 2_049]


def cameras(capi):
    T, R_rect, P = calibration()
    return [capi.camera_params_from_calibration(P[k], R_rect, T, 15.0) for k in ("00", "01", "02", "03")]


@pytest.mark.parametrize("n_cam", [1, 4])
@pytest.mark.parametrize("with_cloud", [True, False])
@pytest.mark.parametrize("shift", [0, 1])
def test_deskew_project_batch_equals_per_frame_calls(capi, oracle, cuda, n_cam, with_cloud, shift):
    """One launch over the batch == per-frame kmc_b200_deskew_project_frame(4)_device calls, bit for bit (ragged frames, empty
    frames, odd offsets; shift = 1 moves every buffer off 32-byte alignment so the 128-bit path runs too)."""
    torch = cuda
    pts, offsets, frames = make_batch(oracle, SIZES, 5100)
    params = batch_params(capi, frames)
    cams = cameras(capi)[:n_cam] if n_cam == 4 else [cameras(capi)[2]]
    n = len(pts)
    st = torch.cuda.current_stream().cuda_stream

    def buf():
        return torch.zeros((n + 1, 4), dtype=torch.float32, device="cuda")[shift:shift + n]

    d_in = buf()
    d_in.copy_(torch.from_numpy(pts))
    d_off, d_par = dev(torch, offsets), dev(torch, params.view(np.uint8))
    cloud_b = buf() if with_cloud else None
    pix_b = [buf() for _ in range(n_cam)]
    capi.deskew_project_batch_device(d_in.data_ptr(), cloud_b.data_ptr() if with_cloud else 0, [p.data_ptr() for p in pix_b], d_off.data_ptr(),
                                     d_par.data_ptr(), len(SIZES), n, cams, 0, st)
    cloud_f = buf() if with_cloud else None
    pix_f = [buf() for _ in range(n_cam)]
    for f in range(len(SIZES)):
        a, b = int(offsets[f]), int(offsets[f + 1])
        if a == b:
            continue
        p = capi.FrameParams.from_buffer_copy(params[f:f + 1].tobytes())
        cl = cloud_f[a:].data_ptr() if with_cloud else 0
        if n_cam == 4:
            capi.deskew_project_frame4_device(d_in[a:].data_ptr(), cl, [q[a:].data_ptr() for q in pix_f], b - a, p, cams, 0, st)
        else:
            capi.deskew_project_frame_device(d_in[a:].data_ptr(), cl, pix_f[0][a:].data_ptr(), b - a, p, cams[0], 0, st)
    torch.cuda.synchronize()
    for c in range(n_cam):
        assert torch.equal(pix_b[c].view(torch.int32), pix_f[c].view(torch.int32)), f"camera {c}"
    if with_cloud:
        assert torch.equal(cloud_b.view(torch.int32), cloud_f.view(torch.int32))
        # and the cloud is the plain batched deskew's
        want = torch.zeros_like(cloud_b)
        capi.deskew_batch_device(d_in.data_ptr(), want.data_ptr(), d_off.data_ptr(), d_par.data_ptr(), len(SIZES), n, 0, st)
        torch.cuda.synchronize()
        assert torch.equal(cloud_b, want)


@pytest.mark.skipif(not rb.available(), reason="oracle/_ref/libkmc_ref.so not built")
def test_deskew_project_batch_against_the_reference_draw_lists(capi, oracle, cuda):
    """Every frame of the batch: deskew with the reference's MotionCompensateFrame, project with the reference's own
    camera_model.cpp (recording cv::circle) — the same points drawn at the same integer pixels with the same colours."""
    torch = cuda
    T, R_rect, P = calibration()
    sizes = [90_000, 70_000, 0, 130_000]
    pts, offsets, frames = make_batch(oracle, sizes, 5200)
    params = batch_params(capi, frames)
    cams = cameras(capi)
    n = len(pts)
    d_in, d_off, d_par = dev(torch, pts), dev(torch, offsets), dev(torch, params.view(np.uint8))
    pix = [torch.zeros_like(d_in) for _ in range(4)]
    capi.deskew_project_batch_device(d_in.data_ptr(), 0, [p.data_ptr() for p in pix], d_off.data_ptr(), d_par.data_ptr(), len(sizes), n, cams, 0,
                                     torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = [p.cpu().numpy() for p in pix]
    from test_deskew_gpu_vs_reference_sources import _draw_list_parity
    for f, (Ts, Te, _, xr) in enumerate(frames):
        a, b = int(offsets[f]), int(offsets[f + 1])
        if a == b:
            continue
        ref_cloud = rb.deskew_xyzi_scan(pts[a:b], Ts, Te, 10.0, 10.1, 10.0 + xr * 0.1)
        _draw_list_parity([g[a:b] for g in got], rb.project_pointcloud_on_frame(ref_cloud, T, R_rect, [P[k] for k in ("00", "01", "02", "03")]))


def f64_batch_inputs(oracle, sizes, seed):
    """Frames in the reference's layout, stored back to back: frame f's column-major N_f x 4 block at 4 * offsets[f]."""
    rng = np.random.default_rng(seed)
    pts, offsets, frames = make_batch(oracle, sizes, seed, mercator=True)
    F = len(sizes)
    times = np.zeros((F, 3))
    cloud = np.zeros(4 * len(pts))
    stamps = np.zeros(len(pts))
    per_frame = []
    for f, (Ts, Te, _, xr) in enumerate(frames):
        a, b = int(offsets[f]), int(offsets[f + 1])
        t0 = 100.0 + 0.1 * f
        times[f] = [t0, t0 + 0.1, t0 + 0.1 * xr]
        c = np.concatenate([pts[a:b, :3].astype(np.float64) + rng.uniform(-1e-7, 1e-7, (b - a, 3)), np.ones((b - a, 1))], axis=1)
        ts = oracle.pseudo_time_stamps(c, t0, t0 + 0.1) if b > a else np.zeros(0)
        if b - a > 10:
            ts[:5] = rng.uniform(t0, t0 + 0.1, 5)  # arbitrary stamps
        cloud[4 * a:4 * b] = c.T.reshape(-1)
        stamps[a:b] = ts
        per_frame.append((c, ts))
    return pts, offsets, frames, times, cloud, stamps, per_frame


def test_f64_batch_device_equals_per_frame_and_the_reference(capi, oracle, cuda):
    torch = cuda
    sizes = [20_001, 0, 1, 4_097, 55_555, 130_000]
    pts, offsets, frames, times, cloud, stamps, per_frame = f64_batch_inputs(oracle, sizes, 5300)
    F, n = len(sizes), len(pts)
    params = capi.params_array([capi.frame_params_from_poses(Ts, Te, times[f, 0], times[f, 1], times[f, 2]) for f, (Ts, Te, _, _) in enumerate(frames)])
    # frame 3 gets a stamp outside its interval and frame 4 a non-homogeneous w: per-frame flags 1 and 2
    stamps[offsets[3] + 100] = times[3, 1] + 1e-3
    cloud[4 * offsets[4] + 3 * sizes[4] + 7] = 2.5
    per_frame[3][1][100] = times[3, 1] + 1e-3
    per_frame[4][0][7, 3] = 2.5
    st = torch.cuda.current_stream().cuda_stream
    d_cloud, d_stamps = dev(torch, cloud), dev(torch, stamps)
    d_out = torch.zeros_like(d_cloud)
    d_off, d_par, d_times = dev(torch, offsets), dev(torch, params.view(np.uint8)), dev(torch, times.reshape(-1))
    d_flags = torch.full((F,), 99, dtype=torch.int32, device="cuda")
    capi.deskew_cloud_f64_batch_device(d_cloud.data_ptr(), d_stamps.data_ptr(), d_out.data_ptr(), d_off.data_ptr(), d_par.data_ptr(),
                                       d_times.data_ptr(), F, n, d_flags.data_ptr(), st)
    torch.cuda.synchronize()
    assert d_flags.cpu().tolist() == [0, 0, 0, 1, 2, 0]
    out = d_out.cpu().numpy()
    one_flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    for f, (Ts, Te, _, _) in enumerate(frames):
        a, b = int(offsets[f]), int(offsets[f + 1])
        if a == b:
            continue
        got = out[4 * a:4 * b].reshape(4, b - a).T
        # per-frame device call: the very same bits
        p = capi.FrameParams.from_buffer_copy(params[f:f + 1].tobytes())
        single = torch.zeros(4 * (b - a), dtype=torch.float64, device="cuda")
        capi.check(capi.lib().kmc_b200_deskew_cloud_f64_device(d_cloud[4 * a:].data_ptr(), d_stamps[a:].data_ptr(), single.data_ptr(), b - a, times[f, 0],
                                                                times[f, 1], times[f, 2], C.byref(p), one_flag.data_ptr(), st))
        torch.cuda.synchronize()
        assert np.array_equal(single.cpu().numpy(), out[4 * a:4 * b]), f"frame {f}"
        if f == 3:
            continue  # the reference aborts on this frame
        c, ts = per_frame[f]
        engine = rb if rb.available() else oracle
        ref = engine.motion_compensate_frame(c, ts, Ts, Te, times[f, 0], times[f, 1], times[f, 2])
        disp = float(np.abs(ref[:, :3] - c[:, :3]).max())
        assert np.abs(got[:, :3] - ref[:, :3]).max() < 2e-7 + 4e-7 * disp, f"frame {f}"  # fp32 displacement added to the double coordinate
        assert np.array_equal(got[:, 3], c[:, 3])


@pytest.mark.parametrize("shift", [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1)])
def test_f64_kernels_two_points_per_thread_equal_one_point_per_thread(capi, oracle, cuda, monkeypatch, shift):
    """The f64 kernels take two points per thread with 128-bit accesses on every 16-byte-aligned column and 64-bit accesses on
    the others.  Every mix of alignments — odd frame sizes, odd frame starts, cloud / stamps / result pointers that are only
    8-byte aligned — gives the bits of the one-point-per-thread kernels (KMC_B200_TUNE f64_pair=0), for the batch and the
    single-frame entry point, and nothing is written outside the result block."""
    torch = cuda
    sizes = [2, 3, 1, 0, 511, 512, 513, 1_025, 40_000, 12_345, 130_001, 6]
    pts, offsets, frames, times, cloud, stamps, per_frame = f64_batch_inputs(oracle, sizes, 811)
    F, n = len(sizes), len(pts)
    params = capi.params_array([capi.frame_params_from_poses(Ts, Te, times[f, 0], times[f, 1], times[f, 2]) for f, (Ts, Te, _, _) in enumerate(frames)])
    st = torch.cuda.current_stream().cuda_stream
    sc, ss, so = shift
    d_cloud_buf, d_stamps_buf = torch.zeros(4 * n + 2, dtype=torch.float64, device="cuda"), torch.zeros(n + 2, dtype=torch.float64, device="cuda")
    d_cloud, d_stamps = d_cloud_buf[sc:sc + 4 * n], d_stamps_buf[ss:ss + n]
    d_cloud.copy_(torch.from_numpy(cloud))
    d_stamps.copy_(torch.from_numpy(stamps))
    d_off, d_par, d_times = dev(torch, offsets), dev(torch, params.view(np.uint8)), dev(torch, times.reshape(-1))
    d_flags = torch.zeros(F, dtype=torch.int32, device="cuda")
    results = {}
    for pair in (0, 1):
        monkeypatch.setenv("KMC_B200_TUNE", f"f64_pair={pair}")
        buf = torch.full((4 * n + 2,), -7.0, dtype=torch.float64, device="cuda")
        d_out = buf[so:so + 4 * n]
        capi.deskew_cloud_f64_batch_device(d_cloud.data_ptr(), d_stamps.data_ptr(), d_out.data_ptr(), d_off.data_ptr(), d_par.data_ptr(),
                                           d_times.data_ptr(), F, n, d_flags.data_ptr(), st)
        torch.cuda.synchronize()
        assert d_flags.cpu().tolist() == [0] * F
        whole = buf.cpu().numpy()
        assert np.all(whole[:so] == -7.0) and np.all(whole[so + 4 * n:] == -7.0)
        single = torch.full((4 * n + 2,), -7.0, dtype=torch.float64, device="cuda")
        one_flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        for f in range(F):
            a, b = int(offsets[f]), int(offsets[f + 1])
            if a == b:
                continue
            p = capi.FrameParams.from_buffer_copy(params[f:f + 1].tobytes())
            capi.check(capi.lib().kmc_b200_deskew_cloud_f64_device(d_cloud[4 * a:].data_ptr(), d_stamps[a:].data_ptr(), single[so + 4 * a:].data_ptr(),
                                                                    b - a, times[f, 0], times[f, 1], times[f, 2], C.byref(p), one_flag.data_ptr(), st))
        torch.cuda.synchronize()
        assert int(one_flag.item()) == 0
        results[pair] = (whole[so:so + 4 * n].copy(), single.cpu().numpy()[so:so + 4 * n].copy())
        assert np.array_equal(results[pair][0], results[pair][1])
    assert np.array_equal(results[0][0].view(np.int64), results[1][0].view(np.int64))
    # in place (out == cloud): every thread reads its own points before it writes them
    in_place = d_cloud.clone() if sc == 0 else d_cloud_buf.clone()[sc:sc + 4 * n]
    capi.deskew_cloud_f64_batch_device(in_place.data_ptr(), d_stamps.data_ptr(), in_place.data_ptr(), d_off.data_ptr(), d_par.data_ptr(),
                                       d_times.data_ptr(), F, n, d_flags.data_ptr(), st)
    torch.cuda.synchronize()
    assert np.array_equal(in_place.cpu().numpy(), results[1][0])
    # and the values are the reference's (largest frame)
    f = sizes.index(130_001)
    a, b = int(offsets[f]), int(offsets[f + 1])
    Ts, Te, _, _ = frames[f]
    c, ts = per_frame[f]
    engine = rb if rb.available() else oracle
    ref = engine.motion_compensate_frame(c, ts, Ts, Te, times[f, 0], times[f, 1], times[f, 2])
    got = results[1][0][4 * a:4 * b].reshape(4, b - a).T
    disp = float(np.abs(ref[:, :3] - c[:, :3]).max())
    assert np.abs(got[:, :3] - ref[:, :3]).max() < 2e-7 + 4e-7 * disp


def test_f64_batch_host_equals_single_frame_host_calls(capi, oracle, cuda):
    sizes = [123_397, 0, 5, 70_001, 9_000]
    _, _, frames, times, _, _, per_frame = f64_batch_inputs(oracle, sizes, 5400)
    params = capi.params_array([capi.frame_params_from_poses(Ts, Te, times[f, 0], times[f, 1], times[f, 2]) for f, (Ts, Te, _, _) in enumerate(frames)])
    per_frame[3][1][17] = times[3, 0] - 1.0   # out-of-range stamp in frame 3
    with capi.Handle(0, 250_000) as h:
        outs, flags, rc = h.deskew_cloud_f64_batch([c for c, _ in per_frame], [t for _, t in per_frame], times, params)
        assert rc == capi.ERR_TIME_OUT_OF_RANGE and flags.tolist() == [0, 0, 0, 1, 0]
        assert "frame 3" in capi.last_error()
        for f, (c, ts) in enumerate(per_frame):
            p = capi.FrameParams.from_buffer_copy(params[f:f + 1].tobytes())
            one, fl, rc1 = h.deskew_cloud_f64(c, ts, times[f, 0], times[f, 1], times[f, 2], p)
            assert fl == flags[f]
            assert one.tobytes() == outs[f].tobytes(), f"frame {f}"
            if f != 3 and len(c):
                Ts, Te = frames[f][0], frames[f][1]
                ref = oracle.motion_compensate_frame(c[::7], ts[::7], Ts, Te, times[f, 0], times[f, 1], times[f, 2])
                assert np.abs(one[::7, :3] - ref[:, :3]).max() < 3e-6
        # argument checks happen before anything runs
        bad_times = times.copy()
        bad_times[2, 2] = bad_times[2, 1] + 1.0
        _, _, rc = h.deskew_cloud_f64_batch([c for c, _ in per_frame], [t for _, t in per_frame], bad_times, params)
        assert rc == capi.ERR_TIME_OUT_OF_RANGE and "frame 2" in capi.last_error()
