"""The CPU oracle against every known-answer vector the reference's own tests hold for the deskew path
(SURVEY 8c).  These pin the oracle; the CUDA path is then compared with the oracle in tests/test_deskew_gpu.py."""
import json
import os

import numpy as np
import pytest

import helpers


def float_eq(a, b):
    """gtest ASSERT_FLOAT_EQ: both rounded to float32, within 4 ULPs."""
    a32, b32 = np.float32(a), np.float32(b)
    if a32 == b32:
        return True
    ia, ib = int(a32.view(np.int32)), int(b32.view(np.int32))
    ia = ia if ia >= 0 else -(ia & 0x7FFFFFFF)
    ib = ib if ib >= 0 else -(ib & 0x7FFFFFFF)
    return abs(ia - ib) <= 4


def matrices_same(oracle, a, b):
    """utilities_for_testing.hpp:6-11 (float-cast trace / off-trace sum of A B^-1, eps 1e-10)."""
    m = a @ oracle.affine_inverse(b)
    return abs(np.float32(np.trace(m)) - np.float32(4.0)) <= 1e-10 and abs(np.float32(m.sum() - np.trace(m))) <= 1e-10


def golden_frame(oracle, kats):
    g = kats["motion_compensate_frame"]
    o = [helpers.oxts7(x) for x in g["oxts"]]
    T_start, T_end = oracle.make_frame_poses(o[0], o[1], o[2], g["stamp_start"], g["stamp_end"])
    return g, np.array(g["cloud"]), T_start, T_end


def test_motion_compensate_frame_golden(oracle, kats):
    # test/test_motion_compensation.cpp:54-76
    g, cloud, T_start, T_end = golden_frame(oracle, kats)
    stamps = oracle.pseudo_time_stamps(cloud, g["stamp_start"], g["stamp_end"])
    out = oracle.motion_compensate_frame(cloud, stamps, T_start, T_end, g["stamp_start"], g["stamp_end"], g["requested_time"])
    for got, want in zip(out.reshape(-1), np.array(g["expected"]).reshape(-1)):
        assert float_eq(got, want), (got, want)


def test_motion_compensate_point_matches_frame(oracle, kats):
    g, cloud, T_start, T_end = golden_frame(oracle, kats)
    stamps = oracle.pseudo_time_stamps(cloud, g["stamp_start"], g["stamp_end"])
    frame = oracle.motion_compensate_frame(cloud, stamps, T_start, T_end, g["stamp_start"], g["stamp_end"], g["requested_time"])
    for i in range(3):
        p = oracle.motion_compensate_point(g["stamp_start"], T_start, g["stamp_end"], T_end, stamps[i], cloud[i], g["requested_time"])
        assert np.array_equal(p, frame[i])


def test_xyzi_pipeline_equals_reference_layout_call(oracle, kats):
    """deskew_xyzi_scan (loader conversion + stamps + frame) == calling the three reference steps one by one."""
    g, cloud, T_start, T_end = golden_frame(oracle, kats)
    xyzi = cloud.astype(np.float32)
    xyzi[:, 3] = 0.5
    a = oracle.deskew_xyzi_scan(xyzi, T_start, T_end, g["stamp_start"], g["stamp_end"], g["requested_time"])
    for got, want in zip(a.reshape(-1), np.array(g["expected"]).reshape(-1)):
        assert float_eq(got, want)


def test_fraction_of_scan_completed(oracle, kats):
    # test/test_timestamp_mocking.cpp:55-57
    g = kats["fraction_of_scan_completed"]
    for p, want in zip(g["points"], g["expected"]):
        assert float_eq(oracle.fraction_of_scan_completed(p), want)


def test_pseudo_time_stamp(oracle, kats):
    # test/test_timestamp_mocking.cpp:71-73, 84-86
    g = kats["pseudo_time_stamp"]
    for p, want in zip(g["points"], g["expected"]):
        assert float_eq(oracle.pseudo_time_stamp(p, g["scan_start"], g["scan_end"]), want)
    stamps = oracle.pseudo_time_stamps(np.array(g["points"]), g["scan_start"], g["scan_end"])
    for got, want in zip(stamps, g["expected"]):
        assert float_eq(got, want)


def test_lie_hat_vee(oracle, kats):
    phi = kats["lie_algebra"]["phi"]
    for a, b in zip(oracle.vee(oracle.hat(phi)), phi):
        assert float_eq(a, b)


def test_lie_so3_log_exp(oracle, kats):
    phi = kats["lie_algebra"]["phi"]
    for a, b in zip(oracle.so3_log(oracle.so3_exp(phi)), phi):
        assert float_eq(a, b)


def test_lie_jacobians_inverse(oracle, kats):
    phi = kats["lie_algebra"]["phi"]
    m = oracle.left_jacobian(phi) @ oracle.inverse_left_jacobian(phi)
    assert float_eq(np.trace(m), 3.0)
    assert abs(np.float32(m.sum() - np.trace(m))) < 1e-6  # ASSERT_FLOAT_EQ(x, 0.0) passes only for |x| < 4 denormal ulps; the reference value is ~1e-17


def test_lie_se3_log_exp(oracle, kats):
    xi = kats["lie_algebra"]["xi"]
    for a, b in zip(oracle.se3_log(oracle.se3_exp(xi)), xi):
        assert float_eq(a, b)


def artificial_pose(oracle, x_rotation, x_translation):
    # test/test_trajectory_interpolation.cpp:24-30: Identity.rotate(AngleAxis(x_rot, UnitX)); translation = (x_tr, 0, 0)
    T = np.eye(4)
    T[:3, :3] = oracle.so3_exp([x_rotation, 0.0, 0.0])
    T[0, 3] = x_translation
    return T


def test_trajectory_interpolation_midpoint(oracle, kats):
    # test/test_trajectory_interpolation.cpp:43-50
    p = kats["trajectory_interpolation_artificial"]["poses"]
    P = [artificial_pose(oracle, q["x_rotation"], q["x_translation"]) for q in p]
    mid = oracle.pose_at_time(p[0]["time"], P[0], p[2]["time"], P[2], p[1]["time"])
    assert matrices_same(oracle, mid, P[1])


def test_relative_pose_between_times(oracle, kats):
    # test/test_trajectory_interpolation.cpp:52-60
    p = kats["trajectory_interpolation_artificial"]["poses"]
    P = [artificial_pose(oracle, q["x_rotation"], q["x_translation"]) for q in p]
    a = oracle.relative_pose_between_times(p[0]["time"], P[0], p[2]["time"], P[2], p[0]["time"], p[1]["time"])
    b = oracle.relative_pose_between_times(p[0]["time"], P[0], p[2]["time"], P[2], p[1]["time"], p[2]["time"])
    assert matrices_same(oracle, a, b)


def test_out_of_range_time_is_the_abort_case(oracle, kats):
    # test/test_trajectory_interpolation.cpp:77-81 (EXPECT_DEATH): the oracle reports instead of aborting
    r = kats["real_scan_frame0"]
    P = oracle.oxts_to_pose(helpers.oxts7(kats["oxts_to_pose"]["oxts"]))
    with pytest.raises(oracle.ReferenceWouldAbort):
        oracle.pose_at_time(r["oxts_stamp"], P, r["oxts_stamp"] + 0.2, P, 0.0)


def test_oxts_to_pose(oracle, kats):
    # test/test_oxts_to_pose.cpp:17-20
    g = kats["oxts_to_pose"]
    P = oracle.oxts_to_pose(helpers.oxts7(g["oxts"]))
    assert float_eq(np.linalg.det(oracle.polar_rotation(P[:3, :3])), g["expected_det"])
    for got, want in zip(P[:3, 3], g["expected_translation"]):
        assert float_eq(got, want)


def test_real_scan_fixture_matches_reference_expectations(kats):
    # test/test_data_io.cpp:53-78
    g = kats["real_scan_frame0"]
    pts = helpers.real_scan()
    assert pts.shape == (g["num_points"], 4)
    for got, want in zip(pts[0], g["first_point"]):
        assert float_eq(got, want)
    for got, want in zip(pts[-1], g["last_point"]):
        assert float_eq(got, want)
    # SURVEY 8c edge case (i): the scan really contains y == -0.0 with x < 0
    assert np.any((pts[:, 0] < 0) & (pts[:, 1] == 0) & np.signbit(pts[:, 1]))


def test_polar_rotation_is_nearest_rotation(oracle):
    """Eigen's Affine-mode rotation(): U diag(1,1,sign det) V^T — checked against numpy's SVD."""
    rng = np.random.default_rng(7)
    for _ in range(200):
        A = rng.normal(size=(3, 3))
        U, _, Vt = np.linalg.svd(A)
        U[:, 2] *= np.sign(np.linalg.det(U @ Vt))
        assert np.abs(U @ Vt - oracle.polar_rotation(A)).max() < 1e-12


def test_oracle_agrees_with_closed_form_on_real_scan(oracle):
    """literal reference algorithm (per-point Log/inverse products) vs Exp((frac - x_req) xi) p in double: the gap is
    the reference's own Mercator-magnitude cancellation (~1e-9 m), far below the 1e-5 m parity bar."""
    pts = helpers.real_scan()[::7]
    T_start, T_end, t0, t1, t2 = helpers.config1_frame()
    ref = oracle.deskew_xyzi_scan(pts, T_start, T_end, t0, t2, t1)
    cf = helpers.closed_form_deskew(pts, np.array(helpers.CONFIG1_TWIST), (t1 - t0) / (t2 - t0))
    assert np.abs(cf - ref[:, :3]).max() < 5e-8
    assert np.array_equal(ref[:, 3], np.ones(len(pts)))


def test_oracle_digest_unchanged(oracle):
    """Guards the oracle against accidental edits (digest written by tests/golden/make_golden.py)."""
    with open(os.path.join(helpers.GOLDEN, "oracle_real_scan_digest.json")) as f:
        d = json.load(f)
    pts = helpers.real_scan()
    T_start, T_end, t0, t1, t2 = helpers.config1_frame()
    idx = np.array(d["sample_index"])
    ref = oracle.deskew_xyzi_scan(pts[idx], T_start, T_end, t0, t2, t1)
    assert np.abs(ref[:, :3] - np.array(d["sample_xyz"])).max() < 1e-9


def test_oracle_rotating_frames_against_matrix_exponential_in_40_digits(oracle):
    """The reference ships no golden with rotation != 0.  Independent pin for rotating frames: the SE(3) closed forms the
    reference uses (Rodrigues, left Jacobian and its inverse, lie_algebra.cpp:22-103) are the matrix exponential /
    logarithm of 4x4 twist matrices, so  GetPoseAtTime(t) = P1 expm(x logm(P1^-1 P2))  can be evaluated with mpmath's
    generic expm/logm at 40 digits — no Rodrigues, no Jacobians, no code shared with the oracle."""
    import mpmath as mp
    mp.mp.dps = 40
    rng = np.random.default_rng(2011)

    def to_mp(a):
        return mp.matrix([[mp.mpf(float(v)) for v in row] for row in np.asarray(a)])

    for case in range(6):
        P1 = helpers.random_pose(rng, mercator=bool(case % 2))
        xi = helpers.random_twist(rng) * (1.0 if case < 4 else 6.0)  # up to ~0.3 rad per scan
        P2 = P1 @ oracle.se3_exp(xi)
        t1, t2 = 100.0, 100.1
        M1, M2 = to_mp(P1), to_mp(P2)
        L = mp.logm(mp.inverse(M1) * M2)
        # lie::Log of the relative pose == the twist read off logm (rho in the last column, phi in the skew part)
        xi_mp = [L[0, 3], L[1, 3], L[2, 3], L[2, 1], L[0, 2], L[1, 0]]
        xi_or = oracle.se3_log(np.linalg.inv(P1) @ P2)
        assert max(abs(float(a) - b) for a, b in zip(xi_mp, xi_or)) < 1e-9 * (1 if case % 2 == 0 else 100)  # Mercator: 6e6 m cancellation
        for x in (0.0, 0.31, 0.5, 1.0):
            want = M1 * mp.expm(mp.mpf(x) * L)
            got = oracle.pose_at_time(t1, P1, t2, P2, t1 + x * (t2 - t1))
            err_rot = max(abs(float(want[r, c]) - got[r, c]) for r in range(3) for c in range(3))
            err_tr = max(abs(float(want[r, 3]) - got[r, 3]) for r in range(3))
            assert err_rot < 1e-13
            assert err_tr < (2e-8 if case % 2 else 1e-12)  # 1 ulp of a 6e6 m coordinate is 1e-9 m
        # the deskew of points: T(x_req)^-1 T(x_i) p, with x_i from the azimuth
        pts = helpers.synthetic_scan(24, 64, case)
        x_req = 0.5
        ref = oracle.deskew_xyzi_scan(pts, P1, P2, t1, t2, t1 + x_req * (t2 - t1))
        for i in range(len(pts)):
            x, y, z = (mp.mpf(float(v)) for v in pts[i, :3])
            frac = (mp.pi - mp.atan2(y, x)) / (2 * mp.pi)
            T = mp.expm((frac - mp.mpf(x_req)) * L)  # P1 cancels; exponentials of one generator commute
            q = T * mp.matrix([x, y, z, 1])
            err = max(abs(float(q[k]) - ref[i, k]) for k in range(3))
            assert err < (5e-8 if case % 2 else 1e-11), (case, i, err)
