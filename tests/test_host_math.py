"""Product host code (double, once per frame) against the oracle and the reference KATs.  No GPU needed."""

import numpy as np
import pytest

import helpers
from test_oracle_golden import artificial_pose, float_eq, matrices_same


def test_lie_matches_oracle(capi, oracle):
    rng = np.random.default_rng(11)
    for k in range(300):
        scale = [1.0, 1e-3, 1e-7, 3.0][k % 4]  # includes the < 1e-6 rad Taylor branches and large angles
        phi = rng.normal(size=3) * scale
        if np.linalg.norm(phi) > 3.0:
            phi *= 3.0 / np.linalg.norm(phi)
        xi = np.concatenate([rng.normal(size=3) * 2, phi])
        assert np.abs(capi.so3_hat(phi) - oracle.hat(phi)).max() == 0
        assert np.abs(capi.so3_exp(phi) - oracle.so3_exp(phi)).max() < 1e-15
        assert np.abs(capi.so3_left_jacobian(phi) - oracle.left_jacobian(phi)).max() < 1e-14
        assert np.abs(capi.so3_inverse_left_jacobian(phi) - oracle.inverse_left_jacobian(phi)).max() < 1e-12
        R = oracle.so3_exp(phi)
        assert np.abs(capi.so3_log(R) - oracle.so3_log(R)).max() < 1e-14
        T = oracle.se3_exp(xi)
        assert np.abs(capi.se3_exp(xi) - T).max() < 1e-14
        assert np.abs(capi.se3_log(T) - oracle.se3_log(T)).max() < 1e-11


def test_reference_lie_kats(capi, kats):
    phi, xi = kats["lie_algebra"]["phi"], kats["lie_algebra"]["xi"]
    assert all(float_eq(a, b) for a, b in zip(capi.so3_vee(capi.so3_hat(phi)), phi))
    assert all(float_eq(a, b) for a, b in zip(capi.so3_log(capi.so3_exp(phi)), phi))
    m = capi.so3_left_jacobian(phi) @ capi.so3_inverse_left_jacobian(phi)
    assert float_eq(np.trace(m), 3.0) and abs(m.sum() - np.trace(m)) < 1e-12
    assert all(float_eq(a, b) for a, b in zip(capi.se3_log(capi.se3_exp(xi)), xi))


def test_se3_log_projects_like_eigen_rotation(capi, oracle):
    """A slightly non-orthonormal linear block: Newton polar (product) == Jacobi-SVD polar (oracle)."""
    rng = np.random.default_rng(3)
    for _ in range(50):
        T = oracle.se3_exp(rng.normal(size=6) * 0.3)
        T[:3, :3] += rng.normal(size=(3, 3)) * 1e-3
        assert np.abs(capi.se3_log(T) - oracle.se3_log(T)).max() < 1e-10


def test_se3_log_rejects_reflection_and_nan(capi):
    T = np.eye(4)
    T[2, 2] = -1.0
    with pytest.raises(capi.KmcError) as e:
        capi.se3_log(T)
    assert e.value.status == capi.ERR_NOT_RIGID
    T = np.eye(4)
    T[0, 3] = np.nan
    with pytest.raises(capi.KmcError):
        capi.se3_log(T)


def test_trajectory_interpolation_kats(capi, oracle, kats):
    p = kats["trajectory_interpolation_artificial"]["poses"]
    P = [artificial_pose(oracle, q["x_rotation"], q["x_translation"]) for q in p]
    mid = capi.pose_at_time(p[0]["time"], P[0], p[2]["time"], P[2], p[1]["time"])
    assert matrices_same(oracle, mid, P[1])
    a = capi.relative_pose_between_times(p[0]["time"], P[0], p[2]["time"], P[2], p[0]["time"], p[1]["time"])
    b = capi.relative_pose_between_times(p[0]["time"], P[0], p[2]["time"], P[2], p[1]["time"], p[2]["time"])
    assert matrices_same(oracle, a, b)


def test_pose_at_time_matches_oracle(capi, oracle):
    rng = np.random.default_rng(5)
    for k in range(100):
        P1 = helpers.random_pose(rng, mercator=bool(k % 2))
        P2 = P1 @ oracle.se3_exp(helpers.random_twist(rng))
        t = rng.uniform(10.0, 10.1)
        got = capi.pose_at_time(10.0, P1, 10.1, P2, t)
        want = oracle.pose_at_time(10.0, P1, 10.1, P2, t)
        assert np.abs(got[:3, :3] - want[:3, :3]).max() < 1e-13
        assert np.abs(got[:3, 3] - want[:3, 3]).max() < 1e-7  # 6e6 m translations: 1 ulp is 1e-9
        r1 = capi.relative_pose_between_times(10.0, P1, 10.1, P2, 10.05, t)
        r2 = oracle.relative_pose_between_times(10.0, P1, 10.1, P2, 10.05, t)
        assert np.abs(r1 - r2).max() < 1e-7


def test_out_of_range_is_an_error_not_an_abort(capi):
    I = np.eye(4)
    with pytest.raises(capi.KmcError) as e:
        capi.pose_at_time(47072.3, I, 47072.5, I, 0.0)  # test_trajectory_interpolation.cpp:77-81
    assert e.value.status == capi.ERR_TIME_OUT_OF_RANGE
    with pytest.raises(capi.KmcError) as e:
        capi.frame_params_from_poses(I, I, 0.0, 0.1, 0.10001)
    assert e.value.status == capi.ERR_TIME_OUT_OF_RANGE
    with pytest.raises(capi.KmcError) as e:
        capi.frame_params_from_poses(I, I, 0.1, 0.1, 0.1)  # t_end == t_start: NaN in the reference
    assert e.value.status == capi.ERR_EMPTY_INTERVAL
    with pytest.raises(capi.KmcError) as e:
        capi.frame_params_from_twist([0, 0, 0, 0, 0, 0], 1.5)
    assert e.value.status == capi.ERR_TIME_OUT_OF_RANGE
    assert "trajectory_interpolation.cpp" in capi.last_error()


def test_fraction_and_stamp_kats(capi, kats):
    g = kats["fraction_of_scan_completed"]
    for p, want in zip(g["points"], g["expected"]):
        assert float_eq(capi.fraction_of_scan_completed(p[0], p[1]), want)
    g = kats["pseudo_time_stamp"]
    for p, want in zip(g["points"], g["expected"]):
        assert float_eq(capi.pseudo_time_stamp(p[0], p[1], g["scan_start"], g["scan_end"]), want)
    assert capi.fraction_of_scan_completed(-17.173, -0.0) == 1.0
    assert capi.fraction_of_scan_completed(-17.173, 0.0) == 0.0
    assert capi.fraction_of_scan_completed(0.0, 0.0) == 0.5


def test_frame_params_record(capi, oracle):
    """The 64-byte record is the double-precision twist of the scan, rounded to float."""
    rng = np.random.default_rng(9)
    for k in range(100):
        P1 = helpers.random_pose(rng, mercator=bool(k % 2))
        xi = helpers.random_twist(rng)
        if k % 10 == 0:
            xi[3:] = 0.0  # pure translation: the reference's own golden case
        if k % 10 == 1:
            xi[3:] *= rng.uniform(1.2, 3.0) / np.linalg.norm(xi[3:])  # > 1 rad per scan: wide path (Log is unique below pi)
        P2 = P1 @ oracle.se3_exp(xi)
        x_req = rng.uniform(0, 1)
        rec = capi.frame_params_from_poses(P1, P2, 5.0, 5.1, 5.0 + 0.1 * x_req)
        rho, phi = xi[:3], xi[3:]
        th2 = phi @ phi
        par = phi * (phi @ rho) / th2 if th2 > 0 else np.zeros(3)
        tol = 2e-6 if k % 2 else 1e-9  # Mercator poses: T_start^-1 T_end loses ~1e-9 m, x 1/|xi|
        assert np.abs(np.array(rec.phi) - phi).max() < 1e-7 * max(1.0, np.abs(phi).max())  # float32 rounding
        # rho_perp + rho_par == rho always; the split along the axis only means something when there is an axis
        # (for theta ~ 1e-17 the recovered axis is rounding noise and S == s makes the split irrelevant)
        assert np.abs(np.array(rec.rho_perp) + np.array(rec.rho_par) - rho).max() < 1e-6 + tol
        if th2 > 1e-12:
            assert np.abs(np.array(rec.rho_perp) - (rho - par)).max() < 1e-6 + tol
            assert np.abs(np.array(rec.rho_par) - par).max() < 1e-6 + tol
        assert np.abs(np.array(rec.phi_x_rho) - np.cross(phi, rho)).max() < 1e-6 * max(1.0, np.abs(np.cross(phi, rho)).max())
        assert abs(rec.theta2 - th2) < 1e-7 * max(1.0, th2)
        assert abs(rec.x_req - x_req) < 1e-6 and abs(rec.c0 - (0.5 - x_req)) < 1e-6
        assert rec.wide == (1.0 if np.float32(th2) > 1.0 else 0.0)
        rec2 = capi.frame_params_from_twist(xi, x_req)
        assert np.array_equal(np.array(rec2.phi), phi.astype(np.float32))


def test_shard_range_partitions_exactly(capi):
    for n, g in [(10000, 8), (10000, 1), (10, 3), (3, 8), (0, 4), (1, 1), (12345, 7)]:
        cover = []
        for i in range(g):
            b, e = capi.shard_range(n, g, i)
            assert 0 <= b <= e <= n
            cover.extend(range(b, e))
            sizes = e - b
            assert n // g <= sizes <= n // g + 1
        assert cover == list(range(n))
    with pytest.raises(capi.KmcError):
        capi.shard_range(10, 0, 0)


def test_synth_frame_params_are_deterministic_and_shard_independent(capi):
    a, xa = capi.synth_frame_params(16, 20110926, 0)
    b, xb = capi.synth_frame_params(8, 20110926, 8)
    assert a[8:].tobytes() == b.tobytes() and np.array_equal(xa[8:], xb)
    assert np.all(xa[:, 0] >= 0) and np.all(xa[:, 0] < 3.0) and np.abs(xa[:, 5]).max() < 0.5
    assert not np.array_equal(xa[0], xa[1])


def test_product_host_math_against_the_compiled_reference_sources(capi, oracle):
    """The product's host doubles (C ABI) against the reference's OWN lie_algebra.cpp / trajectory_interpolation.cpp /
    data_io.cpp compiled into oracle/_ref — no restatement in between.  Skipped where the library has not been built."""
    from oracle import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref/libkmc_ref.so not built (needs /root/reference)")
    rng = np.random.default_rng(19)
    for k in range(200):
        scale = [1.0, 1e-3, 1e-7, 3.0][k % 4]
        phi = rng.normal(size=3) * scale
        if np.linalg.norm(phi) > 3.0:
            phi *= 3.0 / np.linalg.norm(phi)
        xi = np.concatenate([rng.normal(size=3) * 2, phi])
        assert np.abs(capi.so3_exp(phi) - rb.so3_exp(phi)).max() < 1e-15
        assert np.abs(capi.so3_left_jacobian(phi) - rb.left_jacobian(phi)).max() < 1e-14
        assert np.abs(capi.so3_inverse_left_jacobian(phi) - rb.inverse_left_jacobian(phi)).max() < 1e-12
        T = rb.se3_exp(xi)
        assert np.abs(capi.se3_exp(xi) - T).max() < 1e-14
        assert np.abs(capi.se3_log(T) - rb.se3_log(T)).max() < 1e-11
    for mercator in (False, True):
        tol = 1e-8 if mercator else 1e-12
        for _ in range(20):
            P1 = helpers.random_pose(rng, mercator)
            P2 = P1 @ oracle.se3_exp(helpers.random_twist(rng) * rng.uniform(0.1, 6.0))
            for t in (5.0, 5.1, 5.03, 5.0999):
                assert np.abs(capi.pose_at_time(5.0, P1, 5.1, P2, t) - rb.pose_at_time(5.0, P1, 5.1, P2, t)).max() < tol
            got = capi.relative_pose_between_times(5.0, P1, 5.1, P2, 5.05, 5.0123)
            assert np.abs(got - rb.relative_pose_between_times(5.0, P1, 5.1, P2, 5.05, 5.0123)).max() < tol
    for _ in range(20):
        o = [0.0, 49.0 + rng.normal(0, 1), 8.4 + rng.normal(0, 1), 110 + rng.normal(0, 10), *rng.normal(0, 0.1, 2), rng.uniform(-3.1, 3.1)]
        assert np.abs(capi.oxts_to_pose(*o[1:], 0.8) - rb.oxts_to_pose(o, 0.8)).max() < 2e-9  # translations of ~6e6 m
