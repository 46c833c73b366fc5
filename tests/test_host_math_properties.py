"""Algebraic properties of the product's once-per-frame host math (C ABI, no GPU, no oracle): what must hold for ANY input, drawn
by hypothesis.  The reference states the same identities in its own tests for a handful of fixed values
(test/test_lie_algebra.cpp:8-90 round trips, test/test_trajectory_interpolation.cpp:43-75 midpoint / equal relative poses);
here they run over the whole input domain, including Mercator-magnitude poses (translations of 10^6 m) and rotations up to pi."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, assume, given, settings
from hypothesis import strategies as st

import helpers

# deterministic in the suite (same examples every run, nothing written to disk); KMC_HYPOTHESIS_RANDOM=1 [KMC_HYPOTHESIS_EXAMPLES=n]
# draws fresh ones for a soak run
SETTINGS = dict(max_examples=int(os.environ.get("KMC_HYPOTHESIS_EXAMPLES", "150")), deadline=None, database=None,
                derandomize=os.environ.get("KMC_HYPOTHESIS_RANDOM", "0") != "1",
                suppress_health_check=[HealthCheck.function_scoped_fixture])

angle = st.floats(min_value=0.0, max_value=3.1, allow_nan=False)
unit = st.tuples(st.floats(-1, 1), st.floats(-1, 1), st.floats(-1, 1)).filter(lambda v: 1e-3 < np.linalg.norm(v) < 2.0)
shift = st.tuples(st.floats(-50, 50), st.floats(-50, 50), st.floats(-50, 50))
seeds = st.integers(min_value=0, max_value=2**31 - 1)
fraction = st.floats(min_value=0.0, max_value=1.0, allow_nan=False)


def twist(rho, axis, theta):
    a = np.asarray(axis, dtype=np.float64)
    return np.concatenate([np.asarray(rho, dtype=np.float64), a / np.linalg.norm(a) * theta])


def rigid_error(T):
    R = T[:3, :3]
    return max(np.abs(R @ R.T - np.eye(3)).max(), abs(np.linalg.det(R) - 1.0), np.abs(T[3] - [0, 0, 0, 1]).max())


@settings(**SETTINGS)
@given(rho=shift, axis=unit, theta=angle)
def test_exp_gives_a_rigid_transform_and_log_inverts_it(capi, rho, axis, theta):
    xi = twist(rho, axis, theta)
    T = capi.se3_exp(xi)
    assert rigid_error(T) < 1e-12
    back = capi.se3_log(T)
    assert np.abs(back - xi).max() < 1e-8 * (1.0 + np.abs(xi).max())
    # Exp(-xi) is the inverse
    assert np.abs(capi.se3_exp(-xi) @ T - np.eye(4)).max() < 1e-10 * (1.0 + np.abs(xi[:3]).max())


@settings(**SETTINGS)
@given(axis=unit, theta=angle)
def test_so3_exp_log_and_jacobians(capi, axis, theta):
    phi = twist((0, 0, 0), axis, theta)[3:]
    R = capi.so3_exp(phi)
    assert np.abs(R @ R.T - np.eye(3)).max() < 1e-13 and abs(np.linalg.det(R) - 1) < 1e-13
    assert np.abs(capi.so3_log(R) - phi).max() < 1e-8
    assert np.abs(capi.so3_vee(capi.so3_hat(phi)) - phi).max() == 0.0
    J, Jinv = capi.so3_left_jacobian(phi), capi.so3_inverse_left_jacobian(phi)
    assert np.abs(J @ Jinv - np.eye(3)).max() < 1e-9
    assert np.abs(R @ phi - phi).max() < 1e-12  # the axis is fixed by the rotation


@settings(**SETTINGS)
@given(seed=seeds, rho=shift, axis=unit, theta=st.floats(0.0, 1.0), mercator=st.booleans(), a=fraction, b=fraction, c=fraction)
def test_interpolation_is_a_one_parameter_group(capi, seed, rho, axis, theta, mercator, a, b, c):
    """GetPoseAtTime(t) = P1 Exp(x Log(P1^-1 P2)) (trajectory_interpolation.cpp:31-41): end points reproduce the poses,
    RelativePoseBetweenTimes(a, a) = I, (a->b)(b->c) = (a->c), (a->b)^-1 = (b->a), and none of it depends on where on the globe
    the trajectory sits (only P1^-1 P2 enters)."""
    P1 = helpers.random_pose(np.random.default_rng(seed), mercator=mercator)
    P2 = P1 @ capi.se3_exp(twist(rho, axis, theta))
    t1, t2 = 100.0, 100.1
    at = lambda x: t1 + x * (t2 - t1)  # noqa: E731
    scale = 1.0 + np.abs(P1[:3, 3]).max()
    assert np.abs(capi.pose_at_time(t1, P1, t2, P2, t1) - P1).max() < 1e-15 * scale + 1e-12
    assert np.abs(capi.pose_at_time(t1, P1, t2, P2, t2) - P2).max() < 1e-15 * scale * 8 + 1e-9
    rel = lambda x, y: capi.relative_pose_between_times(t1, P1, t2, P2, at(x), at(y))  # noqa: E731
    tol = 2e-9 * (1.0 + 50.0) + 4e-16 * scale * 50  # rounding of the two absolute poses that are differenced
    assert np.abs(rel(a, a) - np.eye(4)).max() < tol
    ab, bc, ac, ba = rel(a, b), rel(b, c), rel(a, c), rel(b, a)
    assert rigid_error(ab) < 1e-9
    assert np.abs(ab @ bc - ac).max() < tol
    assert np.abs(ab @ ba - np.eye(4)).max() < tol


@settings(**SETTINGS)
@given(seed=seeds, rho=shift, axis=unit, theta=st.one_of(st.just(0.0), st.floats(1e-6, 0.6)), x_req=fraction)
def test_frame_record_depends_on_the_relative_motion_only(capi, seed, rho, axis, theta, x_req):
    """The 64-byte per-frame record is built from Log(T_start^-1 T_end) (motion_compensation.cpp:16-28 needs nothing else):
    moving the whole trajectory by any rigid transform — the origin versus a Mercator-magnitude pose — leaves it unchanged to
    float precision, and it equals the record built from the twist directly."""
    xi = twist(rho, axis, theta)
    G = helpers.random_pose(np.random.default_rng(seed), mercator=True)
    t1, t2 = 5.0, 5.1
    here = capi.frame_params_from_poses(np.eye(4), capi.se3_exp(xi), t1, t2, t1 + x_req * (t2 - t1))
    there = capi.frame_params_from_poses(G, G @ capi.se3_exp(xi), t1, t2, t1 + x_req * (t2 - t1))
    direct = capi.frame_params_from_twist(xi, x_req)
    as_floats = lambda p: np.frombuffer(bytes(p), dtype=np.float32)  # noqa: E731
    h, t, d = as_floats(here), as_floats(there), as_floats(direct)
    assert np.abs(h - d).max() < 2e-6 * (1.0 + np.abs(d).max())
    # 1e6 m translations cancel in T_start^-1 T_end with ~1e-9 m left over: far below one float ulp of a 50 m/frame motion.
    # (theta = 0 is the pure translation: phi comes out of the pose product as ~1e-17 of noise and must not define an axis.)
    assert np.abs(t - d).max() < 2e-6 * (1.0 + np.abs(d).max())
    assert np.abs((t[4:7] + t[8:11]) - xi[:3]).max() < 1e-6 * (1.0 + np.abs(xi[:3]).max())  # rho_perp + rho_par = rho


@settings(**SETTINGS)
@given(x=st.floats(-200, 200, allow_nan=False), y=st.floats(-200, 200, allow_nan=False), k=st.floats(1e-3, 1e3))
def test_fraction_of_scan_is_a_function_of_the_direction(capi, x, y, k):
    """FractionOfScanCompleted (timestamp_mocking.cpp:46-54) = (pi - atan2(y, x)) / 2 pi: in [0, 1], scale invariant,
    and a half turn apart for opposite directions."""
    assume(max(abs(x), abs(y)) > 1e-290)  # the origin has no direction, and a denormal scaled by k loses it
    f = capi.fraction_of_scan_completed(x, y)
    assert 0.0 <= f <= 1.0
    assert f == pytest.approx((np.pi - np.arctan2(y, x)) / (2 * np.pi), abs=1e-15)
    assert capi.fraction_of_scan_completed(k * x, k * y) == pytest.approx(f, abs=1e-12)
    g = capi.fraction_of_scan_completed(-x, -y)
    assert abs(abs(f - g) - 0.5) < 1e-12
    s = capi.pseudo_time_stamp(x, y, 10.0, 10.1)
    assert 10.0 <= s <= 10.1 and s == pytest.approx(10.0 + f * 0.1, abs=1e-12)
