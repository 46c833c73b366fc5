"""The oracle restatement against the reference's OWN hot-path sources (oracle/_ref/libkmc_ref.so).

`make -C oracle ref` compiles motion_compensation.cpp, trajectory_interpolation.cpp, lie_algebra.cpp and
timestamp_mocking.cpp untouched from /root/reference (Eigen 3 if installed, else this repository's Eigen shim — see
oracle/ref_binding.py).  These tests pin oracle/kmc_oracle.cpp to that library function by function and end to end,
including the ROTATING frames for which the reference's own tests hold no golden vector.  They skip when the library
has not been built (a checkout without /root/reference); on the GPU box it arrives prebuilt.
"""
from __future__ import annotations

import numpy as np
import pytest

from oracle import binding as ob
from oracle import ref_binding as rb
import helpers as h

pytestmark = pytest.mark.skipif(not rb.available(), reason="oracle/_ref/libkmc_ref.so not built (needs /root/reference)")


def test_provider_is_recorded():
    assert rb.eigen_provider() in ("eigen3", "shim")


def test_reference_sources_reproduce_the_reference_kats():
    """The compiled reference sources give the reference's own golden numbers (test/test_motion_compensation.cpp:54-76,
    test/test_timestamp_mocking.cpp:55-57,71-73,84-86) — i.e. the build recipe did not change their behaviour."""
    k = h.kats()
    g = k["fraction_of_scan_completed"]
    for p, want in zip(g["points"], g["expected"]):
        assert rb.fraction_of_scan_completed(p) == pytest.approx(want, abs=1e-15)
    g = k["pseudo_time_stamp"]
    for p, want in zip(g["points"], g["expected"]):
        assert rb.pseudo_time_stamp(p, g["scan_start"], g["scan_end"]) == pytest.approx(want, abs=1e-15)
    g = k["motion_compensate_frame"]
    o = [h.oxts7(x) for x in g["oxts"]]
    T_start, T_end = ob.make_frame_poses(o[0], o[1], o[2], g["stamp_start"], g["stamp_end"])  # data_io.cpp:253-269
    cloud = np.array(g["cloud"], dtype=np.float64)
    ts = np.array([rb.pseudo_time_stamp(p, g["stamp_start"], g["stamp_end"]) for p in g["cloud"]])
    out = rb.motion_compensate_frame(cloud, ts, T_start, T_end, g["stamp_start"], g["stamp_end"], g["requested_time"])
    want = np.array(g["expected"])
    assert np.array_equal(out.astype(np.float32), want.astype(np.float32)) or np.allclose(out, want, rtol=5e-7, atol=0)


@pytest.mark.parametrize("seed", range(8))
def test_lie_primitives(seed):
    rng = np.random.default_rng(100 + seed)
    scale = [1.0, 1e-3, 1e-7, 2.5][seed % 4]  # includes angles below the 1e-6 Taylor switch and large ones
    phi = rng.normal(0, 1, 3) * scale
    xi = np.concatenate([rng.normal(0, 2, 3), phi])
    for name in ("hat", "so3_exp", "left_jacobian", "inverse_left_jacobian"):
        assert np.allclose(getattr(rb, name)(phi), getattr(ob, name)(phi), rtol=0, atol=1e-14), name
    R = ob.so3_exp(phi)
    assert np.allclose(rb.so3_log(R), ob.so3_log(R), rtol=0, atol=1e-14)
    assert np.allclose(rb.vee(ob.hat(phi)), phi, rtol=0, atol=0)
    T = ob.se3_exp(xi)
    assert np.allclose(rb.se3_exp(xi), T, rtol=0, atol=1e-14)
    assert np.allclose(rb.se3_log(T), ob.se3_log(T), rtol=0, atol=1e-13)


@pytest.mark.parametrize("mercator", [False, True])
def test_interpolation(mercator):
    rng = np.random.default_rng(7 + mercator)
    for _ in range(20):
        P1 = h.random_pose(rng, mercator)
        P2 = P1 @ ob.se3_exp(h.random_twist(rng) * rng.uniform(0.1, 8.0))
        t1, t2 = 100.0, 100.1
        tol = 1e-8 if mercator else 1e-13  # |t| ~ 6e6 m: both sides lose ~1e-9 m to cancellation
        for t in (t1, t2, 100.05, 100.0123):
            assert np.allclose(rb.pose_at_time(t1, P1, t2, P2, t), ob.pose_at_time(t1, P1, t2, P2, t), rtol=0, atol=tol)
        a, q = 100.05, 100.0871
        assert np.allclose(rb.relative_pose_between_times(t1, P1, t2, P2, a, q),
                           ob.relative_pose_between_times(t1, P1, t2, P2, a, q), rtol=0, atol=tol)
        p = np.array([*rng.normal(0, 30, 3), 1.0])
        assert np.allclose(rb.motion_compensate_point(t1, P1, t2, P2, q, p, a), ob.motion_compensate_point(t1, P1, t2, P2, q, p, a),
                           rtol=0, atol=tol * 100)
    with pytest.raises(ob.ReferenceWouldAbort):
        rb.pose_at_time(0.0, np.eye(4), 1.0, np.eye(4), 1.5)


def test_fraction_of_scan_edge_points():
    for p in h.edge_points():
        q = [float(p[0]), float(p[1]), float(p[2]), 1.0]
        assert rb.fraction_of_scan_completed(q) == ob.fraction_of_scan_completed(q)


def test_real_scan_config1_rotating_frame():
    """BASELINE config 1 (real 123 397-point scan, Mercator-magnitude pose, 13 m/s + 0.5 rad/s): the restatement and the
    reference sources agree to 1e-8 m (bounded by the reference's own cancellation at |t| ~ 6e6 m)."""
    pts = h.real_scan()
    T_start, T_end, t0, t1, t2 = h.config1_frame()
    for t_req in (t1, t0, t2):
        ref = rb.deskew_xyzi_scan(pts, T_start, T_end, t0, t2, t_req)
        orc = ob.deskew_xyzi_scan(pts, T_start, T_end, t0, t2, t_req)
        assert np.all(ref[:, 3] == 1.0)
        assert float(np.abs(ref - orc).max()) < 1e-8


@pytest.mark.parametrize("seed", range(4))
def test_synthetic_scans_config2(seed):
    rng = np.random.default_rng(20110926 + seed)
    pts = np.concatenate([h.synthetic_scan(20_000, seed=20110926 + seed), h.edge_points()])
    xi = h.random_twist(rng) * [1.0, 1.0, 0.0, 6.0][seed]  # seed 2: xi = 0 (identity motion); seed 3: > 1 rad per scan
    T_start = np.eye(4) if seed % 2 == 0 else h.random_pose(rng)
    T_end = T_start @ ob.se3_exp(xi)
    for t_req in (0.05, 0.0, 0.1, 0.03):
        ref = rb.deskew_xyzi_scan(pts, T_start, T_end, 0.0, 0.1, t_req)
        orc = ob.deskew_xyzi_scan(pts, T_start, T_end, 0.0, 0.1, t_req)
        assert float(np.abs(ref - orc).max()) < 1e-11
        closed = h.closed_form_deskew(pts, xi, t_req / 0.1)
        assert float(np.abs(ref[:, :3] - closed).max()) < 1e-9


def test_arbitrary_stamps_frame():
    """MotionCompensateFrame with caller-supplied per-point stamps (the FROM_W mode of the kernels)."""
    rng = np.random.default_rng(5)
    n = 5_000
    cloud = np.concatenate([rng.normal(0, 25, (n, 3)), np.ones((n, 1))], axis=1)
    ts = rng.uniform(10.0, 10.1, n)
    ts[:2] = (10.0, 10.1)
    P1 = h.random_pose(rng)
    P2 = P1 @ ob.se3_exp(h.random_twist(rng))
    ref = rb.motion_compensate_frame(cloud, ts, P1, P2, 10.0, 10.1, 10.04)
    orc = ob.motion_compensate_frame(cloud, ts, P1, P2, 10.0, 10.1, 10.04)
    assert float(np.abs(ref - orc).max()) < 1e-11
