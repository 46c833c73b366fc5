"""C++ mirror of the reference API (include/kitti_motion_compensation/*.hpp + libkitti_motion_compensation_lib.so):
the reference's own gtest bodies, re-expressed in tests/cpp/*.cpp, compiled with g++ and run here."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REAL_SCAN = os.path.join(ROOT, "tests", "golden", "kitti_2011_09_26_drive_0005_frame0.bin")


def run_binary(path, *args):
    r = subprocess.run([path, *args], capture_output=True, text=True, timeout=600)
    print(r.stdout[-4000:])
    print(r.stderr[-4000:])
    return r


def test_dropin_library_exports_the_reference_symbols():
    from kitti_motion_compensation_b200 import build
    lib = build.build_dropin()
    out = subprocess.run(["nm", "-DC", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    for sym in ["kmc::MotionCompensateFrame(kmc::Frame const&, double)",
                "kmc::MotionCompensatePoint(kmc::trajectory_interpolation::TrajectoryInterpolator const&, double,",
                "kmc::GetPseudoTimeStamps(", "kmc::GetPseudoTimeStamp(", "kmc::FractionOfScanCompleted(",
                "kmc::trajectory_interpolation::TrajectoryInterpolator::GetPoseAtTime(double) const",
                "kmc::trajectory_interpolation::TrajectoryInterpolator::RelativePoseBetweenTimes(double, double) const",
                "kmc::trajectory_interpolation::InterpolateTrajectory(kmc::Oxts const&, kmc::Oxts const&, double)",
                "kmc::lie::Hat(", "kmc::lie::Vee(", "kmc::lie::Exp(", "kmc::lie::Log(", "kmc::lie::LeftJacobian(",
                "kmc::lie::InverseLeftJacobian(", "kmc::OxtsToPose(kmc::Oxts const&, double)", "kmc::MakeFrame(",
                "kmc::KittiPclLoader::LoadPointcloud(", "kmc::WritePointcloud(", "kmc::MotionCompensateRun(",
                "kmc::LoadLidarExtrinsics(", "kmc::viz::LoadCameraCalibrations(", "kmc::viz::CalibrationLinesToCalibration(",
                "kmc::viz::ProjectPointcloudOnCamera("]:
        assert sym in out, sym


def test_cpp_mirror_host():
    """test/test_lie_algebra.cpp, artificial-pose + abort tests of test/test_trajectory_interpolation.cpp,
    test/test_oxts_to_pose.cpp — host doubles, no GPU."""
    from kitti_motion_compensation_b200 import build
    r = run_binary(build.build_cpp_test("test_dropin_host"))
    assert r.returncode == 0
    assert " 0 failed" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_gpu():
    """test/test_motion_compensation.cpp, test/test_timestamp_mocking.cpp, loader/writer round trip, a generated run
    through MotionCompensateRun and the DataHandle float path — all through the CUDA kernels."""
    from kitti_motion_compensation_b200 import build
    r = run_binary(build.build_cpp_test("test_dropin_gpu"), "", REAL_SCAN)
    assert r.returncode == 0
    assert " 0 failed" in r.stdout
    assert "RealScanTest.LoadAndMotionCompensate" in r.stdout


def test_motion_compensate_frame_has_no_cpu_fallback():
    """On a box without a GPU the C++ MotionCompensateFrame must raise (std::runtime_error -> terminate), not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from kitti_motion_compensation_b200 import build
    r = run_binary(build.build_cpp_test("test_dropin_gpu"), "DataHandleTest", REAL_SCAN)
    assert r.returncode != 0
    assert "no usable CUDA device" in r.stderr or "CUDA" in r.stderr


def test_eigen_shim_semantics():
    """The fallback Eigen subset (column-major storage, comma initialiser, row proxies, Affine-mode inverse / rotation())."""
    from kitti_motion_compensation_b200 import build
    tests = os.path.join(ROOT, "tests", "cpp")
    out = os.path.join(tests, "_build", "test_eigen_shim")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run([build._cxx(), "-std=c++17", "-O1", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-I", tests, "-o", out, os.path.join(tests, "test_eigen_shim.cpp")], check=True)
    r = run_binary(out)
    assert r.returncode == 0 and " 0 failed" in r.stdout
