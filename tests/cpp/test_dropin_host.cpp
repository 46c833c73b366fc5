// Host-only part of the C++ mirror tests: the reference's test/test_lie_algebra.cpp and the artificial-pose half of
// test/test_trajectory_interpolation.cpp, re-expressed against this repository's drop-in headers.  No GPU needed.
#include <filesystem>
#include <fstream>

#include "kitti_motion_compensation/camera_model.hpp"
#include "kitti_motion_compensation/data_io.hpp"
#include "kitti_motion_compensation/lie_algebra.hpp"
#include "kitti_motion_compensation/trajectory_interpolation.hpp"
#include "kitti_motion_compensation/utilities_for_testing.hpp"
#include "kitti_motion_compensation/utils.hpp"
#include "mini_gtest.hpp"

using namespace kmc;

// ---- test/test_lie_algebra.cpp:5-47 ------------------------------------------------------------------------------
TEST(LieAlgebraTest, HatAndVeeInverses) {
  Eigen::Vector3d const phi_in{0.1, 0.2, 0.3};
  Eigen::Vector3d const phi_out{kmc::lie::Vee(kmc::lie::Hat(phi_in))};
  for (int i = 0; i < 3; ++i) ASSERT_FLOAT_EQ(phi_in(i), phi_out(i));
}

TEST(LieAlgebraTest, So3LogAndExpInverse) {
  Eigen::Vector3d const phi_in{0.1, 0.2, 0.3};
  Eigen::Vector3d const phi_out{kmc::lie::Log(kmc::lie::Exp(phi_in))};
  for (int i = 0; i < 3; ++i) ASSERT_FLOAT_EQ(phi_in(i), phi_out(i));
}

TEST(LieAlgebraTest, So3LeftJacobiansInverse) {
  Eigen::Vector3d const phi_in{0.1, 0.2, 0.3};
  Eigen::Matrix3d const j{kmc::lie::LeftJacobian(phi_in)};
  Eigen::Matrix3d const j_inv{kmc::lie::InverseLeftJacobian(phi_in)};
  auto const identity{j * j_inv};
  ASSERT_FLOAT_EQ(identity.trace(), 3.0);
  ASSERT_NEAR(identity.sum() - identity.trace(), 0.0, 1e-12);
}

TEST(LieAlgebraTest, Se3LogAndExpInverse) {
  kmc::Twist xi_in;
  xi_in << 0.1, 0.2, 0.3, 0.4, 0.5, 0.6;
  kmc::Twist const xi_out{kmc::lie::Log(kmc::lie::Exp(xi_in))};
  for (int i = 0; i < 6; ++i) ASSERT_FLOAT_EQ(xi_in(i), xi_out(i));
}

TEST(LieAlgebraTest, TwistOrderingIsRhoThenPhi) {
  kmc::Twist xi;
  xi << 1.0, 2.0, 3.0, 0.0, 0.0, 0.0;  // pure translation
  Eigen::Affine3d const T{kmc::lie::Exp(xi)};
  ASSERT_FLOAT_EQ(T.translation()(0), 1.0);
  ASSERT_FLOAT_EQ(T.translation()(1), 2.0);
  ASSERT_FLOAT_EQ(T.translation()(2), 3.0);
  ASSERT_FLOAT_EQ(T.linear().trace(), 3.0);
}

// ---- test/test_trajectory_interpolation.cpp:10-60 ------------------------------------------------------------------
class TrajectoryInterpolationFixtureArtificialPoses : public ::testing::Test {
 protected:
  void SetUp() override {
    time_0_ = 0;
    pose_0_ = ArtificialPose(0, 0);
    time_1_ = 50;
    pose_1_ = ArtificialPose(0.5, 0.5);
    time_2_ = 100;
    pose_2_ = ArtificialPose(1.0, 1.0);
  }

  static Affine3d ArtificialPose(double const x_rotation, double const x_translation) {
    Eigen::Affine3d pose{Affine3d::Identity()};
    pose.rotate(Eigen::AngleAxisd{x_rotation, Eigen::Vector3d::UnitX()});
    pose.translation() = Eigen::Vector3d{x_translation, 0, 0};
    return pose;
  }

  Time time_0_, time_1_, time_2_;
  Affine3d pose_0_{Affine3d::Identity()}, pose_1_{Affine3d::Identity()}, pose_2_{Affine3d::Identity()};
};

TEST_F(TrajectoryInterpolationFixtureArtificialPoses, TestInterpolationClassPoseConstructor) {
  auto const interpolator{trajectory_interpolation::TrajectoryInterpolator(time_0_, pose_0_, time_2_, pose_2_)};
  Affine3d const interpolated_pose_1{interpolator.GetPoseAtTime(time_1_)};
  ASSERT_TRUE(utilities_for_testing::TransformationMatricesAreTheSame(interpolated_pose_1, pose_1_));
}

TEST_F(TrajectoryInterpolationFixtureArtificialPoses, TestRelativePoseBetweenTimes) {
  auto const interpolator{trajectory_interpolation::TrajectoryInterpolator(time_0_, pose_0_, time_2_, pose_2_)};
  auto const tf_0_1{interpolator.RelativePoseBetweenTimes(time_0_, time_1_)};
  auto const tf_1_2{interpolator.RelativePoseBetweenTimes(time_1_, time_2_)};
  ASSERT_TRUE(utilities_for_testing::TransformationMatricesAreTheSame(tf_0_1, tf_1_2));
}

// test/test_trajectory_interpolation.cpp:77-81 — the abort contract (the packets there are stamped ~47072 s)
TEST_F(TrajectoryInterpolationFixtureArtificialPoses, TestOutOfRangeTime) {
  auto const interpolator{trajectory_interpolation::TrajectoryInterpolator(47072.35, pose_0_, 47072.56, pose_2_)};
  EXPECT_DEATH(interpolator.GetPoseAtTime(0), "c");
  EXPECT_DEATH(interpolator.RelativePoseBetweenTimes(47072.4, 47072.6), "c");
}

// ---- test/test_oxts_to_pose.cpp:8-21 (values of the shipped packet 0) ----------------------------------------------
TEST(OxtsToPoseTest, LoadKnownPoseProperly) {
  Oxts const oxts{47072.349659964, 49.011212804408, 8.4228850417969, 112.83492279053, 0.022447, 1e-05, -1.2219096732051,
                  1.1384311814592, 3.5147680214713, 0.037625160413037};
  auto const pose{OxtsToPose(oxts, 1.0)};
  ASSERT_FLOAT_EQ(pose.rotation().determinant(), 1.0);
  ASSERT_FLOAT_EQ(pose.translation().x(), 937631.25);
  ASSERT_FLOAT_EQ(pose.translation().y(), 6276764);
  ASSERT_FLOAT_EQ(pose.translation().z(), 112.83492);
}

TEST(OxtsInterpolationTest, InterpolateTrajectoryBetweenPackets) {
  // test/test_motion_compensation.cpp:29-31 fixture: 1e-5 deg of longitude per 0.1 s = 1.11319 m east
  Oxts const o0{Time(0.05), 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  Oxts const o1{Time(0.15), 0.0, 0.00001, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  Affine3d const mid{trajectory_interpolation::InterpolateTrajectory(o0, o1, 0.1)};
  ASSERT_FLOAT_EQ(mid.translation().x(), 0.5 * 1.1131949);
  ASSERT_NEAR(mid.translation().y(), 0.0, 1e-8);
}

// ---- utils ------------------------------------------------------------------------------------------------------------
TEST(UtilsTest, StringHelpers) {
  ASSERT_TRUE(IdToZeroPaddedString(15) == "0000000015");
  ASSERT_TRUE(IdToZeroPaddedString(123, 2) == "123");
  auto const tokens{TokenizeString("2011-09-26 13:04:32.283701593")};
  ASSERT_EQ(tokens.size(), size_t{2});
  ASSERT_NEAR(MmHhSsToSeconds(tokens[1]), 47072.283701593, 1e-9);  // test/test_data_io.cpp:49
}

// ---- calibration parsers (data_io.cpp:168-210, 321-406) on files with the reference's calibration values -----------------
static std::filesystem::path WriteCalibrationFolder() {
  namespace fs = std::filesystem;
  fs::path const dir{fs::temp_directory_path() / "kmc_b200_test_calib"};
  fs::create_directories(dir);
  std::ofstream velo(dir / "calib_velo_to_cam.txt");
  velo << "calib_time: 15-Mar-2012 11:37:16\n"
       << "R: 7.533745e-03 -9.999714e-01 -6.166020e-04 1.480249e-02 7.280733e-04 -9.998902e-01 9.998621e-01 7.523790e-03 1.480755e-02\n"
       << "T: -4.069766e-03 -7.631618e-02 -2.717806e-01\n"
       << "delta_f: 0.000000e+00 0.000000e+00\n";
  std::ofstream cam(dir / "calib_cam_to_cam.txt");
  cam << "calib_time: 09-Jan-2012 13:57:47\ncorner_dist: 9.950000e-02\n";
  const char* p_rect[4] = {"0.000000e+00 0.000000e+00 7.215377e+02 1.728540e+02 0.000000e+00", "-3.875744e+02 0.000000e+00 7.215377e+02 1.728540e+02 0.000000e+00",
                           "4.485728e+01 0.000000e+00 7.215377e+02 1.728540e+02 2.163791e-01", "-3.395242e+02 0.000000e+00 7.215377e+02 1.728540e+02 2.199936e+00"};
  const char* p_tail[4] = {"0.000000e+00", "0.000000e+00", "2.745884e-03", "2.729905e-03"};
  for (int c = 0; c < 4; ++c) {
    cam << "S_0" << c << ": 1.392000e+03 5.120000e+02\n"
        << "K_0" << c << ": 9.842439e+02 0.000000e+00 6.900000e+02 0.000000e+00 9.808141e+02 2.331966e+02 0.000000e+00 0.000000e+00 1.000000e+00\n"
        << "D_0" << c << ": -3.728755e-01 2.037299e-01 2.219027e-03 1.383707e-03 -7.233722e-02\n"
        << "R_0" << c << ": 1.000000e+00 0.000000e+00 0.000000e+00 0.000000e+00 1.000000e+00 0.000000e+00 0.000000e+00 0.000000e+00 1.000000e+00\n"
        << "T_0" << c << ": " << -0.5 * c << " 0.000000e+00 0.000000e+00\n"
        << "S_rect_0" << c << ": 1.242000e+03 3.750000e+02\n"
        << "R_rect_0" << c << ": 9.999239e-01 9.837760e-03 -7.445048e-03 -9.869795e-03 9.999421e-01 -4.278459e-03 7.402527e-03 4.351614e-03 9.999631e-01\n"
        << "P_rect_0" << c << ": 7.215377e+02 0.000000e+00 6.095593e+02 " << p_rect[c] << " 0.000000e+00 0.000000e+00 1.000000e+00 " << p_tail[c] << "\n";
  }
  return dir;
}

TEST(CalibrationIoTest, LoadLidarExtrinsicsAndCameraCalibrations) {
  auto const dir{WriteCalibrationFolder()};
  Eigen::Affine3d const tf{LoadLidarExtrinsics(dir, true)};
  ASSERT_FLOAT_EQ(tf.linear()(0, 1), -9.999714e-01);
  ASSERT_FLOAT_EQ(tf.linear()(2, 0), 9.998621e-01);
  ASSERT_FLOAT_EQ(tf.translation()(2), -2.717806e-01);
  viz::CameraCalibrations const cams{viz::LoadCameraCalibrations(dir)};
  ASSERT_FLOAT_EQ(cams.camera_00.S_rect(0), 1242.0);
  ASSERT_FLOAT_EQ(cams.camera_00.R_rect(0, 1), 9.837760e-03);
  ASSERT_FLOAT_EQ(cams.camera_00.R_rect(1, 0), -9.869795e-03);
  ASSERT_FLOAT_EQ(cams.camera_02.P_rect(0, 0), 7.215377e+02);
  ASSERT_FLOAT_EQ(cams.camera_02.P_rect(0, 2), 6.095593e+02);
  ASSERT_FLOAT_EQ(cams.camera_02.P_rect(0, 3), 4.485728e+01);
  ASSERT_FLOAT_EQ(cams.camera_02.P_rect(1, 3), 2.163791e-01);
  ASSERT_FLOAT_EQ(cams.camera_02.P_rect(2, 3), 2.745884e-03);
  ASSERT_FLOAT_EQ(cams.camera_03.P_rect(0, 3), -3.395242e+02);
  ASSERT_FLOAT_EQ(cams.camera_01.T(0), -0.5);
  ASSERT_FLOAT_EQ(cams.camera_00.D(4), -7.233722e-02);
  std::filesystem::remove_all(dir);
}

int main(int argc, char** argv) {
  testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
