// GPU part of the C++ mirror tests: the reference's test/test_motion_compensation.cpp and
// test/test_timestamp_mocking.cpp (whose fixture builds a Frame, i.e. runs GetPseudoTimeStamps), the loader/writer
// round trip of test/test_data_io.cpp:115-149, and a whole MotionCompensateRun on a generated run folder —
// all through this repository's drop-in headers, i.e. through the CUDA kernels.  argv[2] = path of the real scan.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iterator>
#include <memory>
#include <random>
#include <thread>
#include <vector>

#include "kitti_motion_compensation/camera_model.hpp"
#include "kitti_motion_compensation/data_handle.hpp"
#include "kitti_motion_compensation/data_io.hpp"
#include "kitti_motion_compensation/data_types.hpp"
#include "kitti_motion_compensation/handlers.hpp"
#include "kitti_motion_compensation/lie_algebra.hpp"
#include "kitti_motion_compensation/motion_compensation.hpp"
#include "kitti_motion_compensation/timestamp_mocking.hpp"
#include "kitti_motion_compensation/utils.hpp"
#include "mini_gtest.hpp"

using namespace kmc;
namespace fs = std::filesystem;

static std::string g_real_scan;

// test/test_motion_compensation.cpp:10-52 / test/test_timestamp_mocking.cpp:10-48
static Frame MakeMotionCompensationTestFrame() {
  Oxts const odometry_0{Time(0.05), 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  Oxts const odometry_1{Time(0.15), 0.0, 0.00001, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  Oxts const odometry_2{Time(0.25), 0.0, 0.00002, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};

  Pointcloud cloud_1 = MatrixX4d(3, 4);
  cloud_1.row(0) = Vector4d{0.0, 5.0, 0.0, 1.0};
  cloud_1.row(1) = Vector4d{5.0, 0.0, 0.0, 1.0};
  cloud_1.row(2) = Vector4d{0.0, -5.0, 0.0, 1.0};

  Time const stamp_start{Time(0.1)};
  Time const stamp_middle{Time(0.15)};
  Time const stamp_end{Time(0.2)};

  VectorXd const timestamps_1 = GetPseudoTimeStamps(cloud_1, stamp_start, stamp_end);
  VectorXd const intensities_1 = VectorXd(3);

  LidarScan const scan_1{stamp_start, stamp_middle, stamp_end, cloud_1, intensities_1, timestamps_1};
  return MakeFrame(odometry_0, odometry_1, odometry_2, scan_1);
}

class TestFrameFixture : public ::testing::Test {
 protected:
  void SetUp() override { frame_ = std::make_unique<Frame>(MakeMotionCompensationTestFrame()); }
  std::unique_ptr<Frame> frame_;
};

// test/test_motion_compensation.cpp:54-76
TEST_F(TestFrameFixture, MotionCompensateFrame) {
  Time const requested_time{frame_->scan.stamp_middle};
  Pointcloud const motion_compensated_cloud{MotionCompensateFrame(*frame_, requested_time)};

  Vector4d const point_1{motion_compensated_cloud.row(0)};
  ASSERT_FLOAT_EQ(point_1(0), -0.27829874);
  ASSERT_FLOAT_EQ(point_1(1), 5.0);
  ASSERT_FLOAT_EQ(point_1(2), 0.0);
  ASSERT_FLOAT_EQ(point_1(3), 1.0);

  Vector4d const point_2{motion_compensated_cloud.row(1)};
  for (int c = 0; c < 4; ++c) ASSERT_FLOAT_EQ(point_2(c), frame_->scan.cloud.row(1)(c));

  Vector4d const point_3{motion_compensated_cloud.row(2)};
  ASSERT_FLOAT_EQ(point_3(0), 0.27829874);
  ASSERT_FLOAT_EQ(point_3(1), -5.0);
  ASSERT_FLOAT_EQ(point_3(2), 0.0);
  ASSERT_FLOAT_EQ(point_3(3), 1.0);
}

TEST_F(TestFrameFixture, FramePathAgreesWithSinglePointPath) {
  Time const requested_time{frame_->scan.stamp_middle};
  Pointcloud const cloud{MotionCompensateFrame(*frame_, requested_time)};
  TrajectoryInterpolator const interpolator(frame_->scan.stamp_start, frame_->T_start, frame_->scan.stamp_end, frame_->T_end);
  for (Index i = 0; i < 3; ++i) {
    Vector4d const p{MotionCompensatePoint(interpolator, frame_->scan.timestamps(i), frame_->scan.cloud.row(i), requested_time)};
    for (int c = 0; c < 4; ++c) ASSERT_NEAR(cloud.row(i)(c), p(c), 1e-6);
  }
}

TEST_F(TestFrameFixture, OutOfRangeRequestedTimeAborts) {
  EXPECT_DEATH(MotionCompensateFrame(*frame_, 0.3), "c");
  EXPECT_DEATH(MotionCompensateFrame(*frame_, 0.0), "c");
}

TEST_F(TestFrameFixture, OutOfRangePointStampAborts) {
  Frame bad{*frame_};
  bad.scan.timestamps(1) = 0.25;  // the reference asserts inside GetPoseAtTime(point_stamp)
  EXPECT_DEATH(MotionCompensateFrame(bad, 0.15), "c");
}

// test/test_timestamp_mocking.cpp:50-87
TEST(FractionOfScanCompletedTest, XXX) {
  Frame const test_frame{MakeMotionCompensationTestFrame()};
  Pointcloud const& test_cloud{test_frame.scan.cloud};
  ASSERT_FLOAT_EQ(FractionOfScanCompleted(test_cloud.row(0)), 0.25);
  ASSERT_FLOAT_EQ(FractionOfScanCompleted(test_cloud.row(1)), 0.5);
  ASSERT_FLOAT_EQ(FractionOfScanCompleted(test_cloud.row(2)), 0.75);
}

TEST(PsuedoTimeStampTest, XXX) {
  Frame const test_frame{MakeMotionCompensationTestFrame()};
  Pointcloud const& test_cloud{test_frame.scan.cloud};
  Time const scan_start{test_frame.scan.stamp_start};
  Time const scan_end{test_frame.scan.stamp_end};
  ASSERT_FLOAT_EQ(GetPseudoTimeStamp(test_cloud.row(0), scan_start, scan_end), 0.125);
  ASSERT_FLOAT_EQ(GetPseudoTimeStamp(test_cloud.row(1), scan_start, scan_end), 0.15);
  ASSERT_FLOAT_EQ(GetPseudoTimeStamp(test_cloud.row(2), scan_start, scan_end), 0.175);
}

TEST(PsuedoTimeStampFrameInitializationTest, XXX) {
  Frame const test_frame{MakeMotionCompensationTestFrame()};  // stamps come from the CUDA GetPseudoTimeStamps
  ASSERT_FLOAT_EQ(test_frame.scan.timestamps(0), 0.125);
  ASSERT_FLOAT_EQ(test_frame.scan.timestamps(1), 0.15);
  ASSERT_FLOAT_EQ(test_frame.scan.timestamps(2), 0.175);
}

// ---- real scan through the reference-shaped API ----------------------------------------------------------------------
// test/test_data_io.cpp:53-78 (loader expectations) + MotionCompensateFrame vs the double single-point path
TEST(RealScanTest, LoadAndMotionCompensate) {
  KittiPclLoader loader;
  auto const [cloud, intensities] = loader.LoadPointcloud(g_real_scan);
  ASSERT_EQ(cloud.rows(), 123397);
  ASSERT_FLOAT_EQ(cloud.row(0)(0), 22.719);
  ASSERT_FLOAT_EQ(cloud.row(0)(1), 0.031);
  ASSERT_FLOAT_EQ(cloud.row(0)(2), 0.977);
  ASSERT_FLOAT_EQ(intensities(0), 0.32);
  ASSERT_FLOAT_EQ(cloud.row(123396)(0), 5.634);
  ASSERT_FLOAT_EQ(cloud.row(123396)(3), 1.0);

  Time const start{47072.283701593}, middle{47072.335337762}, end{47072.386973931};
  VectorXd const stamps{GetPseudoTimeStamps(cloud, start, end)};
  Oxts const oxts{47072.349659964, 49.011212804408, 8.4228850417969, 112.83492279053, 0.022447, 1e-05, -1.2219096732051, 0, 0, 0};
  Affine3d const T_start{OxtsToPose(oxts)};
  Twist xi;
  xi << 1.34, 0.03, -0.01, -0.003, 0.004, 0.05;
  Affine3d const T_end{T_start * lie::Exp(xi)};
  LidarScan const scan{start, middle, end, cloud, intensities, stamps};
  Frame const frame(T_start, T_end, scan);
  Pointcloud const out{MotionCompensateFrame(frame, middle)};
  ASSERT_EQ(out.rows(), cloud.rows());

  TrajectoryInterpolator const interpolator(start, T_start, end, T_end);
  double worst = 0;
  for (Index i = 0; i < cloud.rows(); i += 997) {
    Vector4d const p{MotionCompensatePoint(interpolator, stamps(i), cloud.row(i), middle)};
    for (int c = 0; c < 3; ++c) worst = std::fmax(worst, std::fabs(out.row(i)(c) - p(c)));
    ASSERT_FLOAT_EQ(out.row(i)(3), 1.0);
  }
  std::printf("    max |dxyz| vs double single-point path: %.3e m\n", worst);
  ASSERT_TRUE(worst < 1e-5);
  // wall time of the reference-shaped call (H2D of the double cloud + stamps, kernel, D2H), after the warm-up above
  auto const t_begin = std::chrono::steady_clock::now();
  int const reps = 20;
  for (int r = 0; r < reps; ++r) {
    Pointcloud const again{MotionCompensateFrame(frame, middle)};
    ASSERT_EQ(again.rows(), cloud.rows());
  }
  double const ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count() / reps;
  std::printf("    kmc::MotionCompensateFrame(Frame, t): %.3f ms per 123 397-point frame = %.1f Mpoints/s\n", ms, 123397.0 / ms / 1e3);
}

// The reference's MotionCompensateFrame touches only its arguments, i.e. it is re-entrant; here concurrent callers lease
// separate handles and share the host thread pool.  Six threads, different requested times and a non-homogeneous column in
// one of them: every result equals the same call made alone.
TEST(RealScanTest, ConcurrentCallersGetTheSingleCallerResult) {
  KittiPclLoader loader;
  auto [cloud, intensities] = loader.LoadPointcloud(g_real_scan);
  Time const start{47072.283701593}, middle{47072.335337762}, end{47072.386973931};
  VectorXd const stamps{GetPseudoTimeStamps(cloud, start, end)};
  Oxts const oxts{47072.349659964, 49.011212804408, 8.4228850417969, 112.83492279053, 0.022447, 1e-05, -1.2219096732051, 0, 0, 0};
  Affine3d const T_start{OxtsToPose(oxts)};
  Twist xi;
  xi << 1.34, 0.03, -0.01, -0.003, 0.004, 0.05;
  Affine3d const T_end{T_start * lie::Exp(xi)};
  constexpr int kCallers = 6;
  std::vector<Frame> frames;
  std::vector<Time> requested;
  for (int c = 0; c < kCallers; ++c) {
    Pointcloud mine = cloud;
    if (c == 3) mine(1234, 3) = 2.0;  // w != 1 in one caller's cloud: R p + t w for that point
    frames.emplace_back(T_start, T_end, LidarScan{start, middle, end, mine, intensities, stamps});
    requested.push_back(start + (end - start) * static_cast<double>(c) / (kCallers - 1));
  }
  std::vector<Pointcloud> alone;
  for (int c = 0; c < kCallers; ++c) alone.push_back(MotionCompensateFrame(frames[static_cast<size_t>(c)], requested[static_cast<size_t>(c)]));
  std::vector<int> mismatches(kCallers, 0);
  std::vector<std::thread> callers;
  for (int c = 0; c < kCallers; ++c) {
    callers.emplace_back([&, c] {
      for (int rep = 0; rep < 10; ++rep) {
        Pointcloud const out{MotionCompensateFrame(frames[static_cast<size_t>(c)], requested[static_cast<size_t>(c)])};
        Pointcloud const& want = alone[static_cast<size_t>(c)];
        for (Index i = 0; i < out.rows(); ++i)
          for (int k = 0; k < 4; ++k)
            if (out(i, k) != want(i, k)) ++mismatches[static_cast<size_t>(c)];
      }
    });
  }
  for (auto& t : callers) t.join();
  for (int c = 0; c < kCallers; ++c) ASSERT_EQ(mismatches[static_cast<size_t>(c)], 0);
  ASSERT_TRUE(alone[3](1234, 3) == 2.0);
  ASSERT_TRUE(alone[0](0, 0) != alone[5](0, 0));  // different requested times do give different clouds
}

// kmc::MotionCompensateFrames (addition): the loop of handlers.cpp:55-64 as one pipeline; bit-identical to the per-frame calls.
TEST(RealScanTest, FramesBatchEqualsPerFrameCalls) {
  KittiPclLoader loader;
  auto [cloud, intensities] = loader.LoadPointcloud(g_real_scan);
  Time const start{47072.283701593}, middle{47072.335337762}, end{47072.386973931};
  VectorXd const stamps{GetPseudoTimeStamps(cloud, start, end)};
  Oxts const oxts{47072.349659964, 49.011212804408, 8.4228850417969, 112.83492279053, 0.022447, 1e-05, -1.2219096732051, 0, 0, 0};
  Affine3d const T_start{OxtsToPose(oxts)};
  std::vector<Frame> frames;
  std::vector<Time> requested;
  for (int k = 0; k < 5; ++k) {
    Twist xi;
    xi << 1.0 + 0.2 * k, 0.03, -0.01, -0.003, 0.004, 0.05 - 0.02 * k;
    Pointcloud mine = MatrixX4d(k == 2 ? 0 : cloud.rows() - 1000 * k, 4);  // different sizes, one empty frame
    VectorXd my_stamps(mine.rows()), my_intensities(mine.rows());
    for (Index i = 0; i < mine.rows(); ++i) {
      mine.row(i) = Vector4d{cloud(i, 0), cloud(i, 1), cloud(i, 2), 1.0};
      my_stamps(i) = stamps(i);
      my_intensities(i) = intensities(i);
    }
    frames.emplace_back(T_start, T_start * lie::Exp(xi), LidarScan{start, middle, end, mine, my_intensities, my_stamps});
    requested.push_back(k % 2 ? start : middle);
  }
  std::vector<const Frame*> pointers;
  for (auto const& f : frames) pointers.push_back(&f);
  std::vector<Pointcloud> const batch{MotionCompensateFrames(pointers, requested)};
  ASSERT_EQ(batch.size(), frames.size());
  for (size_t k = 0; k < frames.size(); ++k) {
    Pointcloud const one{MotionCompensateFrame(frames[k], requested[k])};
    ASSERT_EQ(batch[k].rows(), one.rows());
    int mismatches{0};
    for (Index i = 0; i < one.rows(); ++i)
      for (int c = 0; c < 4; ++c)
        if (batch[k](i, c) != one(i, c)) ++mismatches;
    ASSERT_EQ(mismatches, 0);
  }
  ASSERT_EQ(batch[2].rows(), 0);
  ASSERT_TRUE(MotionCompensateFrames({}, {}).empty());
  EXPECT_DEATH(MotionCompensateFrames(pointers, std::vector<Time>(5, end + 1.0)), "c");
}

// test/test_data_io.cpp:115-149 — write -> read round trip of the float32 xyzi format
TEST(DataIoTest, SavePointcloud) {
  fs::path const dir{fs::temp_directory_path() / "kmc_b200_test_io"};
  fs::create_directories(dir);
  Pointcloud cloud = MatrixX4d(3, 4);
  cloud.row(0) = Vector4d{1.5, -2.25, 3.125, 1.0};
  cloud.row(1) = Vector4d{-40.0, 0.5, -1.75, 1.0};
  cloud.row(2) = Vector4d{0.0, 100.0, 2.0, 1.0};
  VectorXd intensities(3);
  intensities(0) = 0.25;
  intensities(1) = 0.5;
  intensities(2) = 0.99;
  WritePointcloud(dir, 7, cloud, intensities);
  KittiPclLoader loader;
  auto const [back, back_i] = loader.LoadPointcloud(dir / "0000000007.bin");
  ASSERT_EQ(back.rows(), 3);
  for (Index r = 0; r < 3; ++r) {
    for (int c = 0; c < 4; ++c) ASSERT_FLOAT_EQ(back.row(r)(c), cloud.row(r)(c));
    ASSERT_FLOAT_EQ(back_i(r), intensities(r));
  }
  ASSERT_THROW(loader.LoadPointcloud(dir / "missing.bin"), std::runtime_error);
  fs::remove_all(dir);
}

// ---- handlers.cpp:41-65 on a generated three-frame run ----------------------------------------------------------------------
static void WriteLines(fs::path const& file, std::vector<std::string> const& lines) {
  std::ofstream out(file);
  for (auto const& l : lines) out << l << '\n';
}

TEST(HandlersTest, MotionCompensateRunOnGeneratedFolder) {
  fs::path const run{fs::temp_directory_path() / "kmc_b200_test_run"};
  fs::remove_all(run);
  fs::create_directories(run / "velodyne_points/data");
  fs::create_directories(run / "oxts/data");
  // five frames at 10 Hz; oxts at the middle of every scan, moving 1e-5 deg of longitude per frame (1.113 m east)
  std::vector<std::string> starts, middles, ends, oxts_stamps;
  char buf[128];
  for (int i = 0; i < 5; ++i) {
    std::snprintf(buf, sizeof(buf), "2011-09-26 13:04:%012.9f", 32.0 + 0.1 * i);
    starts.push_back(buf);
    std::snprintf(buf, sizeof(buf), "2011-09-26 13:04:%012.9f", 32.05 + 0.1 * i);
    middles.push_back(buf);
    oxts_stamps.push_back(buf);
    std::snprintf(buf, sizeof(buf), "2011-09-26 13:04:%012.9f", 32.1 + 0.1 * i);
    ends.push_back(buf);
    std::snprintf(buf, sizeof(buf), "0.0 %.10f 0.0 0.0 0.0 0.0 0 0 11.13 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 0 4 10 4 4 0", 1e-5 * i);
    WriteLines(run / "oxts/data" / (IdToZeroPaddedString(i) + ".txt"), {buf});
    std::ofstream bin(run / "velodyne_points/data" / (IdToZeroPaddedString(i) + ".bin"), std::ios::binary);
    float const pts[3][4] = {{0.f, 5.f, 0.f, 0.1f * i}, {5.f, 0.f, 0.f, 0.2f}, {0.f, -5.f, 0.f, 0.3f}};
    bin.write(reinterpret_cast<const char*>(pts), sizeof(pts));
  }
  WriteLines(run / "velodyne_points/timestamps_start.txt", starts);
  WriteLines(run / "velodyne_points/timestamps.txt", middles);
  WriteLines(run / "velodyne_points/timestamps_end.txt", ends);
  WriteLines(run / "oxts/timestamps.txt", oxts_stamps);

  MotionCompensateRun(run);

  ASSERT_EQ(NumberOfFilesInDirectory(run / "velodyne_points/data_motion_compensated"), size_t{5});
  KittiPclLoader loader;
  for (int i = 1; i <= 3; ++i) {  // same geometry as the reference's golden frame: +-0.25 of 1.11319 m
    auto const [cloud, intensities] = loader.LoadPointcloud(run / "velodyne_points/data_motion_compensated" / (IdToZeroPaddedString(i) + ".bin"));
    ASSERT_EQ(cloud.rows(), 3);
    ASSERT_FLOAT_EQ(cloud.row(0)(0), -0.27829874);
    ASSERT_FLOAT_EQ(cloud.row(0)(1), 5.0);
    ASSERT_FLOAT_EQ(cloud.row(1)(0), 5.0);
    ASSERT_NEAR(cloud.row(1)(1), 0.0, 1e-6);
    ASSERT_FLOAT_EQ(cloud.row(2)(0), 0.27829874);
    ASSERT_FLOAT_EQ(intensities(0), 0.1f * i);
  }
  // first and last frame are copied through; the last one is the LAST frame's cloud (the reference writes the first)
  auto const [last, last_i] = loader.LoadPointcloud(run / "velodyne_points/data_motion_compensated/0000000004.bin");
  ASSERT_FLOAT_EQ(last.row(0)(0), 0.0);
  ASSERT_FLOAT_EQ(last_i(0), 0.4);
  fs::remove_all(run);
}

// ---- DataHandle -----------------------------------------------------------------------------------------------------------------
TEST(DataHandleTest, FloatPathMatchesFrameApi) {
  Frame const frame{MakeMotionCompensationTestFrame()};
  kmc::b200::DataHandle handle(0, 1024);
  ASSERT_TRUE(handle.capacity() >= 1024);
  auto const Ts{frame.T_start.matrix()};
  auto const Te{frame.T_end.matrix()};
  kmc_b200_frame_params const params{kmc::b200::FrameParamsFromPoses(Ts.data(), Te.data(), 0.1, 0.2, 0.15)};
  float const in[12] = {0.f, 5.f, 0.f, 0.7f, 5.f, 0.f, 0.f, 0.8f, 0.f, -5.f, 0.f, 0.9f};
  float out[12];
  handle.DeskewScan(in, out, 3, params);
  ASSERT_FLOAT_EQ(out[0], -0.27829874);
  ASSERT_FLOAT_EQ(out[8], 0.27829874);
  ASSERT_FLOAT_EQ(out[3], 0.7);
  ASSERT_FLOAT_EQ(out[11], 0.9);
  ASSERT_THROW(kmc::b200::FrameParamsFromPoses(Ts.data(), Te.data(), 0.1, 0.2, 0.5), std::runtime_error);
}

TEST(DataHandleTest, BinFilesPipelineEqualsPerFileCalls) {
  Frame const frame{MakeMotionCompensationTestFrame()};
  auto const Ts{frame.T_start.matrix()};
  auto const Te{frame.T_end.matrix()};
  fs::path const dir{fs::temp_directory_path() / "kmc_b200_test_bin_files"};
  fs::remove_all(dir);
  fs::create_directories(dir);
  std::vector<std::string> in, out_many, out_single;
  std::vector<kmc_b200_frame_params> params;
  for (int k = 0; k < 7; ++k) {
    in.push_back((dir / ("in" + std::to_string(k) + ".bin")).string());
    out_many.push_back((dir / ("many" + std::to_string(k) + ".bin")).string());
    out_single.push_back((dir / ("single" + std::to_string(k) + ".bin")).string());
    std::ofstream bin(in.back(), std::ios::binary);
    for (int i = 0; i < 100 * k; ++i) {  // file 0 is empty
      float const p[4] = {5.f * std::cos(0.05f * i), 5.f * std::sin(0.05f * i), 0.01f * k, 0.001f * i};
      bin.write(reinterpret_cast<const char*>(p), sizeof(p));
    }
    params.push_back(kmc::b200::FrameParamsFromPoses(Ts.data(), Te.data(), 0.1, 0.2, 0.1 + 0.015 * k));
  }
  kmc::b200::DataHandle handle(0, 700);  // two or three of these files per staging slot
  std::vector<std::int64_t> const points{handle.DeskewBinFiles(in, out_many, params, 2)};
  for (int k = 0; k < 7; ++k) {
    ASSERT_EQ(points[static_cast<size_t>(k)], std::int64_t{100} * k);
    ASSERT_EQ(handle.DeskewBinFile(in[static_cast<size_t>(k)], out_single[static_cast<size_t>(k)], params[static_cast<size_t>(k)]), std::int64_t{100} * k);
    std::ifstream a(out_many[static_cast<size_t>(k)], std::ios::binary), b(out_single[static_cast<size_t>(k)], std::ios::binary);
    std::string const sa((std::istreambuf_iterator<char>(a)), std::istreambuf_iterator<char>());
    std::string const sb((std::istreambuf_iterator<char>(b)), std::istreambuf_iterator<char>());
    ASSERT_EQ(sa.size(), static_cast<size_t>(1600 * k));
    ASSERT_TRUE(sa == sb);
  }
  ASSERT_THROW(handle.DeskewBinFiles({(dir / "missing.bin").string()}, {out_many[0]}, {params[0]}), std::runtime_error);
  fs::remove_all(dir);
}

// ---- camera_model.cpp:5-95 without the drawing: the draw list of the real scan on camera 02 ------------------------------
TEST(CameraModelTest, ProjectPointcloudOnCamera) {
  KittiPclLoader loader;
  auto const [cloud, intensities] = loader.LoadPointcloud(g_real_scan);
  viz::CameraCalibration cam;
  cam.P_rect << 7.215377e+02, 0.0, 6.095593e+02, 4.485728e+01, 0.0, 7.215377e+02, 1.728540e+02, 2.163791e-01, 0.0, 0.0, 1.0, 2.745884e-03;
  Eigen::Matrix3d r_rect;
  r_rect << 9.999239e-01, 9.837760e-03, -7.445048e-03, -9.869795e-03, 9.999421e-01, -4.278459e-03, 7.402527e-03, 4.351614e-03, 9.999631e-01;
  Eigen::Matrix3d R;
  R << 7.533745e-03, -9.999714e-01, -6.166020e-04, 1.480249e-02, 7.280733e-04, -9.998902e-01, 9.998621e-01, 7.523790e-03, 1.480755e-02;
  Eigen::Affine3d tf{Eigen::Affine3d::Identity()};
  tf.linear() = R;
  tf.translation() = Eigen::Vector3d{-4.069766e-03, -7.631618e-02, -2.717806e-01};

  auto const draw{viz::ProjectPointcloudOnCamera(cloud, cam, r_rect, tf, 15.0)};
  ASSERT_EQ(static_cast<Index>(draw.size()), cloud.rows());
  int kept = 0, checked = 0;
  for (Index i = 0; i < cloud.rows(); i += 13) {
    Eigen::Vector3d const p{cloud.row(i)(0), cloud.row(i)(1), cloud.row(i)(2)};
    Eigen::Vector3d const rect{r_rect * (tf * p)};  // R_rect_00 * (R|T) * X   (camera_model.cpp:75-81)
    bool const keep = !((rect(2) < 0.01) || (rect(2) > 15.0) || (rect(1) > 1.25));  // :21-23
    bool const near_edge = std::fabs(rect(2) - 0.01) < 1e-4 || std::fabs(rect(2) - 15.0) < 1e-4 || std::fabs(rect(1) - 1.25) < 1e-4;
    if (!near_edge) ASSERT_TRUE(draw[static_cast<size_t>(i)].kept() == keep);
    if (!keep || rect(2) < 0.5) continue;
    ++kept;
    double const pu = cam.P_rect(0, 0) * rect(0) + cam.P_rect(0, 1) * rect(1) + cam.P_rect(0, 2) * rect(2) + cam.P_rect(0, 3);
    double const pv = cam.P_rect(1, 0) * rect(0) + cam.P_rect(1, 1) * rect(1) + cam.P_rect(1, 2) * rect(2) + cam.P_rect(1, 3);
    double const pw = cam.P_rect(2, 0) * rect(0) + cam.P_rect(2, 1) * rect(1) + cam.P_rect(2, 2) * rect(2) + cam.P_rect(2, 3);
    if (std::fabs(pu / pw) > 2000.0) continue;  // far off the 1242 x 375 image
    ++checked;
    ASSERT_NEAR(draw[static_cast<size_t>(i)].u, pu / pw, 2e-2);
    ASSERT_NEAR(draw[static_cast<size_t>(i)].v, pv / pw, 2e-2);
    ASSERT_NEAR(draw[static_cast<size_t>(i)].depth, rect(2), 1e-5);
    ASSERT_NEAR(draw[static_cast<size_t>(i)].colour, 255.0 * rect(2) / 14.99, 1e-3);
  }
  std::printf("    %d sampled points kept by the reference's filters, %d compared on the image\n", kept, checked);
  ASSERT_TRUE(checked > 100);
}

int main(int argc, char** argv) {
  testing::InitGoogleTest(&argc, argv);
  g_real_scan = argc > 2 ? argv[2] : "tests/golden/kitti_2011_09_26_drive_0005_frame0.bin";
  return RUN_ALL_TESTS();
}
