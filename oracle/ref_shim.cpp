// ref_shim.cpp — C entry points over the UNMODIFIED reference sources, compiled only by `make -C oracle ref` on a box
// that has <Eigen/Dense> (this image does not, so this file has never been compiled here — see DESIGN.md §5).
// It exposes the same two calls as libkmc_oracle.so so that tests/bench can swap the restatement for the real thing:
//   kmc_ref_deskew_xyzi_scan   = KittiPclLoader conversion + kmc::GetPseudoTimeStamps + kmc::MotionCompensateFrame
//   kmc_ref_timed_frames       = the same over many scans on std::threads, timed
// TEST INFRASTRUCTURE ONLY; outputs go to oracle/_ref/.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <thread>
#include <vector>

#include "kitti_motion_compensation/data_types.hpp"
#include "kitti_motion_compensation/motion_compensation.hpp"
#include "kitti_motion_compensation/timestamp_mocking.hpp"

#include <stdexcept>

#include "kitti_motion_compensation/data_io.hpp"

// trajectory_interpolation.cpp:21-25 (the Oxts constructor) references kmc::OxtsToPose, which lives in data_io.cpp
// together with the OpenCV image loaders and is therefore not compiled into oracle/_ref.  The deskew path never takes
// that constructor; this definition only satisfies the dynamic linker.
namespace kmc {
Eigen::Affine3d OxtsToPose(Oxts const&, double const) { throw std::logic_error("OxtsToPose is not part of oracle/_ref"); }
}  // namespace kmc

namespace {

kmc::Affine3d FromColMajor(const double* m) {
  kmc::Affine3d T{kmc::Affine3d::Identity()};
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 3; ++r) T.matrix()(r, c) = m[c * 4 + r];
  return T;
}

kmc::Pointcloud Deskew(const float* xyzi, int64_t n, const double* T_start, const double* T_end, double t0, double t2, double t_req) {
  kmc::Pointcloud cloud{kmc::MatrixX4d(n, 4)};
  kmc::VectorXd intensities{kmc::VectorXd(n)};
  for (int64_t i = 0; i < n; ++i) {  // data_io.cpp:126-135
    cloud(i, 0) = xyzi[4 * i];
    cloud(i, 1) = xyzi[4 * i + 1];
    cloud(i, 2) = xyzi[4 * i + 2];
    cloud(i, 3) = 1.0;
    intensities(i) = xyzi[4 * i + 3];
  }
  kmc::VectorXd const stamps{kmc::GetPseudoTimeStamps(cloud, t0, t2)};
  kmc::LidarScan const scan{t0, t_req, t2, cloud, intensities, stamps};
  kmc::Frame const frame(FromColMajor(T_start), FromColMajor(T_end), scan);
  return kmc::MotionCompensateFrame(frame, t_req);
}

}  // namespace

extern "C" {

int kmc_ref_deskew_xyzi_scan(const float* xyzi, int64_t n, const double* T_start, const double* T_end, double t0, double t2,
                             double t_req, double* out_xyz1) {
  kmc::Pointcloud const res{Deskew(xyzi, n, T_start, T_end, t0, t2, t_req)};
  for (int64_t i = 0; i < n; ++i)
    for (int c = 0; c < 4; ++c) out_xyz1[4 * i + c] = res(i, c);
  return 0;
}

double kmc_ref_timed_frames(const float* xyzi, int64_t points_per_frame, int32_t n_frames, const double* T_start,
                            const double* T_end, const double* stamps3, int32_t n_threads, double* checksum) {
  std::atomic<int32_t> next{0};
  std::vector<double> partial(static_cast<size_t>(n_threads < 1 ? 1 : n_threads), 0.0);
  auto worker = [&](int tid) {
    double acc = 0.0;
    for (;;) {
      int32_t const f = next.fetch_add(1);
      if (f >= n_frames) break;
      kmc::Pointcloud const res{Deskew(xyzi + static_cast<int64_t>(f) * points_per_frame * 4, points_per_frame, T_start + 16 * f,
                                       T_end + 16 * f, stamps3[3 * f], stamps3[3 * f + 1], stamps3[3 * f + 2])};
      acc += res.sum();
    }
    partial[static_cast<size_t>(tid)] = acc;
  };
  auto const a = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (size_t t = 0; t < partial.size(); ++t) pool.emplace_back(worker, static_cast<int>(t));
  for (auto& th : pool) th.join();
  auto const b = std::chrono::steady_clock::now();
  double total = 0.0;
  for (double p : partial) total += p;
  if (checksum) *checksum = total;
  return std::chrono::duration<double>(b - a).count();
}

}  // extern "C"
