// ref_shim.cpp — C entry points over the UNMODIFIED reference sources, compiled only by `make -C oracle ref`:
// against Eigen 3 where <Eigen/Dense> exists, otherwise (this image) against ref_stub/Eigen/Dense, which forwards to
// this repository's eigen_shim.hpp — see DESIGN.md §5 for what that does and does not prove.
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

#include <cstring>
#include <stdexcept>
#include <string>

#include "kitti_motion_compensation/camera_model.hpp"
#include "kitti_motion_compensation/data_io.hpp"
#include "kitti_motion_compensation/handlers.hpp"
#include "kitti_motion_compensation/lie_algebra.hpp"
#include "kitti_motion_compensation/trajectory_interpolation.hpp"

namespace {

kmc::Affine3d FromColMajor(const double* m) {
  kmc::Affine3d T{kmc::Affine3d::Identity()};
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T.linear()(r, c) = m[c * 4 + r];
    T.translation()(r) = m[12 + r];
  }
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

// ---- the primitives of the path, one entry point per reference function (same shapes as the kmc_oracle_* calls) ----
static void Mat3Out(Eigen::Matrix3d const& m, double* out) {
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) out[c * 3 + r] = m(r, c);
}
static Eigen::Matrix3d Mat3In(const double* in) {
  Eigen::Matrix3d m;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) m(r, c) = in[c * 3 + r];
  return m;
}
static void AffineOut(kmc::Affine3d const& T, double* out) {
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) out[c * 4 + r] = T.matrix()(r, c);
}

const char* kmc_ref_eigen_provider(void) {
#ifdef EIGEN_WORLD_VERSION
  return "eigen3";
#else
  return "shim";
#endif
}

void kmc_ref_hat(const double phi[3], double out[9]) { Mat3Out(kmc::lie::Hat(Eigen::Vector3d{phi[0], phi[1], phi[2]}), out); }
void kmc_ref_vee(const double m[9], double out[3]) {
  Eigen::Vector3d const v{kmc::lie::Vee(Mat3In(m))};
  for (int i = 0; i < 3; ++i) out[i] = v(i);
}
void kmc_ref_so3_exp(const double phi[3], double out[9]) { Mat3Out(kmc::lie::Exp(Eigen::Vector3d{phi[0], phi[1], phi[2]}), out); }
void kmc_ref_so3_log(const double R[9], double out[3]) {
  Eigen::Vector3d const v{kmc::lie::Log(Mat3In(R))};
  for (int i = 0; i < 3; ++i) out[i] = v(i);
}
void kmc_ref_left_jacobian(const double phi[3], double out[9]) {
  Mat3Out(kmc::lie::LeftJacobian(Eigen::Vector3d{phi[0], phi[1], phi[2]}), out);
}
void kmc_ref_inverse_left_jacobian(const double phi[3], double out[9]) {
  Mat3Out(kmc::lie::InverseLeftJacobian(Eigen::Vector3d{phi[0], phi[1], phi[2]}), out);
}
void kmc_ref_se3_exp(const double xi[6], double out[16]) {
  kmc::Twist t;
  for (int i = 0; i < 6; ++i) t(i) = xi[i];
  AffineOut(kmc::lie::Exp(t), out);
}
void kmc_ref_se3_log(const double T[16], double xi[6]) {
  kmc::Twist const t{kmc::lie::Log(FromColMajor(T))};
  for (int i = 0; i < 6; ++i) xi[i] = t(i);
}
// The reference aborts on a time outside [t1, t2] (trajectory_interpolation.cpp:32); callers check the range first.
void kmc_ref_pose_at_time(double t1, const double P1[16], double t2, const double P2[16], double t, double out[16]) {
  kmc::TrajectoryInterpolator const interp(t1, FromColMajor(P1), t2, FromColMajor(P2));
  AffineOut(interp.GetPoseAtTime(t), out);
}
void kmc_ref_relative_pose_between_times(double t1, const double P1[16], double t2, const double P2[16], double anchor,
                                         double query, double out[16]) {
  kmc::TrajectoryInterpolator const interp(t1, FromColMajor(P1), t2, FromColMajor(P2));
  AffineOut(interp.RelativePoseBetweenTimes(anchor, query), out);
}
double kmc_ref_fraction_of_scan_completed(const double p[4]) {
  return kmc::FractionOfScanCompleted(kmc::Vector4d{p[0], p[1], p[2], p[3]});
}
double kmc_ref_pseudo_time_stamp(const double p[4], double start, double end) {
  return kmc::GetPseudoTimeStamp(kmc::Vector4d{p[0], p[1], p[2], p[3]}, start, end);
}
void kmc_ref_motion_compensate_point(double t1, const double P1[16], double t2, const double P2[16], double point_stamp,
                                     const double p[4], double requested_time, double out[4]) {
  kmc::TrajectoryInterpolator const interp(t1, FromColMajor(P1), t2, FromColMajor(P2));
  kmc::Vector4d const r{kmc::MotionCompensatePoint(interp, point_stamp, kmc::Vector4d{p[0], p[1], p[2], p[3]}, requested_time)};
  for (int i = 0; i < 4; ++i) out[i] = r(i);
}
// MotionCompensateFrame on the reference's own layout: column-major n x 4 cloud + n stamps -> column-major n x 4.
void kmc_ref_motion_compensate_frame(const double* cloud_colmajor, const double* stamps, int64_t n, const double T_start[16],
                                     const double T_end[16], double t0, double t2, double t_req, double* out_colmajor) {
  kmc::Pointcloud cloud{kmc::MatrixX4d(n, 4)};
  kmc::VectorXd ts{kmc::VectorXd(n)}, intensities{kmc::VectorXd(n)};
  for (int64_t i = 0; i < n; ++i) {
    for (int c = 0; c < 4; ++c) cloud(i, c) = cloud_colmajor[c * n + i];
    ts(i) = stamps[i];
    intensities(i) = 0.0;
  }
  kmc::LidarScan const scan{t0, t_req, t2, cloud, intensities, ts};
  kmc::Frame const frame(FromColMajor(T_start), FromColMajor(T_end), scan);
  kmc::Pointcloud const res{kmc::MotionCompensateFrame(frame, t_req)};
  for (int64_t i = 0; i < n; ++i)
    for (int c = 0; c < 4; ++c) out_colmajor[c * n + i] = res(i, c);
}

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
      for (kmc::Index i = 0; i < res.rows(); ++i)
        for (int c = 0; c < 4; ++c) acc += res(i, c);
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


// ---- data_io.cpp / handlers.cpp: the rows either side of the path --------------------------------
static kmc::Oxts OxtsFrom7(const double o[7]) { return kmc::Oxts{o[0], o[1], o[2], o[3], o[4], o[5], o[6], 0.0, 0.0, 0.0}; }

void kmc_ref_oxts_to_pose(const double oxts7[7], double scale, double out[16]) { AffineOut(kmc::OxtsToPose(OxtsFrom7(oxts7), scale), out); }

// MakeFrame (data_io.cpp:253-269): start / end pose of a scan from the three surrounding OxTS packets.  Aborts (the
// reference's assert) when a scan stamp lies outside its packet interval — callers check first.
void kmc_ref_make_frame_poses(const double o_prev[7], const double o_cur[7], const double o_next[7], double stamp_start,
                              double stamp_end, double T_start[16], double T_end[16]) {
  kmc::LidarScan scan{};
  scan.stamp_start = stamp_start;
  scan.stamp_middle = 0.5 * (stamp_start + stamp_end);
  scan.stamp_end = stamp_end;
  kmc::Frame const frame{kmc::MakeFrame(OxtsFrom7(o_prev), OxtsFrom7(o_cur), OxtsFrom7(o_next), scan)};
  AffineOut(frame.T_start, T_start);
  AffineOut(frame.T_end, T_end);
}

void kmc_ref_interpolate_trajectory(const double o1[7], const double o2[7], double t, double out[16]) {
  AffineOut(kmc::trajectory_interpolation::InterpolateTrajectory(OxtsFrom7(o1), OxtsFrom7(o2), t), out);
}

double kmc_ref_load_time_stamp(const char* file, int64_t frame_id) { return kmc::LoadTimeStamp(kmc::Path(file), static_cast<size_t>(frame_id)); }

void kmc_ref_load_oxts(const char* run_folder, int64_t frame_id, double out7[7]) {
  kmc::Oxts const o{kmc::LoadOxts(kmc::Path(run_folder), static_cast<size_t>(frame_id))};
  double const v[7] = {o.stamp, o.lat, o.lon, o.alt, o.roll, o.pitch, o.yaw};
  std::memcpy(out7, v, sizeof v);
}

// KittiPclLoader::LoadPointcloud (data_io.cpp:101-138): returns the number of points; fills at most `capacity` rows of
// out_xyz1 (row-major n x 4 doubles) and out_intensity.  Pass capacity 0 to only count.
int64_t kmc_ref_load_pointcloud(const char* file, double* out_xyz1, double* out_intensity, int64_t capacity) {
  kmc::KittiPclLoader loader;
  auto const [cloud, intensities] = loader.LoadPointcloud(kmc::Path(file));
  int64_t const n = cloud.rows();
  for (int64_t i = 0; i < n && i < capacity; ++i) {
    for (int c = 0; c < 4; ++c) out_xyz1[4 * i + c] = cloud(i, c);
    out_intensity[i] = intensities(i);
  }
  return n;
}

// WritePointcloud (data_io.cpp:287-313) of a row-major n x 4 double cloud + intensities.
void kmc_ref_write_pointcloud(const char* folder, int64_t frame_id, const double* xyz1, const double* intensity, int64_t n) {
  kmc::Pointcloud cloud{kmc::MatrixX4d(n, 4)};
  kmc::VectorXd in{kmc::VectorXd(n)};
  for (int64_t i = 0; i < n; ++i) {
    for (int c = 0; c < 4; ++c) cloud(i, c) = xyz1[4 * i + c];
    in(i) = intensity[i];
  }
  kmc::WritePointcloud(kmc::Path(folder), static_cast<size_t>(frame_id), cloud, in);
}

// handlers.cpp:41-65 on a KITTI run folder; returns the wall time in seconds.
double kmc_ref_motion_compensate_run(const char* run_folder) {
  auto const a = std::chrono::steady_clock::now();
  kmc::MotionCompensateRun(kmc::Path(run_folder));
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
}

// ---- camera_model.cpp:38-95 + :5-36 ---------------------------------------------------------------
// ProjectPointcloudOnFrame on a column-major n x 4 cloud.  cv::circle is the recording stub of ref_stub/opencv2, so what
// comes back is the reference's own draw list per camera: integer pixel centres and the three colour channels, in call
// order.  uv_out[k] holds 2 ints and color_out[k] 3 doubles per drawn point (capacity n each); counts[k] = points drawn.
void kmc_ref_project_pointcloud_on_frame(const double* cloud_colmajor, int64_t n, const double tf_c00_lo[16],
                                         const double R_rect_00[9], const double P_rect[4][12], int32_t* const uv_out[4],
                                         double* const color_out[4], int64_t counts[4]) {
  kmc::LidarScan scan{};
  scan.cloud = kmc::MatrixX4d(n, 4);
  for (int64_t i = 0; i < n; ++i)
    for (int c = 0; c < 4; ++c) scan.cloud(i, c) = cloud_colmajor[c * n + i];
  scan.intensities = kmc::VectorXd(n);
  scan.timestamps = kmc::VectorXd(n);
  kmc::Image const blank{0.0, cv::Mat(375, 1242)};
  kmc::Frame const frame(kmc::Affine3d::Identity(), kmc::Affine3d::Identity(), scan, kmc::Images{blank, blank, blank, blank});
  kmc::viz::CameraCalibrations calib{};
  kmc::viz::CameraCalibration* cams[4] = {&calib.camera_00, &calib.camera_01, &calib.camera_02, &calib.camera_03};
  for (int k = 0; k < 4; ++k) {
    cams[k]->R_rect = Eigen::Matrix3d::Identity();
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) cams[k]->P_rect(r, c) = P_rect[k][c * 3 + r];  // column-major 3 x 4
  }
  calib.camera_00.R_rect = Mat3In(R_rect_00);
  kmc::Images const drawn{kmc::viz::ProjectPointcloudOnFrame(frame, calib, FromColMajor(tf_c00_lo))};
  kmc::Image const* imgs[4] = {&drawn.image_00, &drawn.image_01, &drawn.image_02, &drawn.image_03};
  for (int k = 0; k < 4; ++k) {
    auto const& list = imgs[k]->image.drawn;
    counts[k] = static_cast<int64_t>(list.size());
    for (size_t j = 0; j < list.size(); ++j) {
      uv_out[k][2 * j] = list[j].center.x;
      uv_out[k][2 * j + 1] = list[j].center.y;
      for (int c = 0; c < 3; ++c) color_out[k][3 * j + c] = list[j].color.val[c];
    }
  }
}


// LoadLidarExtrinsics + LoadCameraCalibrations (data_io.cpp:168-210, 321-406) on a KITTI calibration folder: the
// velodyne -> camera_00 transform, R_rect_00, the four P_rect (3x4, column-major each) and S_rect_00.
void kmc_ref_load_calibration(const char* folder, double T_velo_to_cam[16], double R_rect_00[9], double P_rect[4][12], double S_rect_00[2]) {
  AffineOut(kmc::LoadLidarExtrinsics(kmc::Path(folder), true), T_velo_to_cam);
  kmc::viz::CameraCalibrations const c{kmc::viz::LoadCameraCalibrations(kmc::Path(folder))};
  kmc::viz::CameraCalibration const* cams[4] = {&c.camera_00, &c.camera_01, &c.camera_02, &c.camera_03};
  Mat3Out(c.camera_00.R_rect, R_rect_00);
  for (int k = 0; k < 4; ++k)
    for (int col = 0; col < 4; ++col)
      for (int r = 0; r < 3; ++r) P_rect[k][col * 3 + r] = cams[k]->P_rect(r, col);
  S_rect_00[0] = c.camera_00.S_rect(0);
  S_rect_00[1] = c.camera_00.S_rect(1);
}

}  // extern "C"
