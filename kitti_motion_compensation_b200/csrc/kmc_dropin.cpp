// kmc_dropin.cpp — C++ mirror of the reference's public API for the deskew path, implemented on the C ABI of
// libkmc_b200 (kmc_b200.h).  Built as libkitti_motion_compensation_lib.so, the reference's library name
// (reference CMakeLists.txt:27-29), so an application linking the reference can link this instead.
//
// What runs where:
//   * MotionCompensateFrame, GetPseudoTimeStamps, MotionCompensateRun  -> CUDA kernels (no CPU fallback; a missing
//     device is a std::runtime_error)
//   * lie::*, TrajectoryInterpolator, OxtsToPose, single-point helpers -> host doubles: once-per-frame scalars
//   * file loaders / writers                                           -> host I/O
// Reference behaviours that are kept: out-of-range interpolation times abort the process
// (trajectory_interpolation.cpp:9,32 keeps assert in release builds); loader failures throw std::runtime_error
// (data_io.cpp:51,104,109); an unreadable time-stamp file prints and exits (data_io.cpp:27-30).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "kitti_motion_compensation/camera_model.hpp"
#include "kitti_motion_compensation/data_handle.hpp"
#include "kitti_motion_compensation/data_io.hpp"
#include "kitti_motion_compensation/handlers.hpp"
#include "kitti_motion_compensation/lie_algebra.hpp"
#include "kitti_motion_compensation/motion_compensation.hpp"
#include "kitti_motion_compensation/timestamp_mocking.hpp"
#include "kitti_motion_compensation/trajectory_interpolation.hpp"
#include "kitti_motion_compensation/utils.hpp"
#include "kmc_b200.h"

#define KMC_EXPORT __attribute__((visibility("default")))

namespace {

std::atomic<int> g_device{0};

[[noreturn]] void AbortOutOfRange(const char* where, double t, double t1, double t2) {
  std::fprintf(stderr,
               "%s: time %.9f is outside of the two poses being interpolated between [%.9f, %.9f] "
               "(the reference asserts here: trajectory_interpolation.cpp:32)\n",
               where, t, t1, t2);
  std::abort();
}

void ThrowUnlessOk(int status, const char* what) { kmc::b200::ThrowOnError(status, what); }

// Eigen::Matrix3d / Vector3d / Affine3d <-> the column-major double buffers of the C ABI
void ToBuffer(Eigen::Matrix3d const& m, double out[9]) {
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) out[c * 3 + r] = m(r, c);
}
Eigen::Matrix3d Matrix3FromBuffer(const double in[9]) {
  Eigen::Matrix3d m;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) m(r, c) = in[c * 3 + r];
  return m;
}
void ToBuffer(Eigen::Affine3d const& T, double out[16]) {
  for (int c = 0; c < 3; ++c) {
    for (int r = 0; r < 3; ++r) out[c * 4 + r] = T.linear()(r, c);
    out[c * 4 + 3] = 0.0;
  }
  for (int r = 0; r < 3; ++r) out[12 + r] = T.translation()(r);
  out[15] = 1.0;
}
Eigen::Affine3d AffineFromBuffer(const double in[16]) {
  Eigen::Affine3d T{Eigen::Affine3d::Identity()};
  Eigen::Matrix3d L;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) L(r, c) = in[c * 4 + r];
  T.linear() = L;
  T.translation() = Eigen::Vector3d{in[12], in[13], in[14]};
  return T;
}

kmc_b200_handle* DefaultHandle() {
  kmc_b200_handle* h = nullptr;
  ThrowUnlessOk(kmc_b200_default_handle(g_device.load(), &h), "kmc_b200_default_handle");
  return h;
}

// The reference's MotionCompensateFrame is re-entrant (it touches only its arguments).  A handle serialises its callers,
// so concurrent callers each lease their own: the process-wide default handle first, further handles (reference loader
// capacity, data_io.hpp:17) created on demand and kept for reuse.
class HandleLease {
 public:
  HandleLease() : device_{g_device.load()} {
    {
      std::lock_guard<std::mutex> lock(Mu());
      auto& idle = Idle();
      for (size_t i = 0; i < idle.size(); ++i) {
        if (idle[i].first == device_) {
          handle_ = idle[i].second;
          idle.erase(idle.begin() + static_cast<std::ptrdiff_t>(i));
          return;
        }
      }
      if (!DefaultTaken()[device_ & 63]) {
        DefaultTaken()[device_ & 63] = true;
        is_default_ = true;
      }
    }
    if (is_default_) {
      int const rc = kmc_b200_default_handle(device_, &handle_);
      if (rc != KMC_B200_OK) {
        std::lock_guard<std::mutex> lock(Mu());
        DefaultTaken()[device_ & 63] = false;
      }
      ThrowUnlessOk(rc, "kmc_b200_default_handle");
    } else {
      ThrowUnlessOk(kmc_b200_handle_create(device_, 250000, &handle_), "kmc_b200_handle_create");
    }
  }
  ~HandleLease() {
    std::lock_guard<std::mutex> lock(Mu());
    if (is_default_) DefaultTaken()[device_ & 63] = false;
    else Idle().emplace_back(device_, handle_);
  }
  HandleLease(HandleLease const&) = delete;
  HandleLease& operator=(HandleLease const&) = delete;
  kmc_b200_handle* get() const { return handle_; }

 private:
  static std::mutex& Mu() {
    static std::mutex mu;
    return mu;
  }
  static std::vector<std::pair<int, kmc_b200_handle*>>& Idle() {
    static auto* idle = new std::vector<std::pair<int, kmc_b200_handle*>>;  // never destroyed: outlives static teardown
    return *idle;
  }
  static bool* DefaultTaken() {
    static bool taken[64] = {};
    return taken;
  }
  int device_;
  kmc_b200_handle* handle_ = nullptr;
  bool is_default_ = false;
};

}  // namespace

// ===================================================================================================================
namespace kmc::lie {

KMC_EXPORT Eigen::Matrix3d Hat(Eigen::Vector3d const& a) {
  double in[3] = {a(0), a(1), a(2)}, out[9];
  kmc_b200_so3_hat(in, out);
  return Matrix3FromBuffer(out);
}

KMC_EXPORT Eigen::Vector3d Vee(Eigen::Matrix3d const& a) {
  double in[9], out[3];
  ToBuffer(a, in);
  kmc_b200_so3_vee(in, out);
  return Eigen::Vector3d{out[0], out[1], out[2]};
}

KMC_EXPORT Eigen::Matrix3d Exp(Eigen::Vector3d const& phi) {
  double in[3] = {phi(0), phi(1), phi(2)}, out[9];
  kmc_b200_so3_exp(in, out);
  return Matrix3FromBuffer(out);
}

KMC_EXPORT Eigen::Vector3d Log(Eigen::Matrix3d const& R) {
  double in[9], out[3];
  ToBuffer(R, in);
  kmc_b200_so3_log(in, out);
  return Eigen::Vector3d{out[0], out[1], out[2]};
}

KMC_EXPORT Eigen::Matrix3d LeftJacobian(Eigen::Vector3d const& phi) {
  double in[3] = {phi(0), phi(1), phi(2)}, out[9];
  kmc_b200_so3_left_jacobian(in, out);
  return Matrix3FromBuffer(out);
}

KMC_EXPORT Eigen::Matrix3d InverseLeftJacobian(Eigen::Vector3d const& phi) {
  double in[3] = {phi(0), phi(1), phi(2)}, out[9];
  kmc_b200_so3_inverse_left_jacobian(in, out);
  return Matrix3FromBuffer(out);
}

KMC_EXPORT Eigen::Affine3d Exp(Twist const& xi) {
  double in[6], out[16];
  for (int i = 0; i < 6; ++i) in[i] = xi(i);
  kmc_b200_se3_exp(in, out);
  return AffineFromBuffer(out);
}

KMC_EXPORT Twist Log(Eigen::Affine3d const& T) {
  double in[16], out[6];
  ToBuffer(T, in);
  if (kmc_b200_se3_log(in, out) != KMC_B200_OK) {
    // a linear block without a proper-rotation polar factor: Eigen would return garbage; surface NaNs instead
    for (double& v : out) v = std::nan("");
  }
  Twist xi;
  for (int i = 0; i < 6; ++i) xi(i) = out[i];
  return xi;
}

}  // namespace kmc::lie

// ===================================================================================================================
namespace kmc::trajectory_interpolation {

KMC_EXPORT Affine3d InterpolateTrajectory(Oxts const& odometry_1, Oxts const& odometry_2, Time const time) {
  return TrajectoryInterpolator(odometry_1, odometry_2).GetPoseAtTime(time);
}

KMC_EXPORT TrajectoryInterpolator::TrajectoryInterpolator(Oxts const& odometry_1, Oxts const& odometry_2)
    : time_1_{odometry_1.stamp}, pose_1_{OxtsToPose(odometry_1)}, time_2_{odometry_2.stamp}, pose_2_{OxtsToPose(odometry_2)} {}

KMC_EXPORT TrajectoryInterpolator::TrajectoryInterpolator(Time const time_1, Affine3d const& pose_1, Time const time_2,
                                                          Affine3d const& pose_2)
    : time_1_{time_1}, pose_1_{pose_1}, time_2_{time_2}, pose_2_{pose_2} {}

KMC_EXPORT Affine3d TrajectoryInterpolator::GetPoseAtTime(Time const time) const {
  double p1[16], p2[16], out[16];
  ToBuffer(pose_1_, p1);
  ToBuffer(pose_2_, p2);
  int const rc = kmc_b200_pose_at_time(time_1_, p1, time_2_, p2, time, out);
  if (rc == KMC_B200_ERR_TIME_OUT_OF_RANGE || rc == KMC_B200_ERR_EMPTY_INTERVAL)
    AbortOutOfRange("TrajectoryInterpolator::GetPoseAtTime", time, time_1_, time_2_);
  ThrowUnlessOk(rc, "kmc_b200_pose_at_time");
  return AffineFromBuffer(out);
}

KMC_EXPORT Affine3d TrajectoryInterpolator::RelativePoseBetweenTimes(Time const anchor_time, Time const query_time) const {
  return GetPoseAtTime(anchor_time).inverse() * GetPoseAtTime(query_time);
}

}  // namespace kmc::trajectory_interpolation

// ===================================================================================================================
namespace kmc {

KMC_EXPORT void SetMotionCompensationDevice(int device) { g_device.store(device); }

KMC_EXPORT double FractionOfScanCompleted(Eigen::Vector4d const point) {
  return kmc_b200_fraction_of_scan_completed(point(0), point(1));
}

KMC_EXPORT Time GetPseudoTimeStamp(Eigen::Vector4d const point, Time const scan_start, Time const scan_end) {
  return kmc_b200_pseudo_time_stamp(point(0), point(1), scan_start, scan_end);
}

// CUDA: the cloud is column-major, so x and y are its first two contiguous columns.
KMC_EXPORT VectorXd GetPseudoTimeStamps(Pointcloud const& cloud, Time const start_time, Time const end_time) {
  Index const n{cloud.rows()};
  VectorXd stamps(n);
  if (n == 0) return stamps;
  ThrowUnlessOk(kmc_b200_pseudo_time_stamps_xy_host(DefaultHandle(), cloud.data(), cloud.data() + n, n, start_time, end_time,
                                                    stamps.data()),
                "kmc_b200_pseudo_time_stamps_xy_host");
  return stamps;
}

KMC_EXPORT Vector4d MotionCompensatePoint(TrajectoryInterpolator const& trajectory_interpolator, Time const point_stamp,
                                          Vector4d const& point, Time const requested_time) {
  Affine3d const correction{trajectory_interpolator.RelativePoseBetweenTimes(requested_time, point_stamp)};
  return correction * point;
}

// CUDA, on the reference's own layout: the column-major double cloud and the per-point stamp vector go to the device as
// they are (kmc_b200_deskew_cloud_f64_host), the kernel forms each point's trajectory fraction from ITS stamp — whatever
// the caller stored in scan.timestamps is honoured, not only azimuth-derived stamps — and adds the fp32 displacement to
// the double coordinate.  No host-side layout conversion, no float32 rounding of the result.
KMC_EXPORT Pointcloud MotionCompensateFrame(Frame const& frame, Time const requested_time) {
  Time const t1{frame.scan.stamp_start}, t2{frame.scan.stamp_end};
  Index const n{frame.scan.cloud.rows()};

  double p1[16], p2[16];
  ToBuffer(frame.T_start, p1);
  ToBuffer(frame.T_end, p2);
  kmc_b200_frame_params params{};
  int rc = kmc_b200_frame_params_from_poses(p1, p2, t1, t2, requested_time, &params);
  if (rc == KMC_B200_ERR_TIME_OUT_OF_RANGE || rc == KMC_B200_ERR_EMPTY_INTERVAL) {
    if (n == 0 && rc == KMC_B200_ERR_TIME_OUT_OF_RANGE) return Pointcloud{MatrixX4d(0, 4)};  // the reference's loop never runs
    AbortOutOfRange("MotionCompensateFrame", requested_time, t1, t2);
  }
  ThrowUnlessOk(rc, "kmc_b200_frame_params_from_poses");

  if (frame.scan.timestamps.size() != n) throw std::invalid_argument("MotionCompensateFrame: scan.timestamps and scan.cloud differ in length");
  Pointcloud result{MatrixX4d(n, 4)};
  if (n == 0) return result;

  int flags{0};
  HandleLease const lease;
  rc = kmc_b200_deskew_cloud_f64_host(lease.get(), frame.scan.cloud.data(), frame.scan.timestamps.data(), result.data(), n, t1, t2,
                                      requested_time, &params, &flags);
  if (flags & 1) AbortOutOfRange("MotionCompensateFrame (a point stamp)", std::nan(""), t1, t2);  // GetPoseAtTime(point_stamp) asserts
  ThrowUnlessOk(rc, "kmc_b200_deskew_cloud_f64_host");
  return result;
}

KMC_EXPORT std::vector<Pointcloud> MotionCompensateFrames(std::vector<const Frame*> const& frames, std::vector<Time> const& requested_times) {
  if (frames.size() != requested_times.size()) throw std::invalid_argument("MotionCompensateFrames: frames and requested_times differ in length");
  size_t const F{frames.size()};
  std::vector<Pointcloud> results;
  results.reserve(F);
  std::vector<kmc_b200_frame_params> params(F);
  std::vector<const double*> clouds(F), stamps(F);
  std::vector<double*> outs(F);
  std::vector<std::int64_t> n_points(F);
  std::vector<double> times(3 * F);
  for (size_t k{0}; k < F; ++k) {
    if (!frames[k]) throw std::invalid_argument("MotionCompensateFrames: null frame");
    Frame const& frame{*frames[k]};
    Time const t1{frame.scan.stamp_start}, t2{frame.scan.stamp_end};
    Index const n{frame.scan.cloud.rows()};
    if (frame.scan.timestamps.size() != n) throw std::invalid_argument("MotionCompensateFrames: scan.timestamps and scan.cloud differ in length");
    double p1[16], p2[16];
    ToBuffer(frame.T_start, p1);
    ToBuffer(frame.T_end, p2);
    int const rc{kmc_b200_frame_params_from_poses(p1, p2, t1, t2, requested_times[k], &params[k])};
    if (rc == KMC_B200_ERR_TIME_OUT_OF_RANGE || rc == KMC_B200_ERR_EMPTY_INTERVAL) {
      if (n == 0 && rc == KMC_B200_ERR_TIME_OUT_OF_RANGE) {  // the reference's loop never runs for an empty cloud
        params[k] = kmc_b200_frame_params{};
        times[3 * k] = 0.0, times[3 * k + 1] = 1.0, times[3 * k + 2] = 0.5;
      } else {
        AbortOutOfRange("MotionCompensateFrames", requested_times[k], t1, t2);
      }
    } else {
      ThrowUnlessOk(rc, "kmc_b200_frame_params_from_poses");
      times[3 * k] = t1, times[3 * k + 1] = t2, times[3 * k + 2] = requested_times[k];
    }
    results.emplace_back(MatrixX4d(n, 4));
    clouds[k] = frame.scan.cloud.data();
    stamps[k] = frame.scan.timestamps.data();
    outs[k] = results.back().data();
    n_points[k] = n;
  }
  if (F == 0) return results;
  std::vector<int> flags(F, 0);
  HandleLease const lease;
  int const rc{kmc_b200_deskew_cloud_f64_batch_host(lease.get(), clouds.data(), stamps.data(), outs.data(), n_points.data(), times.data(),
                                                    params.data(), static_cast<std::int32_t>(F), flags.data())};
  for (size_t k{0}; k < F; ++k)
    if (flags[k] & 1) AbortOutOfRange("MotionCompensateFrames (a point stamp)", std::nan(""), times[3 * k], times[3 * k + 1]);
  ThrowUnlessOk(rc, "kmc_b200_deskew_cloud_f64_batch_host");
  return results;
}

// ---- utils ------------------------------------------------------------------------------------------------------------
KMC_EXPORT std::string IdToZeroPaddedString(size_t const id, size_t const pad) {
  std::string digits{std::to_string(id)};
  if (digits.size() < pad) digits.insert(digits.begin(), pad - digits.size(), '0');
  return digits;
}

KMC_EXPORT std::vector<std::string> TokenizeString(std::string raw_string) {
  std::vector<std::string> tokens;
  std::istringstream stream(raw_string);
  for (std::string tok; std::getline(stream, tok, ' ');) tokens.push_back(tok);
  return tokens;
}

KMC_EXPORT double MmHhSsToSeconds(std::string const mm_hh_ss) {
  // "HH:MM:SS.fffffffff"
  int const hours{std::stoi(mm_hh_ss.substr(0, 2))};
  int const minutes{std::stoi(mm_hh_ss.substr(3, 2))};
  double const seconds{std::stod(mm_hh_ss.substr(6))};
  return static_cast<double>(3600 * hours + 60 * minutes) + seconds;
}

// ---- data_io ----------------------------------------------------------------------------------------------------------
KMC_EXPORT Time LoadTimeStamp(Path const timestamp_file, size_t const frame_id) {
  std::ifstream in(timestamp_file);
  if (!in.is_open()) {
    std::cout << "Failed to open timestamp file: " << timestamp_file << '\n';
    std::exit(0);  // the reference's convention (data_io.cpp:27-30)
  }
  std::string line;
  for (size_t i{0}; i <= frame_id; ++i) std::getline(in, line);
  std::vector<std::string> const tokens{TokenizeString(line)};  // "2011-09-26 13:04:32.283701593"
  if (tokens.size() < 2) throw std::runtime_error("Time stamp file has no line for frame " + std::to_string(frame_id) + ": " + timestamp_file.string());
  return Time(MmHhSsToSeconds(tokens[1]));
}

KMC_EXPORT Oxts LoadOxts(Path const folder, size_t const frame_id) {
  Time const stamp{LoadTimeStamp(folder / Path("oxts/timestamps.txt"), frame_id)};
  Path const file(folder / Path("oxts/data/" + IdToZeroPaddedString(frame_id) + ".txt"));
  std::ifstream in(file);
  if (!in.is_open()) throw std::runtime_error("The Oxts file you tried to load did not open: " + file.string());
  std::string line;
  std::getline(in, line);
  std::vector<std::string> const v{TokenizeString(line)};
  if (v.size() < 11) throw std::runtime_error("The Oxts file is malformed: " + file.string());
  // lat lon alt roll pitch yaw vn ve vf vl vu ...
  return Oxts{stamp,           std::stod(v[0]), std::stod(v[1]), std::stod(v[2]), std::stod(v[3]),
              std::stod(v[4]), std::stod(v[5]), std::stod(v[8]), std::stod(v[9]), std::stod(v[10])};
}

KMC_EXPORT Eigen::Affine3d OxtsToPose(Oxts const& odometry, double const scale) {
  double out[16];
  kmc_b200_oxts_to_pose(odometry.lat, odometry.lon, odometry.alt, odometry.roll, odometry.pitch, odometry.yaw, scale, out);
  return AffineFromBuffer(out);
}

KMC_EXPORT KittiPclLoader::KittiPclLoader() : data_{new float[pcl_buffer_size]} {}
KMC_EXPORT KittiPclLoader::~KittiPclLoader() { delete[] data_; }

KMC_EXPORT std::tuple<Pointcloud, VectorXd> KittiPclLoader::LoadPointcloud(Path const& file) {
  std::ifstream in{file, std::ios::in | std::ios::binary | std::ios::ate};
  if (!in.is_open()) throw std::runtime_error("Unable to open requested KITTI pointcloud binary file: " + file.string());
  std::int64_t const bytes{static_cast<std::int64_t>(in.tellg())};
  // data_io.cpp:107-112: any multiple of 4 bytes is accepted and a trailing partial point is ignored.  The reference would
  // overrun its 250 000-point buffer on a larger file; that is an error here.
  if (bytes < 0 || bytes % 4 != 0 || static_cast<size_t>(bytes) > pcl_buffer_size * sizeof(float))
    throw std::runtime_error("Opened KITTI pointcloud binary file is incorrectly formatted: " + file.string());
  in.seekg(0, std::ios::beg);
  in.read(reinterpret_cast<char*>(data_), bytes);
  Index const n{static_cast<Index>(bytes / 16)};
  Pointcloud cloud{MatrixX4d(n, 4)};
  VectorXd intensities(n);
  double* c = cloud.data();
  for (Index i = 0; i < n; ++i) {
    c[i] = data_[4 * i];
    c[n + i] = data_[4 * i + 1];
    c[2 * n + i] = data_[4 * i + 2];
    c[3 * n + i] = 1.0;
    intensities(i) = data_[4 * i + 3];
  }
  return {cloud, intensities};
}

KMC_EXPORT LidarScan LoadLidarScan(Path const folder, size_t const frame_id) {
  Time const start{LoadTimeStamp(folder / Path("velodyne_points/timestamps_start.txt"), frame_id)};
  Time const middle{LoadTimeStamp(folder / Path("velodyne_points/timestamps.txt"), frame_id)};
  Time const end{LoadTimeStamp(folder / Path("velodyne_points/timestamps_end.txt"), frame_id)};
  KittiPclLoader loader;
  auto [cloud, intensities] = loader.LoadPointcloud(folder / Path("velodyne_points/data/" + IdToZeroPaddedString(frame_id) + ".bin"));
  VectorXd stamps{GetPseudoTimeStamps(cloud, start, end)};
  return LidarScan{start, middle, end, cloud, intensities, stamps};
}

KMC_EXPORT Frame MakeFrame(kmc::Oxts const& odometry_n_m_1, kmc::Oxts const& odometry_n, kmc::Oxts const& odometry_n_p_1,
                           kmc::LidarScan const& lidar_scan, std::optional<kmc::Images> const camera_images) {
  Affine3d const start_pose{trajectory_interpolation::InterpolateTrajectory(odometry_n_m_1, odometry_n, lidar_scan.stamp_start)};
  Affine3d const end_pose{trajectory_interpolation::InterpolateTrajectory(odometry_n, odometry_n_p_1, lidar_scan.stamp_end)};
  return Frame(start_pose, end_pose, lidar_scan, camera_images);
}

KMC_EXPORT Frame LoadSingleFrame(Path const data_folder, size_t const frame_id, bool const load_images) {
  if (load_images) throw std::invalid_argument("LoadSingleFrame: image loading is outside the scope of this implementation");
  Oxts const prev{LoadOxts(data_folder, frame_id - 1)}, cur{LoadOxts(data_folder, frame_id)}, next{LoadOxts(data_folder, frame_id + 1)};
  return MakeFrame(prev, cur, next, LoadLidarScan(data_folder, frame_id));
}

KMC_EXPORT void WritePointcloud(Path const data_folder, size_t const frame_id, Pointcloud const& pointcloud,
                                VectorXd const& intensities) {
  Index const n{pointcloud.rows()};
  std::vector<float> buf(static_cast<size_t>(4 * n));
  const double* c = pointcloud.data();
  for (Index i = 0; i < n; ++i) {
    buf[static_cast<size_t>(4 * i)] = static_cast<float>(c[i]);
    buf[static_cast<size_t>(4 * i + 1)] = static_cast<float>(c[n + i]);
    buf[static_cast<size_t>(4 * i + 2)] = static_cast<float>(c[2 * n + i]);
    buf[static_cast<size_t>(4 * i + 3)] = static_cast<float>(intensities(i));
  }
  std::ofstream out(data_folder / Path(IdToZeroPaddedString(frame_id) + ".bin"), std::ios::out | std::ios::binary);
  out.write(reinterpret_cast<const char*>(buf.data()), static_cast<std::streamsize>(buf.size() * sizeof(float)));
}

// ---- handlers -------------------------------------------------------------------------------------------------------------
KMC_EXPORT std::size_t NumberOfFilesInDirectory(std::filesystem::path path) {
  return static_cast<std::size_t>(std::distance(std::filesystem::directory_iterator{path}, std::filesystem::directory_iterator{}));
}

namespace {
void CopyFile(Path const& from, Path const& to) {
  std::filesystem::copy_file(from, to, std::filesystem::copy_options::overwrite_existing);
}
}  // namespace

KMC_EXPORT void CopyOverUncompensatedFirstAndLastFrame(Path const run_folder) {
  Path const velodyne{run_folder / Path{"velodyne_points"}};
  size_t const n{NumberOfFilesInDirectory(velodyne / Path("data"))};
  if (n == 0) return;
  for (size_t id : {size_t{0}, n - 1}) {
    // float32 xyzi in, float32 xyzi out: a byte copy is exactly the reference's load + WritePointcloud round trip
    CopyFile(velodyne / Path("data/" + IdToZeroPaddedString(id) + ".bin"),
             velodyne / Path("data_motion_compensated/" + IdToZeroPaddedString(id) + ".bin"));
  }
}

// The reference handles one frame at a time (load three oxts packets, load + expand the scan, deskew, write).  Here the
// whole run goes through kmc_b200_motion_compensate_run: text files read once, scans streamed disk -> pinned memory ->
// GPU -> pinned memory -> disk in groups of ~16 with reads, copies, kernels and writes overlapped.  The handle for runs
// is separate from the per-frame default handle because its staging slots hold a group of scans, not one.
KMC_EXPORT void MotionCompensateRun(Path const run_folder) {
  constexpr std::int64_t kRunSlotPoints = 2'000'000;  // 32 MB per staging buffer, ~16 KITTI scans per group
  static std::mutex mu;
  static kmc_b200_handle* run_handle = nullptr;
  static int run_handle_device = -1;
  std::lock_guard<std::mutex> lock(mu);
  if (!run_handle || run_handle_device != g_device.load()) {
    if (run_handle) kmc_b200_handle_destroy(run_handle);
    run_handle = nullptr;
    ThrowUnlessOk(kmc_b200_handle_create(g_device.load(), kRunSlotPoints, &run_handle), "kmc_b200_handle_create");
    run_handle_device = g_device.load();
  }
  kmc_b200_run_stats stats{};
  // one line per frame as the run advances (handlers.cpp:63); file k of the pipeline is frame k + 1
  kmc_b200_handle_set_file_callback(
      run_handle, [](std::int32_t file_index, std::int64_t, void*) { std::cout << "Motion compensated pointcloud number: " << file_index + 1 << std::endl; },
      nullptr);
  int const rc = kmc_b200_motion_compensate_run(run_handle, run_folder.c_str(), 0, &stats);
  kmc_b200_handle_set_file_callback(run_handle, nullptr, nullptr);
  if (rc == KMC_B200_ERR_TIME_OUT_OF_RANGE || rc == KMC_B200_ERR_EMPTY_INTERVAL) {
    std::fprintf(stderr, "MotionCompensateRun: %s (the reference asserts here: trajectory_interpolation.cpp:32)\n", kmc_b200_last_error());
    std::abort();
  }
  if (rc == KMC_B200_ERR_IO && std::string(kmc_b200_last_error()).rfind("failed to open timestamp file", 0) == 0) {
    std::cout << kmc_b200_last_error() << '\n';
    std::exit(0);  // the reference's convention for an unreadable time-stamp file (data_io.cpp:27-30)
  }
  ThrowUnlessOk(rc, "kmc_b200_motion_compensate_run");
}

// ---- calibration files + projection (camera_model.cpp, data_io.cpp:168-210,321-406) -----------------------------------------
namespace {
std::vector<double> NumbersAfterLabel(std::string const& line, size_t expected, char const* what) {
  std::vector<std::string> const tokens{TokenizeString(line)};  // "R_rect_00: 9.99e-01 ..."
  if (tokens.size() < expected + 1) throw std::runtime_error(std::string("Calibration line is too short for ") + what + ": " + line);
  std::vector<double> values;
  for (size_t i{1}; i <= expected; ++i) values.push_back(std::stod(tokens[i]));
  return values;
}
template <class M>
void FillRowMajor(M& m, std::vector<double> const& v) {
  for (Index r = 0; r < m.rows(); ++r)
    for (Index c = 0; c < m.cols(); ++c) m(r, c) = v[static_cast<size_t>(r * m.cols() + c)];
}
}  // namespace

KMC_EXPORT Eigen::Affine3d LoadLidarExtrinsics(kmc::Path const data_folder, bool const to_cam) {
  Path const file{data_folder / Path(to_cam ? "calib_velo_to_cam.txt" : "calib_imu_to_velo.txt")};
  std::ifstream in(file);
  if (!in.is_open()) {
    std::cout << "Failed to open camera calibration file: " << file << '\n';
    std::exit(0);  // the reference's convention (data_io.cpp:178-181)
  }
  std::string line;
  std::getline(in, line);  // calib_time
  std::getline(in, line);
  Eigen::Matrix3d R;
  FillRowMajor(R, NumbersAfterLabel(line, 9, "R"));
  std::getline(in, line);
  std::vector<double> const t{NumbersAfterLabel(line, 3, "T")};
  Eigen::Affine3d T{Eigen::Affine3d::Identity()};
  T.linear() = R;
  T.translation() = Eigen::Vector3d{t[0], t[1], t[2]};
  return T;
}

}  // namespace kmc

namespace kmc::viz {

KMC_EXPORT CameraCalibration CalibrationLinesToCalibration(std::vector<std::string> const lines) {
  if (lines.size() < 8) throw std::runtime_error("A camera calibration block has eight lines (S K D R T S_rect R_rect P_rect)");
  CameraCalibration c;
  std::vector<double> v{NumbersAfterLabel(lines[0], 2, "S")};
  c.S = Eigen::Vector2d{v[0], v[1]};
  FillRowMajor(c.K, NumbersAfterLabel(lines[1], 9, "K"));
  FillRowMajor(c.D, NumbersAfterLabel(lines[2], 5, "D"));
  FillRowMajor(c.R, NumbersAfterLabel(lines[3], 9, "R"));
  v = NumbersAfterLabel(lines[4], 3, "T");
  c.T = Eigen::Vector3d{v[0], v[1], v[2]};
  v = NumbersAfterLabel(lines[5], 2, "S_rect");
  c.S_rect = Eigen::Vector2d{v[0], v[1]};
  FillRowMajor(c.R_rect, NumbersAfterLabel(lines[6], 9, "R_rect"));
  FillRowMajor(c.P_rect, NumbersAfterLabel(lines[7], 12, "P_rect"));
  return c;
}

KMC_EXPORT CameraCalibrations LoadCameraCalibrations(kmc::Path const data_folder) {
  Path const file{data_folder / Path("calib_cam_to_cam.txt")};
  std::ifstream in(file);
  if (!in.is_open()) {
    std::cout << "Failed to open camera calibration file: " << file << '\n';
    std::exit(0);  // data_io.cpp:379-382
  }
  std::string line;
  std::getline(in, line);  // calib_time
  std::getline(in, line);  // corner_dist
  std::vector<CameraCalibration> cameras;
  for (int cam = 0; cam < 4; ++cam) {
    std::vector<std::string> block;
    for (int i = 0; i < 8; ++i) {
      std::getline(in, line);
      block.push_back(line);
    }
    cameras.push_back(CalibrationLinesToCalibration(block));
  }
  return CameraCalibrations{cameras[0], cameras[1], cameras[2], cameras[3]};
}

KMC_EXPORT std::vector<ProjectedPoint> ProjectPointcloudOnCamera(Pointcloud const& cloud, CameraCalibration const& camera,
                                                                 Eigen::Matrix3d const& r_rect_00, Eigen::Affine3d const& tf_c00_lo,
                                                                 double const max_range) {
  double p_rect[12], r_rect[9], tf[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 3; ++r) p_rect[c * 3 + r] = camera.P_rect(r, c);
  ToBuffer(r_rect_00, r_rect);
  ToBuffer(tf_c00_lo, tf);
  kmc_b200_camera_params params{};
  ThrowUnlessOk(kmc_b200_camera_params_from_calibration(p_rect, r_rect, tf, max_range, &params), "kmc_b200_camera_params_from_calibration");
  Index const n{cloud.rows()};
  std::vector<ProjectedPoint> result(static_cast<size_t>(n));
  if (n == 0) return result;
  std::vector<float> xyzi(static_cast<size_t>(4 * n));
  const double* c = cloud.data();
  for (Index i = 0; i < n; ++i) {
    xyzi[static_cast<size_t>(4 * i)] = static_cast<float>(c[i]);
    xyzi[static_cast<size_t>(4 * i + 1)] = static_cast<float>(c[n + i]);
    xyzi[static_cast<size_t>(4 * i + 2)] = static_cast<float>(c[2 * n + i]);
    xyzi[static_cast<size_t>(4 * i + 3)] = 0.0f;
  }
  static_assert(sizeof(ProjectedPoint) == 4 * sizeof(float), "ProjectedPoint mirrors the kernel's float4 record");
  ThrowUnlessOk(kmc_b200_project_frame_host(DefaultHandle(), xyzi.data(), reinterpret_cast<float*>(result.data()), n, &params),
                "kmc_b200_project_frame_host");
  return result;
}

}  // namespace kmc::viz
