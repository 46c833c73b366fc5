// data_io.hpp — KITTI raw-data loaders and writers on either side of the deskew path (reference:
// include/kitti_motion_compensation/data_io.hpp:9-100, src/.../data_io.cpp:18-166,253-313).  These are the rows
// SURVEY.md 8(f) marks "next": host-side file parsing and once-per-frame pose preparation, kept API-compatible so a
// run can be driven from raw OxTS + velodyne files.  Image and camera-calibration loading (OpenCV) is out of scope.
//
// Error conventions follow the reference: unreadable oxts / velodyne files throw std::runtime_error; an unreadable
// time-stamp file prints a message and exits (data_io.cpp:27-30).
#pragma once

#include <cstddef>
#include <optional>
#include <string>
#include <tuple>

#include "kitti_motion_compensation/data_types.hpp"

namespace kmc {

Time LoadTimeStamp(Path const timestamp_file, size_t const frame_id);

Oxts LoadOxts(Path const folder, size_t const frame_id);

// Mercator projection + Rz(yaw) Ry(pitch) Rx(roll), as pykitti (computed by kmc_b200_oxts_to_pose)
Eigen::Affine3d OxtsToPose(Oxts const &odometry, double const scale = 1.0);

constexpr size_t float_entries_per_point{4};                         // x y z intensity
constexpr size_t pcl_buffer_size{250000 * float_entries_per_point};  // staging capacity in floats

// Reads a velodyne .bin (n x 4 float32) into the reference's in-memory layout: column-major double cloud with a
// homogeneous ones column + a separate intensity vector.  Owns its staging buffer; non-copyable.
class KittiPclLoader {
 public:
  KittiPclLoader();
  ~KittiPclLoader();
  KittiPclLoader(KittiPclLoader const &other) = delete;
  KittiPclLoader &operator=(KittiPclLoader const &other) = delete;

  std::tuple<Pointcloud, VectorXd> LoadPointcloud(Path const &file);

 private:
  float *data_;
};

// start/middle/end stamps + cloud + pseudo time stamps (GetPseudoTimeStamps runs on the GPU)
LidarScan LoadLidarScan(Path const folder, size_t const frame_id);

// T_start from (oxts n-1, oxts n) at stamp_start, T_end from (oxts n, oxts n+1) at stamp_end
Frame MakeFrame(kmc::Oxts const &odometry_n_m_1, kmc::Oxts const &odometry_n, kmc::Oxts const &odometry_n_p_1,
                kmc::LidarScan const &lidar_scan, std::optional<kmc::Images> const camera_images = std::nullopt);

// load_images must be false: image loading is outside this repository's scope (throws std::invalid_argument)
Frame LoadSingleFrame(Path const data_folder, size_t const frame_id, bool const load_images = false);

// <data_folder>/<10-digit id>.bin, n x 4 float32: x y z from the cloud, the 4th float from `intensities`
void WritePointcloud(Path const data_folder, size_t const frame_id, Pointcloud const &pointcloud,
                     VectorXd const &intensities);

}  // namespace kmc
