// data_types.hpp — vocabulary types of the deskew path, mirroring the reference's
// include/kitti_motion_compensation/data_types.hpp:9-91 (Pointcloud, Time, Twist, Affine3d, Oxts, LidarScan, Frame).
//
// Real Eigen is used when it is installed; otherwise eigen_shim.hpp provides the members the callers touch.
// The camera/image types of the reference (cv::Mat based, data_types.hpp:61-72,95-116) belong to its visualisation
// tools, which are outside the scope of this repository: Image/Images are kept as stamp-only placeholders so that
// Frame keeps the reference's constructor signature.
#pragma once

#include <filesystem>
#include <optional>

#if defined(KMC_USE_EIGEN_SHIM) || !__has_include(<Eigen/Dense>)
#include "kitti_motion_compensation/eigen_shim.hpp"
#define KMC_B200_HAS_REAL_EIGEN 0
#else
#include <Eigen/Dense>
#define KMC_B200_HAS_REAL_EIGEN 1
#endif

namespace kmc {

// N x 4 column-major doubles; the 4th column is the homogeneous 1 (NOT the intensity).
using Pointcloud = Eigen::MatrixX4d;
// seconds since midnight
using Time = double;
using Path = std::filesystem::path;
// [rho (translation part) ; phi (rotation part)]
using Twist = Eigen::Matrix<double, 6, 1>;

using Affine3d = Eigen::Affine3d;
using MatrixX4d = Eigen::MatrixX4d;
using VectorXd = Eigen::VectorXd;
using Vector4d = Eigen::Vector4d;
using Index = Eigen::Index;

// One OxTS GPS/IMU packet (fields of the KITTI raw development kit that the path uses).
struct Oxts {
  Time stamp;
  double lat;
  double lon;
  double alt;
  double roll;
  double pitch;
  double yaw;
  double vf;
  double vl;
  double vu;
};

struct LidarScan {
  Time stamp_start;
  Time stamp_middle;  // camera trigger
  Time stamp_end;

  Pointcloud cloud;
  VectorXd intensities;  // carried beside the cloud, never enters the math
  VectorXd timestamps;   // one stamp per point (pseudo stamps from the azimuth for KITTI)
};

struct Image {
  Time stamp;
};

struct Images {
  Image image_00;
  Image image_01;
  Image image_02;
  Image image_03;
};

struct Frame {
  Frame(Affine3d const &start_pose, Affine3d const &end_pose, kmc::LidarScan const &lidar_scan,
        std::optional<kmc::Images> const camera_images = std::nullopt)
      : T_start{start_pose}, T_end{end_pose}, scan{lidar_scan}, images{camera_images} {}

  Affine3d T_start;  // sensor pose at scan.stamp_start
  Affine3d T_end;    // sensor pose at scan.stamp_end
  kmc::LidarScan scan;
  std::optional<kmc::Images> images;
};

}  // namespace kmc
