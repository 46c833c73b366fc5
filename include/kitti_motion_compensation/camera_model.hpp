// camera_model.hpp — projection of a scan onto the rectified KITTI cameras (reference:
// include/kitti_motion_compensation/camera_model.hpp:5-11, src/.../camera_model.cpp:5-95, calibration types
// data_types.hpp:95-116, calibration parsers data_io.cpp:168-210,321-406).
//
// The reference draws the projected points into cv::Mat images with OpenCV.  This implementation stops one step
// earlier: it returns, per point, the pixel coordinates, the depth in front of the camera and the reference's range
// colour (or "culled"), computed by the CUDA projection kernel (kmc_b200_project_frame_host) — the draw list.  Painting
// circles into an image stays with the caller.  Y = P_rect_xx * R_rect_00 * (R|T)_velo_to_cam * X, cull when
// z_rect < 0.01, z_rect > max_range or y_rect > 1.25 (camera_model.cpp:21-23).
#pragma once

#include <string>
#include <vector>

#include "kitti_motion_compensation/data_types.hpp"

namespace kmc::viz {

// pinhole projection matrix of a rectified camera
typedef Eigen::Matrix<double, 3, 4> P;

// one "xx" block of calib_cam_to_cam.txt
struct CameraCalibration {
  Eigen::Vector2d S;
  Eigen::Matrix3d K;
  Eigen::Matrix<double, 5, 1> D;
  Eigen::Matrix3d R;
  Eigen::Vector3d T;
  Eigen::Vector2d S_rect;
  Eigen::Matrix3d R_rect;
  P P_rect;
};

struct CameraCalibrations {
  CameraCalibration camera_00;
  CameraCalibration camera_01;
  CameraCalibration camera_02;
  CameraCalibration camera_03;
};

// the eight "S_xx: ... P_rect_xx: ..." lines of one camera -> CameraCalibration
CameraCalibration CalibrationLinesToCalibration(std::vector<std::string> const calibration_lines);

// <data_folder>/calib_cam_to_cam.txt
CameraCalibrations LoadCameraCalibrations(kmc::Path const data_folder);

// What the reference would draw for one point.
struct ProjectedPoint {
  float u;       // pixel column (the reference truncates to int when drawing)
  float v;       // pixel row
  float depth;   // z in the rectified camera_00 frame, metres
  float colour;  // 255 * depth / (max_range - 0.01) for kept points, negative when the reference skips the point
  bool kept() const { return colour >= 0.0f; }
};

// Per-point part of ProjectPointcloudOnFrame + ProjectPointcloudOnImage for one camera (CUDA).
// cloud: N x 4 column-major doubles (x y z 1); tf_c00_lo: velodyne -> camera_00; r_rect_00: camera_00.R_rect.
std::vector<ProjectedPoint> ProjectPointcloudOnCamera(Pointcloud const &cloud, CameraCalibration const &camera,
                                                      Eigen::Matrix3d const &r_rect_00, Eigen::Affine3d const &tf_c00_lo,
                                                      double const max_range = 15.0);

}  // namespace kmc::viz

namespace kmc {

// <data_folder>/calib_velo_to_cam.txt (to_cam) or calib_imu_to_velo.txt: the "R:" and "T:" lines as an Affine3d
Eigen::Affine3d LoadLidarExtrinsics(kmc::Path const data_folder, bool const to_cam = true);

}  // namespace kmc
