// timestamp_mocking.hpp — azimuth -> pseudo time stamp (reference: include/.../timestamp_mocking.hpp:7-11,
// src/.../timestamp_mocking.cpp:46-63).  KITTI scans carry no per-point time; the spinning sensor starts at the back
// of the vehicle, so  fraction = (pi - atan2(y, x)) / 2pi  and  stamp = start + fraction * (end - start).
//
// The fused deskew kernel computes the fraction itself (KMC_B200_TIME_FROM_AZIMUTH) and never materialises stamps;
// GetPseudoTimeStamps exists for callers that want the stamps and runs as a double-precision CUDA kernel over the
// cloud's x and y columns.
#pragma once

#include "kitti_motion_compensation/data_types.hpp"

namespace kmc {

// in [0, 1]; 0.5 is straight ahead (the camera trigger direction)
double FractionOfScanCompleted(const Eigen::Vector4d xyz1);

// one point (host)
Time GetPseudoTimeStamp(const Eigen::Vector4d xyz1, const Time t_scan_begin, const Time t_scan_finish);

// every row of the cloud (CUDA)
VectorXd GetPseudoTimeStamps(const Pointcloud& xyz1_rows, const Time t_scan_begin, const Time t_scan_finish);

}  // namespace kmc
