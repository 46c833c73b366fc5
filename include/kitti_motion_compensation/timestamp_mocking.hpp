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

double FractionOfScanCompleted(Eigen::Vector4d const point);

Time GetPseudoTimeStamp(Eigen::Vector4d const point, Time const scan_start, Time const scan_end);

VectorXd GetPseudoTimeStamps(Pointcloud const &cloud, Time const start_time, Time const end_time);

}  // namespace kmc
