// motion_compensation.hpp — the deskew entry points, same names and signatures as the reference
// (include/kitti_motion_compensation/motion_compensation.hpp:8-13, src/.../motion_compensation.cpp:9-28).
//
// MotionCompensateFrame maps every point, measured in the sensor frame at its own capture time, into the sensor frame
// at the requested time:  p' = T(t_requested)^-1 T(t_i) p  with T(.) the constant-twist interpolation between
// frame.T_start and frame.T_end.  Here it runs as ONE fused CUDA kernel on a B200 (libkmc_b200: float4 xyz + per-point
// trajectory fraction in, float4 out); there is no CPU fallback — without a usable device it throws std::runtime_error.
// Out-of-range times abort, as in the reference.  Coordinates are carried in float32 on the device: the result equals
// the reference's double result to < 1e-5 m for KITTI-range clouds (inputs loaded from .bin files are float32 anyway).
#pragma once

#include "kitti_motion_compensation/data_types.hpp"
#include "kitti_motion_compensation/trajectory_interpolation.hpp"

namespace kmc {

using TrajectoryInterpolator = trajectory_interpolation::TrajectoryInterpolator;

// One point (host, double): interpolator.RelativePoseBetweenTimes(t_requested, t_point) applied to xyz1.
// Scalar convenience API; MotionCompensateFrame does not call it.
Vector4d MotionCompensatePoint(const TrajectoryInterpolator& interpolator, const Time t_point, const Vector4d& xyz1,
                               const Time t_requested);

// A whole scan (CUDA).
Pointcloud MotionCompensateFrame(const Frame& frame, const Time t_requested);

// Selects the CUDA device used by the calls above (default 0).  Addition of this implementation.
void SetMotionCompensationDevice(int device_ordinal);

}  // namespace kmc
