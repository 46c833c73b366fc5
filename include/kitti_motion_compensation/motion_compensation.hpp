// motion_compensation.hpp — the deskew entry points, same names and signatures as the reference
// (include/kitti_motion_compensation/motion_compensation.hpp:8-13, src/.../motion_compensation.cpp:9-28).
//
// MotionCompensateFrame maps every point, measured in the sensor frame at its own capture time, into the sensor frame
// at the requested time:  p' = T(t_requested)^-1 T(t_i) p  with T(.) the constant-twist interpolation between
// frame.T_start and frame.T_end.  Here it runs as ONE fused CUDA kernel on a B200, directly on the reference's layout
// (column-major double cloud + per-point stamps in, column-major double cloud out; kmc_b200_deskew_cloud_f64_host);
// there is no CPU fallback — without a usable device it throws std::runtime_error.  Out-of-range times abort, as in the
// reference.  The per-point displacement is computed in fp32 and added to the double coordinate: the result equals
// the reference's double result to ~2e-7 m for KITTI-range clouds.
#pragma once

#include <vector>

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

// Several frames at once (addition of this implementation): what a caller that loops MotionCompensateFrame over the frames
// of a run does (handlers.cpp:55-64, 67-92), as ONE pipeline — the host passes, transfers and kernels of neighbouring frames
// overlap instead of draining at the end of every frame (kmc_b200_deskew_cloud_f64_batch_host).  result[k] is bit-identical
// to MotionCompensateFrame(*frames[k], requested_times[k]); the same abort rules apply to every frame.
std::vector<Pointcloud> MotionCompensateFrames(const std::vector<const Frame*>& frames, const std::vector<Time>& requested_times);

}  // namespace kmc
