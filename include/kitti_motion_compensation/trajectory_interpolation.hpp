// trajectory_interpolation.hpp — constant-twist (screw motion) interpolation between two stamped poses
// (reference: include/kitti_motion_compensation/trajectory_interpolation.hpp:9-31,
//  src/.../trajectory_interpolation.cpp:14-51).
//
//   GetPoseAtTime(t)                  = P1 * Exp( x * Log(P1^-1 P2) ),  x = (t - t1) / (t2 - t1)
//   RelativePoseBetweenTimes(a, q)    = GetPoseAtTime(a)^-1 * GetPoseAtTime(q)
//
// Like the reference (which keeps its asserts in release builds, trajectory_interpolation.cpp:9,32), a time outside
// [t1, t2] ABORTS the process.  The C ABI underneath reports KMC_B200_ERR_TIME_OUT_OF_RANGE instead.
#pragma once

#include "kitti_motion_compensation/data_types.hpp"

namespace kmc::trajectory_interpolation {

// pose at `t` on the screw motion between two OxTS packets (poses from OxtsToPose, times from their stamps)
Affine3d InterpolateTrajectory(const Oxts& packet_a, const Oxts& packet_b, const Time t);

class TrajectoryInterpolator {
 public:
  TrajectoryInterpolator(const Oxts& packet_a, const Oxts& packet_b);
  TrajectoryInterpolator(const Time t_a, const Affine3d& pose_a, const Time t_b, const Affine3d& pose_b);

  Affine3d GetPoseAtTime(const Time t) const;
  Affine3d RelativePoseBetweenTimes(const Time t_anchor, const Time t_query) const;

  // additions of this implementation: read-only access for the batched entry points
  Time time_1() const { return time_1_; }
  Time time_2() const { return time_2_; }
  const Affine3d& pose_1() const { return pose_1_; }
  const Affine3d& pose_2() const { return pose_2_; }

 private:
  Time time_1_;
  Affine3d pose_1_;
  Time time_2_;
  Affine3d pose_2_;
};

}  // namespace kmc::trajectory_interpolation
