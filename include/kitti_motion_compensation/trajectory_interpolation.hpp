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

Affine3d InterpolateTrajectory(Oxts const &odometry_1, Oxts const &odometry_2, Time const time);

class TrajectoryInterpolator {
 public:
  // poses from OxtsToPose(odometry_k), times from odometry_k.stamp
  TrajectoryInterpolator(Oxts const &odometry_1, Oxts const &odometry_2);

  TrajectoryInterpolator(Time const time_1, Affine3d const &pose_1, Time const time_2, Affine3d const &pose_2);

  Affine3d GetPoseAtTime(Time const time) const;

  Affine3d RelativePoseBetweenTimes(Time const anchor_time, Time const query_time) const;

  // additions of this implementation (read-only accessors used by the batched deskew entry points)
  Time time_1() const { return time_1_; }
  Time time_2() const { return time_2_; }
  Affine3d const &pose_1() const { return pose_1_; }
  Affine3d const &pose_2() const { return pose_2_; }

 private:
  Time time_1_;
  Affine3d pose_1_;
  Time time_2_;
  Affine3d pose_2_;
};

}  // namespace kmc::trajectory_interpolation
