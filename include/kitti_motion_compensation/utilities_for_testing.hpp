// utilities_for_testing.hpp — pose comparison helper with the semantics of the reference's test utility
// (reference include/kitti_motion_compensation/utilities_for_testing.hpp:4-11): two transforms count as equal when
// A * B^-1 has, after rounding to float, a trace of exactly 4 and off-diagonal entries summing to exactly 0
// (tolerance 1e-10, i.e. effectively float equality).
#pragma once

#include <cmath>

#include "kitti_motion_compensation/data_types.hpp"

namespace kmc::utilities_for_testing {

inline bool FloatEqual(float const a, float const b, float const epsilon = 1e-10) {
  float const gap{std::fabs(a - b)};
  return gap <= epsilon;
}

inline bool TransformationMatricesAreTheSame(Affine3d const& tf1, Affine3d const& tf2) {
  Affine3d const should_be_identity{tf1 * tf2.inverse()};
  double diagonal{1.0};  // the homogeneous corner
  double everything{1.0};
  for (int r = 0; r < 3; ++r) {
    diagonal += should_be_identity.linear()(r, r);
    everything += should_be_identity.translation()(r);
    for (int c = 0; c < 3; ++c) everything += should_be_identity.linear()(r, c);
  }
  bool const trace_is_four{FloatEqual(static_cast<float>(diagonal), 4.0f)};
  bool const rest_is_zero{FloatEqual(static_cast<float>(everything - diagonal), 0.0f)};
  return trace_is_four && rest_is_zero;
}

}  // namespace kmc::utilities_for_testing
