// utilities_for_testing.hpp — comparison helpers with the semantics of the reference's
// include/kitti_motion_compensation/utilities_for_testing.hpp:4-11 (float-cast trace / off-trace sum of A * B^-1).
#pragma once

#include <cmath>

#include "kitti_motion_compensation/data_types.hpp"

namespace kmc::utilities_for_testing {

inline bool FloatEqual(float const a, float const b, float const epsilon = 1e-10) { return std::fabs(a - b) <= epsilon; }

// true when tf1 * tf2^-1 is the identity to float precision
inline bool TransformationMatricesAreTheSame(Affine3d const& tf1, Affine3d const& tf2) {
  auto const product{tf1 * tf2.inverse()};
  auto const m{product.matrix()};
  return FloatEqual(static_cast<float>(m.trace()), 4.0f) and FloatEqual(static_cast<float>(m.sum() - m.trace()), 0.0f);
}

}  // namespace kmc::utilities_for_testing
