// lie_algebra.hpp — so(3)/se(3) helpers with the reference's names and conventions
// (reference: include/kitti_motion_compensation/lie_algebra.hpp:12-26, src/.../lie_algebra.cpp:7-103).
// Twist ordering is [rho; phi]; every closed form switches to its first-order Taylor expansion below 1e-6 rad.
// Implemented in double on the host by libkmc_b200 (kmc_b200_so3_* / kmc_b200_se3_* in kmc_b200.h): these are the
// once-per-frame scalars of the deskew path, not per-point work.
#pragma once

#include "kitti_motion_compensation/data_types.hpp"

namespace kmc::lie {

using Vec3 = Eigen::Vector3d;
using Mat3 = Eigen::Matrix3d;

// skew-symmetric matrix of a 3-vector, and its inverse map
Mat3 Hat(const Vec3& omega);
Vec3 Vee(const Mat3& skew);

// SO(3): Rodrigues exponential and its logarithm
Mat3 Exp(const Vec3& rotation_vector);
Vec3 Log(const Mat3& rotation);

// left Jacobian of SO(3) and its closed-form inverse
Mat3 LeftJacobian(const Vec3& rotation_vector);
Mat3 InverseLeftJacobian(const Vec3& rotation_vector);

// SE(3): rotation Exp(phi), translation J(phi) rho  /  the inverse map.  Log takes the polar projection of the linear
// block, exactly what Eigen's Affine-mode rotation() returns.
Eigen::Affine3d Exp(const Twist& twist);
Twist Log(const Eigen::Affine3d& transform);

}  // namespace kmc::lie
