// lie_algebra.hpp — so(3)/se(3) helpers with the reference's names and conventions
// (reference: include/kitti_motion_compensation/lie_algebra.hpp:12-26, src/.../lie_algebra.cpp:7-103).
// Twist ordering is [rho; phi]; every closed form switches to its first-order Taylor expansion below 1e-6 rad.
// Implemented in double on the host by libkmc_b200 (kmc_b200_so3_* / kmc_b200_se3_* in kmc_b200.h): these are the
// once-per-frame scalars of the deskew path, not per-point work.
#pragma once

#include "kitti_motion_compensation/data_types.hpp"

namespace kmc::lie {

Eigen::Matrix3d Hat(Eigen::Vector3d const& a);

Eigen::Vector3d Vee(Eigen::Matrix3d const& a);

// SO(3) exponential (Rodrigues)
Eigen::Matrix3d Exp(Eigen::Vector3d const& phi);

// SO(3) logarithm
Eigen::Vector3d Log(Eigen::Matrix3d const& R);

Eigen::Matrix3d LeftJacobian(Eigen::Vector3d const& phi);

Eigen::Matrix3d InverseLeftJacobian(Eigen::Vector3d const& phi);

// SE(3) exponential: rotation Exp(phi), translation J(phi) rho
Eigen::Affine3d Exp(Twist const& xi);

// SE(3) logarithm; the rotation is the polar projection of T's linear block, as Eigen's Affine-mode rotation()
Twist Log(Eigen::Affine3d const& T);

}  // namespace kmc::lie
