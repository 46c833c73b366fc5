// Semantics of include/kitti_motion_compensation/eigen_shim.hpp that the drop-in relies on when real Eigen is absent:
// column-major storage, Eigen's row-major comma initialiser, row proxies of MatrixX4d, Affine-mode inverse/product/
// rotation(), AngleAxisd.  Compiled with KMC_USE_EIGEN_SHIM so it also exercises the shim on boxes that have Eigen.
#define KMC_USE_EIGEN_SHIM 1
#include "kitti_motion_compensation/data_types.hpp"
#include "mini_gtest.hpp"

using namespace kmc;

TEST(EigenShimTest, FixedMatricesAreColumnMajorAndCommaInitIsRowMajor) {
  Eigen::Matrix3d m;
  m << 1, 2, 3, 4, 5, 6, 7, 8, 9;
  ASSERT_EQ(m(0, 1), 2.0);
  ASSERT_EQ(m(1, 0), 4.0);
  ASSERT_EQ(m.data()[1], 4.0);  // column-major: element (1,0) is second in memory
  ASSERT_EQ(m.data()[3], 2.0);
  ASSERT_EQ(m.trace(), 15.0);
  ASSERT_EQ(m.sum(), 45.0);
  ASSERT_EQ(m.transpose()(0, 1), 4.0);
  Eigen::Matrix<double, 3, 4> p;
  p << 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12;
  ASSERT_EQ(p(1, 3), 8.0);
  ASSERT_EQ(p.data()[3 * 3 + 1], 8.0);
  Twist xi;
  xi << 0.1, 0.2, 0.3, 0.4, 0.5, 0.6;
  ASSERT_EQ(xi.topRows(3)(2), 0.3);
  ASSERT_EQ(xi.bottomRows(3)(0), 0.4);
  Twist const cxi{xi};
  Eigen::Vector3d const rho{cxi.topRows(3)}, phi{cxi.bottomRows(3)};
  ASSERT_EQ(rho(1), 0.2);
  ASSERT_EQ(phi(2), 0.6);
  xi.topRows(3) = phi;
  ASSERT_EQ(xi(0), 0.4);
}

TEST(EigenShimTest, MatrixX4dIsColumnMajorWithRowProxies) {
  Pointcloud cloud = MatrixX4d(3, 4);
  cloud.row(0) = Vector4d{1.0, 2.0, 3.0, 1.0};
  cloud.row(2) = Vector4d{7.0, 8.0, 9.0, 1.0};
  ASSERT_EQ(cloud.rows(), 3);
  ASSERT_EQ(cloud.cols(), 4);
  ASSERT_EQ(cloud.data()[0], 1.0);      // x column first
  ASSERT_EQ(cloud.data()[2], 7.0);
  ASSERT_EQ(cloud.data()[3 + 0], 2.0);  // then the y column
  ASSERT_EQ(cloud(2, 2), 9.0);
  Vector4d const p{cloud.row(2)};
  ASSERT_EQ(p(1), 8.0);
  cloud.row(1) = cloud.row(2);
  ASSERT_EQ(cloud(1, 0), 7.0);
  Pointcloud const& view{cloud};
  ASSERT_EQ(view.row(1)(3), 1.0);
  VectorXd v(4);
  v(3) = 2.5;
  ASSERT_EQ(v.size(), 4);
  ASSERT_EQ(v.data()[3], 2.5);
}

TEST(EigenShimTest, AffineFollowsEigenAffineMode) {
  Eigen::Affine3d a{Eigen::Affine3d::Identity()};
  a.rotate(Eigen::AngleAxisd{0.5, Eigen::Vector3d::UnitZ()});
  a.translation() = Eigen::Vector3d{1.0, 2.0, 3.0};
  ASSERT_NEAR(a.linear()(0, 0), std::cos(0.5), 1e-15);
  ASSERT_NEAR(a.linear()(0, 1), -std::sin(0.5), 1e-15);
  Eigen::Affine3d const id{a * a.inverse()};
  ASSERT_NEAR(id.matrix().trace(), 4.0, 1e-14);
  ASSERT_NEAR(id.matrix().sum() - id.matrix().trace(), 0.0, 1e-14);
  // homogeneous apply: top rows L v + t w, w passes through
  Vector4d const q{a * Vector4d{1.0, 0.0, 0.0, 2.0}};
  ASSERT_NEAR(q(0), std::cos(0.5) + 2.0 * 1.0, 1e-15);
  ASSERT_NEAR(q(1), std::sin(0.5) + 2.0 * 2.0, 1e-15);
  ASSERT_EQ(q(3), 2.0);
  // matrix() is column-major 4x4 with last row 0 0 0 1
  auto const m{a.matrix()};
  ASSERT_EQ(m.data()[12], 1.0);
  ASSERT_EQ(m.data()[13], 2.0);
  ASSERT_EQ(m.data()[3], 0.0);
  ASSERT_EQ(m.data()[15], 1.0);
  // general (non-orthonormal) linear part: inverse() is the general inverse, rotation() the polar factor
  Eigen::Affine3d s{a};
  s.linear() = a.linear() * 2.0;
  Eigen::Affine3d const sid{s.inverse() * s};
  ASSERT_NEAR(sid.linear()(1, 1), 1.0, 1e-14);
  ASSERT_NEAR(sid.translation().norm(), 0.0, 1e-14);
  Eigen::Matrix3d const r{s.rotation()};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) ASSERT_NEAR(r(i, j), a.linear()(i, j), 1e-14);
  ASSERT_NEAR(r.determinant(), 1.0, 1e-14);
  // pose = R * pose (data_io.cpp:84 in the reference)
  Eigen::Matrix3d const rz{Eigen::AngleAxisd(0.3, Eigen::Vector3d::UnitZ()) * Eigen::AngleAxisd(0.2, Eigen::Vector3d::UnitY())};
  Eigen::Affine3d pose{Eigen::Affine3d::Identity()};
  pose = rz * pose;
  pose.translation() << 4.0, 5.0, 6.0;
  ASSERT_NEAR(pose.linear()(2, 0), -std::sin(0.2), 1e-15);
  ASSERT_EQ(pose.translation().y(), 5.0);
}

int main(int argc, char** argv) {
  testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
