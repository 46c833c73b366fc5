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

// The members the reference's camera model uses (camera_model.cpp:9-12, 63-81): dynamic N x 4 / N x 3 matrices,
// transposes, a fixed-by-dynamic product, the Affine * (4 x N) product, leftCols, Ones and the colwise divide.
TEST(EigenShimTest, DynamicMatricesOfTheCameraModel) {
  Eigen::MatrixX4d cloud = Eigen::MatrixX4d::Ones(3, 4);
  ASSERT_EQ(cloud.rows(), 3);
  ASSERT_EQ(cloud(2, 3), 1.0);
  Eigen::MatrixX4d src(3, 4);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) src(r, c) = 10.0 * (r + 1) + c;  // rows (10 11 12 13), (20 ..), (30 ..)
  cloud.leftCols(3) = static_cast<Eigen::MatrixX4d const&>(src).leftCols(3);
  ASSERT_EQ(cloud(1, 2), 22.0);
  ASSERT_EQ(cloud(1, 3), 1.0);            // the homogeneous column is untouched
  ASSERT_EQ(cloud.data()[1 * 3 + 2], 31.0);  // column-major: column 1, row 2

  Eigen::Matrix4Xd const t{cloud.transpose()};
  ASSERT_EQ(t.rows(), 4);
  ASSERT_EQ(t.cols(), 3);
  ASSERT_EQ(t(2, 1), 22.0);
  Eigen::MatrixX4d const back{t.transpose()};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) ASSERT_EQ(back(r, c), cloud(r, c));

  // Affine * (4 x N): top rows L p + t w, bottom row passes through
  Eigen::Affine3d T{Eigen::Affine3d::Identity()};
  Eigen::Matrix3d L;
  L << 0, -1, 0, 1, 0, 0, 0, 0, 2;
  T.linear() = L;
  T.translation() << 100.0, 200.0, 300.0;
  Eigen::MatrixX4d const moved{(T * cloud.transpose()).transpose()};
  ASSERT_EQ(moved(0, 0), -11.0 + 100.0);
  ASSERT_EQ(moved(0, 1), 10.0 + 200.0);
  ASSERT_EQ(moved(0, 2), 2.0 * 12.0 + 300.0);
  ASSERT_EQ(moved(2, 3), 1.0);
  Eigen::Vector4d const single{T * Eigen::Vector4d{10.0, 11.0, 12.0, 1.0}};
  for (int c = 0; c < 4; ++c) ASSERT_EQ(moved(0, c), single(c));

  // (3 x 4) * (4 x N), then the perspective divide of every column by the third
  Eigen::Matrix<double, 3, 4> P;
  P << 2, 0, 1, 0.5, 0, 3, 1, 0, 0, 0, 1, 0;
  Eigen::MatrixX3d pixels = (P * cloud.transpose()).transpose();
  ASSERT_EQ(pixels.rows(), 3);
  ASSERT_EQ(pixels(1, 0), 2.0 * 20.0 + 22.0 + 0.5);
  ASSERT_EQ(pixels(1, 1), 3.0 * 21.0 + 22.0);
  ASSERT_EQ(pixels(1, 2), 22.0);
  pixels = pixels.array().colwise() / pixels.col(2).array();
  ASSERT_EQ(pixels(1, 0), (2.0 * 20.0 + 22.0 + 0.5) / 22.0);
  ASSERT_EQ(pixels(1, 2), 1.0);
  ASSERT_EQ(pixels.row(2)(1), (3.0 * 31.0 + 32.0) / 32.0);
  ASSERT_EQ(static_cast<Eigen::MatrixX4d const&>(cloud).row(2)(1), 31.0);
}

int main(int argc, char** argv) {
  testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
