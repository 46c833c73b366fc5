// eigen_shim.hpp — the small subset of Eigen 3 that the callers of the deskew path use.
//
// The reference's public API carries Eigen types (include/kitti_motion_compensation/data_types.hpp:3,14,25-31 in the
// reference repo).  When <Eigen/Dense> is installed the drop-in headers use the real thing; when it is not (this
// image), this file supplies just enough of namespace Eigen — column-major fixed-size matrices, MatrixX4d / MatrixX3d, VectorXd,
// Affine3d, AngleAxisd, comma initialisers — for the reference's call sites and test bodies to compile unchanged.
// It is a data carrier, not a linear-algebra library: the numerics of the path live behind the C ABI (kmc_b200.h).
#pragma once

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <memory>
#include <type_traits>
#include <utility>
#include <vector>

namespace Eigen {

using Index = std::ptrdiff_t;

template <typename Derived>
class CommaInitializer {
 public:
  CommaInitializer(Derived& target, double first) : target_(target), next_(0) { push(first); }
  CommaInitializer& operator,(double v) {
    push(v);
    return *this;
  }
  template <typename Other, typename = decltype(std::declval<const Other&>().size())>
  CommaInitializer& operator,(const Other& block) {
    for (Index i = 0; i < static_cast<Index>(block.size()); ++i) push(block(i));
    return *this;
  }

 private:
  void push(double v) {
    // row-major fill order, as Eigen's comma initialiser
    Index const r = next_ / target_.cols(), c = next_ % target_.cols();
    assert(r < target_.rows());
    target_(r, c) = v;
    ++next_;
  }
  Derived& target_;
  Index next_;
};

// Fixed-size, column-major R x C matrix of doubles.
template <typename Scalar, int R, int C>
class Matrix {
  static_assert(std::is_same<Scalar, double>::value, "the shim only carries doubles");

 public:
  static constexpr int RowsAtCompileTime = R;
  static constexpr int ColsAtCompileTime = C;

  Matrix() : v_{} {}

  // Vector3d{a, b, c}, Vector4d{...}: exactly R*C scalars, column vectors only
  template <typename... T, typename = std::enable_if_t<(sizeof...(T) == R * C) && (C == 1) && (sizeof...(T) > 1) &&
                                                       (std::is_convertible<T, double>::value && ...)>>
  Matrix(T... vals) : v_{static_cast<double>(vals)...} {}

  // from anything row/vector-like with matching length (MatrixX4d::row(i), blocks)
  template <typename Other, typename = std::enable_if_t<!std::is_convertible<Other, double>::value &&
                                                        !std::is_same<std::decay_t<Other>, Matrix>::value>,
            typename = decltype(std::declval<const Other&>().size()), typename = decltype(std::declval<const Other&>()(Index{0}))>
  Matrix(const Other& o) : v_{} {
    assert(static_cast<Index>(o.size()) == R * C);
    for (Index i = 0; i < R * C; ++i) v_[static_cast<size_t>(i)] = o(i);
  }

  static Matrix Zero() { return Matrix(); }
  static Matrix Identity() {
    Matrix m;
    for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1.0;
    return m;
  }
  static Matrix UnitX() { return Unit(0); }
  static Matrix UnitY() { return Unit(1); }
  static Matrix UnitZ() { return Unit(2); }

  Index rows() const { return R; }
  Index cols() const { return C; }
  Index size() const { return R * C; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }

  double& operator()(Index r, Index c) { return v_[static_cast<size_t>(c * R + r)]; }
  double operator()(Index r, Index c) const { return v_[static_cast<size_t>(c * R + r)]; }
  double& operator()(Index i) { return v_[static_cast<size_t>(i)]; }
  double operator()(Index i) const { return v_[static_cast<size_t>(i)]; }
  double& operator[](Index i) { return v_[static_cast<size_t>(i)]; }
  double operator[](Index i) const { return v_[static_cast<size_t>(i)]; }
  double& x() { return v_[0]; }
  double& y() { return v_[1]; }
  double& z() { return v_[2]; }
  double x() const { return v_[0]; }
  double y() const { return v_[1]; }
  double z() const { return v_[2]; }

  CommaInitializer<Matrix> operator<<(double first) { return CommaInitializer<Matrix>(*this, first); }
  template <typename Other, typename = decltype(std::declval<const Other&>().size())>
  Matrix& operator<<(const Other& o) {  // T.translation() << (Matrix3d * Vector3d)
    *this = Matrix(o);
    return *this;
  }

  double norm() const { return std::sqrt(squaredNorm()); }
  double squaredNorm() const {
    double s = 0;
    for (double x : v_) s += x * x;
    return s;
  }
  double sum() const {
    double s = 0;
    for (double x : v_) s += x;
    return s;
  }
  double trace() const {
    double s = 0;
    for (int i = 0; i < (R < C ? R : C); ++i) s += (*this)(i, i);
    return s;
  }
  Matrix<double, C, R> transpose() const {
    Matrix<double, C, R> t;
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < C; ++c) t(c, r) = (*this)(r, c);
    return t;
  }
  const Matrix& matrix() const { return *this; }

  double determinant() const {
    static_assert(R == 3 && C == 3, "determinant() is provided for 3x3 only");
    const Matrix& m = *this;
    return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
           m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
  }
  Matrix inverse() const {
    static_assert(R == 3 && C == 3, "inverse() is provided for 3x3 only");
    const Matrix& m = *this;
    Matrix a;
    a(0, 0) = m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1);
    a(0, 1) = m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2);
    a(0, 2) = m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1);
    a(1, 0) = m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2);
    a(1, 1) = m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0);
    a(1, 2) = m(0, 2) * m(1, 0) - m(0, 0) * m(1, 2);
    a(2, 0) = m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0);
    a(2, 1) = m(0, 1) * m(2, 0) - m(0, 0) * m(2, 1);
    a(2, 2) = m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0);
    double const inv = 1.0 / (m(0, 0) * a(0, 0) + m(1, 0) * a(0, 1) + m(2, 0) * a(0, 2));
    return a * inv;
  }

  // xi.topRows(3) / xi.bottomRows(3) on a 6-vector; read access returns a copy, write access goes through a view
  class RowsView {
   public:
    RowsView(Matrix& m, Index first, Index count) : m_(m), first_(first), count_(count) {}
    template <typename Other>
    RowsView& operator=(const Other& o) {
      for (Index i = 0; i < count_; ++i) m_(first_ + i) = o(i);
      return *this;
    }
    Index size() const { return count_; }
    double operator()(Index i) const { return m_(first_ + i); }

   private:
    Matrix& m_;
    Index first_, count_;
  };
  RowsView topRows(Index n) {
    static_assert(C == 1, "topRows() is provided for column vectors only");
    return RowsView(*this, 0, n);
  }
  RowsView bottomRows(Index n) {
    static_assert(C == 1, "bottomRows() is provided for column vectors only");
    return RowsView(*this, R - n, n);
  }
  Matrix<double, 3, 1> topRows(Index n) const {
    assert(n == 3 && C == 1);
    (void)n;
    return Matrix<double, 3, 1>{v_[0], v_[1], v_[2]};
  }
  Matrix<double, 3, 1> bottomRows(Index n) const {
    assert(n == 3 && C == 1);
    (void)n;
    return Matrix<double, 3, 1>{v_[R - 3], v_[R - 2], v_[R - 1]};
  }

  Matrix operator+(const Matrix& o) const {
    Matrix r;
    for (size_t i = 0; i < v_.size(); ++i) r.v_[i] = v_[i] + o.v_[i];
    return r;
  }
  Matrix operator-(const Matrix& o) const {
    Matrix r;
    for (size_t i = 0; i < v_.size(); ++i) r.v_[i] = v_[i] - o.v_[i];
    return r;
  }
  Matrix operator-() const { return *this * -1.0; }
  Matrix operator*(double s) const {
    Matrix r;
    for (size_t i = 0; i < v_.size(); ++i) r.v_[i] = v_[i] * s;
    return r;
  }
  Matrix operator/(double s) const { return *this * (1.0 / s); }
  Matrix& operator+=(const Matrix& o) { return *this = *this + o; }
  Matrix& operator*=(double s) { return *this = *this * s; }
  template <int K>
  Matrix<double, R, K> operator*(const Matrix<double, C, K>& o) const {
    Matrix<double, R, K> r;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < K; ++j) {
        double acc = 0;
        for (int k = 0; k < C; ++k) acc += (*this)(i, k) * o(k, j);
        r(i, j) = acc;
      }
    return r;
  }
  bool operator==(const Matrix& o) const { return v_ == o.v_; }

 private:
  static Matrix Unit(int i) {
    static_assert(C == 1, "UnitX/Y/Z are for column vectors");
    Matrix m;
    m(i) = 1.0;
    return m;
  }
  std::array<double, static_cast<size_t>(R* C)> v_;
};

template <int R, int C>
Matrix<double, R, C> operator*(double s, const Matrix<double, R, C>& m) {
  return m * s;
}

using Vector2d = Matrix<double, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Vector4d = Matrix<double, 4, 1>;
using Matrix3d = Matrix<double, 3, 3>;
using Matrix4d = Matrix<double, 4, 4>;

// Storage of the dynamic types: like Eigen, a sized constructor allocates WITHOUT initialising (a 123 397 x 4 result
// matrix is 4 MB; zero-filling it would cost as much as the deskew call that fills it).
namespace internal {
template <class T>
struct DefaultInitAllocator : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = DefaultInitAllocator<U>;
  };
  using std::allocator<T>::allocator;
  template <class U>
  void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) {
    ::new (static_cast<void*>(p)) U;
  }
  template <class U, class... Args>
  void construct(U* p, Args&&... args) {
    ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
  }
};
using Storage = std::vector<double, DefaultInitAllocator<double>>;
}  // namespace internal

// Dynamic column vector.
class VectorXd {
 public:
  VectorXd() = default;
  explicit VectorXd(Index n) : v_(static_cast<size_t>(n)) {}
  Index size() const { return static_cast<Index>(v_.size()); }
  Index rows() const { return size(); }
  Index cols() const { return 1; }
  double& operator()(Index i) { return v_[static_cast<size_t>(i)]; }
  double operator()(Index i) const { return v_[static_cast<size_t>(i)]; }
  double& operator[](Index i) { return v_[static_cast<size_t>(i)]; }
  double operator[](Index i) const { return v_[static_cast<size_t>(i)]; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  void resize(Index n) { v_.resize(static_cast<size_t>(n)); }

 private:
  internal::Storage v_;
};

template <int R>
class DynColsMatrix;

// Dynamic-rows x C, COLUMN-major.  C = 4 is kmc::Pointcloud (x column, y column, z column, homogeneous column); C = 3 is
// the pixel matrix of the reference's camera model (camera_model.cpp:9).
template <int C>
class DynRowsMatrix {
 public:
  class RowRef {
   public:
    RowRef(double* base, Index stride) : base_(base), stride_(stride) {}
    RowRef& operator=(const Matrix<double, C, 1>& v) {
      for (int c = 0; c < C; ++c) base_[c * stride_] = v(c);
      return *this;
    }
    RowRef& operator=(const RowRef& o) {
      for (int c = 0; c < C; ++c) base_[c * stride_] = o(c);
      return *this;
    }
    double& operator()(Index c) { return base_[c * stride_]; }
    double operator()(Index c) const { return base_[c * stride_]; }
    Index size() const { return C; }

   private:
    double* base_;
    Index stride_;
  };
  class ConstRowRef {
   public:
    ConstRowRef(const double* base, Index stride) : base_(base), stride_(stride) {}
    double operator()(Index c) const { return base_[c * stride_]; }
    Index size() const { return C; }

   private:
    const double* base_;
    Index stride_;
  };
  // one column; .array() is the identity (the shim has no separate array world)
  class ConstColView {
   public:
    ConstColView(const double* base, Index n) : base_(base), n_(n) {}
    double operator()(Index i) const { return base_[i]; }
    Index size() const { return n_; }
    const ConstColView& array() const { return *this; }

   private:
    const double* base_;
    Index n_;
  };
  // m.array().colwise() / m.col(j).array(): every column divided element-wise by one column (camera_model.cpp:12)
  class ColwiseView {
   public:
    explicit ColwiseView(const DynRowsMatrix& m) : m_(m) {}
    DynRowsMatrix operator/(const ConstColView& d) const {
      assert(d.size() == m_.rows());
      DynRowsMatrix r(m_.rows(), C);
      for (int c = 0; c < C; ++c)
        for (Index i = 0; i < m_.rows(); ++i) r(i, c) = m_(i, c) / d(i);
      return r;
    }

   private:
    const DynRowsMatrix& m_;
  };
  class ArrayView {
   public:
    explicit ArrayView(const DynRowsMatrix& m) : m_(m) {}
    ColwiseView colwise() const { return ColwiseView(m_); }

   private:
    const DynRowsMatrix& m_;
  };
  // the first k columns, read-only and assignable flavours (a.leftCols(3) = b.leftCols(3), camera_model.cpp:64)
  class ConstLeftColsView {
   public:
    ConstLeftColsView(const DynRowsMatrix& m, Index k) : m_(m), k_(k) {}
    Index rows() const { return m_.rows(); }
    Index cols() const { return k_; }
    double operator()(Index r, Index c) const { return m_(r, c); }

   private:
    const DynRowsMatrix& m_;
    Index k_;
  };
  class LeftColsView {
   public:
    LeftColsView(DynRowsMatrix& m, Index k) : m_(m), k_(k) {}
    LeftColsView& operator=(const ConstLeftColsView& o) {
      assert(o.rows() == m_.rows() && o.cols() == k_);
      for (Index c = 0; c < k_; ++c)
        for (Index r = 0; r < m_.rows(); ++r) m_(r, c) = o(r, c);
      return *this;
    }

   private:
    DynRowsMatrix& m_;
    Index k_;
  };

  DynRowsMatrix() = default;
  DynRowsMatrix(Index rows, Index cols) : rows_(rows), v_(static_cast<size_t>(rows * C)) {
    assert(cols == C);
    (void)cols;
  }
  static DynRowsMatrix Ones(Index rows, Index cols) {
    DynRowsMatrix m(rows, cols);
    std::fill(m.v_.begin(), m.v_.end(), 1.0);
    return m;
  }
  static DynRowsMatrix Zero(Index rows, Index cols) {
    DynRowsMatrix m(rows, cols);
    std::fill(m.v_.begin(), m.v_.end(), 0.0);
    return m;
  }
  Index rows() const { return rows_; }
  Index cols() const { return C; }
  Index size() const { return rows_ * C; }
  double& operator()(Index r, Index c) { return v_[static_cast<size_t>(c * rows_ + r)]; }
  double operator()(Index r, Index c) const { return v_[static_cast<size_t>(c * rows_ + r)]; }
  RowRef row(Index r) { return RowRef(v_.data() + r, rows_); }
  ConstRowRef row(Index r) const { return ConstRowRef(v_.data() + r, rows_); }
  ConstColView col(Index c) const { return ConstColView(v_.data() + c * rows_, rows_); }
  ArrayView array() const { return ArrayView(*this); }
  LeftColsView leftCols(Index k) { return LeftColsView(*this, k); }
  ConstLeftColsView leftCols(Index k) const { return ConstLeftColsView(*this, k); }
  DynColsMatrix<C> transpose() const;
  double sum() const {
    double s = 0;
    for (double x : v_) s += x;
    return s;
  }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }

 private:
  Index rows_ = 0;
  internal::Storage v_;
};

// R x dynamic-columns, COLUMN-major: what `cloud.transpose()` is, so that `T * cloud.transpose()` reads as in the reference.
template <int R>
class DynColsMatrix {
 public:
  DynColsMatrix() = default;
  DynColsMatrix(Index rows, Index cols) : cols_(cols), v_(static_cast<size_t>(R * cols)) {
    assert(rows == R);
    (void)rows;
  }
  Index rows() const { return R; }
  Index cols() const { return cols_; }
  double& operator()(Index r, Index c) { return v_[static_cast<size_t>(c * R + r)]; }
  double operator()(Index r, Index c) const { return v_[static_cast<size_t>(c * R + r)]; }
  DynRowsMatrix<R> transpose() const {
    DynRowsMatrix<R> t(cols_, R);
    for (Index c = 0; c < cols_; ++c)
      for (int r = 0; r < R; ++r) t(c, r) = (*this)(r, c);
    return t;
  }

 private:
  Index cols_ = 0;
  internal::Storage v_;
};

template <int C>
DynColsMatrix<C> DynRowsMatrix<C>::transpose() const {
  DynColsMatrix<C> t(C, rows_);
  for (Index r = 0; r < rows_; ++r)
    for (int c = 0; c < C; ++c) t(c, r) = (*this)(r, c);
  return t;
}

// (M x K) * (K x n)
template <int M, int K>
DynColsMatrix<M> operator*(const Matrix<double, M, K>& a, const DynColsMatrix<K>& b) {
  DynColsMatrix<M> r(M, b.cols());
  for (Index j = 0; j < b.cols(); ++j)
    for (int i = 0; i < M; ++i) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += a(i, k) * b(k, j);
      r(i, j) = acc;
    }
  return r;
}

using MatrixX4d = DynRowsMatrix<4>;
using MatrixX3d = DynRowsMatrix<3>;
using Matrix4Xd = DynColsMatrix<4>;
using Matrix3Xd = DynColsMatrix<3>;

class AngleAxisd {
 public:
  AngleAxisd(double angle, const Vector3d& axis) : angle_(angle), axis_(axis) {}
  double angle() const { return angle_; }
  const Vector3d& axis() const { return axis_; }
  Matrix3d toRotationMatrix() const {
    double const c = std::cos(angle_), s = std::sin(angle_), t = 1.0 - c;
    const Vector3d& a = axis_;
    Matrix3d m;
    m(0, 0) = t * a(0) * a(0) + c;
    m(0, 1) = t * a(0) * a(1) - s * a(2);
    m(0, 2) = t * a(0) * a(2) + s * a(1);
    m(1, 0) = t * a(0) * a(1) + s * a(2);
    m(1, 1) = t * a(1) * a(1) + c;
    m(1, 2) = t * a(1) * a(2) - s * a(0);
    m(2, 0) = t * a(0) * a(2) - s * a(1);
    m(2, 1) = t * a(1) * a(2) + s * a(0);
    m(2, 2) = t * a(2) * a(2) + c;
    return m;
  }
  operator Matrix3d() const { return toRotationMatrix(); }
  Matrix3d operator*(const AngleAxisd& o) const { return toRotationMatrix() * o.toRotationMatrix(); }

 private:
  double angle_;
  Vector3d axis_;
};
inline Matrix3d operator*(const Matrix3d& m, const AngleAxisd& a) { return m * a.toRotationMatrix(); }

// Transform<double, 3, Affine>: linear block + translation; the last row is (0 0 0 1).
class Affine3d {
 public:
  Affine3d() : linear_(Matrix3d::Identity()), translation_() {}
  static Affine3d Identity() { return Affine3d(); }

  Matrix3d& linear() { return linear_; }
  const Matrix3d& linear() const { return linear_; }
  Vector3d& translation() { return translation_; }
  const Vector3d& translation() const { return translation_; }

  // 4x4 homogeneous matrix, column-major (a copy: the shim stores the blocks separately)
  Matrix4d matrix() const {
    Matrix4d m;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) m(r, c) = linear_(r, c);
      m(r, 3) = translation_(r);
    }
    m(3, 3) = 1.0;
    return m;
  }
  static Affine3d FromMatrix(const double* colmajor16) {
    Affine3d t;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) t.linear_(r, c) = colmajor16[c * 4 + r];
      t.translation_(r) = colmajor16[12 + r];
    }
    return t;
  }

  // Affine-mode rotation(): the proper-rotation polar factor of the linear block (Eigen takes an SVD; Newton's
  // iteration X <- (X + X^-T)/2 converges to the same matrix for det > 0).
  Matrix3d rotation() const {
    Matrix3d x = linear_;
    for (int it = 0; it < 50; ++it) {
      Matrix3d const next = (x + x.inverse().transpose()) * 0.5;
      double delta = 0;
      for (Index i = 0; i < 9; ++i) delta = std::fmax(delta, std::fabs(next(i) - x(i)));
      x = next;
      if (delta < 1e-15) break;
    }
    return x;
  }

  // Affine-mode inverse(): general inverse of the linear block, translation -(L^-1) t
  Affine3d inverse() const {
    Affine3d r;
    r.linear_ = linear_.inverse();
    r.translation_ = -(r.linear_ * translation_);
    return r;
  }

  Affine3d operator*(const Affine3d& o) const {
    Affine3d r;
    r.linear_ = linear_ * o.linear_;
    r.translation_ = linear_ * o.translation_ + translation_;
    return r;
  }
  Vector3d operator*(const Vector3d& p) const { return linear_ * p + translation_; }
  // homogeneous point: top rows L v3 + t w, w passes through
  Vector4d operator*(const Vector4d& p) const {
    Vector3d const q = linear_ * Vector3d{p(0), p(1), p(2)} + translation_ * p(3);
    return Vector4d{q(0), q(1), q(2), p(3)};
  }
  // 4 x n block of homogeneous columns: top rows L v3 + t w, the bottom row passes through
  Matrix4Xd operator*(const Matrix4Xd& p) const {
    Matrix4Xd r(4, p.cols());
    for (Index j = 0; j < p.cols(); ++j) {
      for (int i = 0; i < 3; ++i)
        r(i, j) = linear_(i, 0) * p(0, j) + linear_(i, 1) * p(1, j) + linear_(i, 2) * p(2, j) + translation_(i) * p(3, j);
      r(3, j) = p(3, j);
    }
    return r;
  }
  Affine3d& operator*=(const Matrix3d& m) {
    linear_ = linear_ * m;
    return *this;
  }
  Affine3d& rotate(const AngleAxisd& a) { return *this *= a.toRotationMatrix(); }
  Affine3d& rotate(const Matrix3d& m) { return *this *= m; }
  Affine3d& translate(const Vector3d& v) {
    translation_ += linear_ * v;
    return *this;
  }

 private:
  Matrix3d linear_;
  Vector3d translation_;
};

inline Affine3d operator*(const Matrix3d& m, const Affine3d& t) {  // pose = R * pose  (data_io.cpp:84 in the reference)
  Affine3d r;
  r.linear() = m * t.linear();
  r.translation() = m * t.translation();
  return r;
}

}  // namespace Eigen
