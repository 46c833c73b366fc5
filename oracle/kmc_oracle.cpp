// kmc_oracle.cpp — CPU restatement of the reference's per-point LiDAR deskew path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under kitti_motion_compensation_b200/ (the product) includes, links
// or calls this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load the library built from it.
//
// Parity status: PINNED for the synthetic goldens the reference's own tests hold
//   (test/test_motion_compensation.cpp:54-76, test/test_timestamp_mocking.cpp:55-57,71-73,84-86,
//    test/test_lie_algebra.cpp:5-47, test/test_trajectory_interpolation.cpp:43-60,
//    test/test_oxts_to_pose.cpp:17-20) — see tests/test_oracle_golden.py.
//   The reference ships NO golden for a frame with rotation != 0 and its own build cannot run in this image
//   (Eigen3 / OpenCV / GTest absent).  Rotating frames are therefore pinned against the reference's own eight
//   source files compiled unmodified into oracle/_ref/libkmc_ref.so (`make -C oracle ref`: Eigen supplied by
//   this repository's shim when Eigen 3 is missing, OpenCV by a recording stub) — function by function and end
//   to end in tests/test_oracle_vs_reference_sources.py — and, independently of any shared code, against a
//   40-digit matrix expm/logm evaluation in tests/test_oracle_golden.py.
//
// What is restated (file:line are relative to /root/reference):
//   src/kitti_motion_compensation/motion_compensation.cpp:9-28      MotionCompensatePoint / MotionCompensateFrame
//   src/kitti_motion_compensation/trajectory_interpolation.cpp:14-51 TrajectoryInterpolator, InterpolateTrajectory
//   src/kitti_motion_compensation/lie_algebra.cpp:7-103              Hat, Vee, Exp, Log, J, J^-1, SE(3) Exp/Log
//   src/kitti_motion_compensation/timestamp_mocking.cpp:46-63        azimuth -> pseudo time stamp
//   src/kitti_motion_compensation/data_io.cpp:68-88, 253-269         OxtsToPose, MakeFrame (fixture building only)
//   src/kitti_motion_compensation/data_io.cpp:124-135, 299-310       float32 xyzi <-> double column-major cloud
//
// The arithmetic of the reference lives partly in Eigen 3 (unpinned apt libeigen3-dev, 3.4.0 on Ubuntu 22.04),
// which is not vendored under /root/reference.  The Eigen semantics that matter are restated from Eigen's
// published behaviour and named where used:
//   * Transform<double,3,Affine>::inverse()  -> GENERAL 3x3 inverse of the linear block (cofactor formula),
//                                               translation = -(L^-1) t                    [AffineInverse]
//   * Transform::rotation() in Affine mode   -> computeRotationScaling: JacobiSVD polar projection
//                                               U diag(1,1,sign det(U V^T)) V^T            [PolarRotation]
//   * Affine3d * Vector4d                    -> top rows L v3 + t w, last row passes w through [AffineApply]
//   * MatrixX4d                              -> COLUMN-major N x 4 doubles
// The SVD below is a two-sided Jacobi written for this file; it is not Eigen's code and can differ from
// Eigen in the last ulp (effect on deskewed coordinates << 1e-12 m).
//
// The redundant per-point work of the reference (Log, SVD and three general inverses per point) is KEPT on
// purpose: this file is the timed "reference CPU path" of bench.py whenever oracle/_ref has not been built.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace kmc_oracle {

struct Vec3 {
  double v[3];
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
};

struct Mat3 {
  double m[3][3];  // m[row][col]
  double& operator()(int r, int c) { return m[r][c]; }
  double operator()(int r, int c) const { return m[r][c]; }
};

// An Eigen::Affine3d: 4x4 whose last row is fixed to (0 0 0 1).
struct Affine {
  Mat3 L;
  Vec3 t;
};

struct Twist {
  double v[6];  // [rho ; phi]  (data_types.hpp:24-25, lie_algebra.cpp:84-85,99-100)
};

// ---------- tiny fixed-size helpers ------------------------------------------------------------------

static inline Mat3 Identity3() { return Mat3{{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }
static inline Mat3 Zero3() { return Mat3{{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}}; }

static inline Mat3 Add(Mat3 const& a, Mat3 const& b) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j];
  return r;
}
static inline Mat3 Sub(Mat3 const& a, Mat3 const& b) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j];
  return r;
}
static inline Mat3 Scale(double s, Mat3 const& a) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = s * a.m[i][j];
  return r;
}
static inline Mat3 Mul(Mat3 const& a, Mat3 const& b) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k) acc += a.m[i][k] * b.m[k][j];
      r.m[i][j] = acc;
    }
  return r;
}
static inline Mat3 Transpose(Mat3 const& a) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
  return r;
}
static inline Vec3 Mul(Mat3 const& a, Vec3 const& x) {
  Vec3 r;
  for (int i = 0; i < 3; ++i) r.v[i] = a.m[i][0] * x.v[0] + a.m[i][1] * x.v[1] + a.m[i][2] * x.v[2];
  return r;
}
static inline Mat3 Outer(Vec3 const& a, Vec3 const& b) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.v[i] * b.v[j];
  return r;
}
static inline double Norm(Vec3 const& a) { return std::sqrt(a.v[0] * a.v[0] + a.v[1] * a.v[1] + a.v[2] * a.v[2]); }
static inline double Trace(Mat3 const& a) { return a.m[0][0] + a.m[1][1] + a.m[2][2]; }
static inline double Det(Mat3 const& a) {
  return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) -
         a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) +
         a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}

// General 3x3 inverse by cofactors — what Eigen's fixed-size inverse() does for 3x3.
static Mat3 Inverse3(Mat3 const& a) {
  Mat3 c;
  c.m[0][0] = a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1];
  c.m[0][1] = a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2];
  c.m[0][2] = a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1];
  c.m[1][0] = a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2];
  c.m[1][1] = a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0];
  c.m[1][2] = a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2];
  c.m[2][0] = a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0];
  c.m[2][1] = a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1];
  c.m[2][2] = a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0];
  double const det = a.m[0][0] * c.m[0][0] + a.m[1][0] * c.m[0][1] + a.m[2][0] * c.m[0][2];
  return Scale(1.0 / det, c);
}

// ---------- Eigen::Affine3d semantics ---------------------------------------------------------------

static inline Affine AffineIdentity() { return Affine{Identity3(), Vec3{{0, 0, 0}}}; }

// Transform<double,3,Affine>::inverse(): general inverse of the linear part (NOT the transpose).
static Affine AffineInverse(Affine const& a) {
  Affine r;
  r.L = Inverse3(a.L);
  Vec3 const lt = Mul(r.L, a.t);
  r.t = Vec3{{-lt.v[0], -lt.v[1], -lt.v[2]}};
  return r;
}

static Affine AffineMul(Affine const& a, Affine const& b) {
  Affine r;
  r.L = Mul(a.L, b.L);
  Vec3 const lt = Mul(a.L, b.t);
  r.t = Vec3{{lt.v[0] + a.t.v[0], lt.v[1] + a.t.v[1], lt.v[2] + a.t.v[2]}};
  return r;
}

// Affine3d * Vector4d (motion_compensation.cpp:13): top three rows L*v3 + t*w, w passes through.
static void AffineApply(Affine const& a, double const p[4], double out[4]) {
  for (int i = 0; i < 3; ++i) out[i] = a.L.m[i][0] * p[0] + a.L.m[i][1] * p[1] + a.L.m[i][2] * p[2] + a.t.v[i] * p[3];
  out[3] = p[3];
}

// Two-sided Jacobi SVD of a 3x3 (A = U diag(s) V^T), used only for the polar projection below.
static void JacobiSvd3(Mat3 const& a, Mat3& u, double s[3], Mat3& v) {
  Mat3 w = a;
  u = Identity3();
  v = Identity3();
  double const eps = 2.220446049250313e-16;
  for (int sweep = 0; sweep < 64; ++sweep) {
    bool rotated = false;
    for (int p = 1; p < 3; ++p) {
      for (int q = 0; q < p; ++q) {
        double const scale = std::max(std::fabs(w.m[p][p]), std::fabs(w.m[q][q]));
        double const thresh = std::max(2.0 * 2.2250738585072014e-308, eps * scale);
        if (std::fabs(w.m[p][q]) <= thresh && std::fabs(w.m[q][p]) <= thresh) continue;
        rotated = true;
        // 2x2 block [[w_qq w_qp],[w_pq w_pp]] -> first symmetrise with a left rotation, then diagonalise.
        double const bqq = w.m[q][q], bqp = w.m[q][p], bpq = w.m[p][q], bpp = w.m[p][p];
        double c1, s1;  // left rotation that makes the block symmetric
        double const tsum = bqq + bpp, tdiff = bpq - bqp;
        if (std::fabs(tdiff) < 2.2250738585072014e-308) {
          c1 = 1.0;
          s1 = 0.0;
        } else {
          double const h = std::hypot(tsum, tdiff);
          c1 = tsum / h;
          s1 = tdiff / h;
        }
        // Apply G1^T on the left of rows (q,p): rows' = [c1 s1; -s1 c1] rows
        for (int k = 0; k < 3; ++k) {
          double const wq = w.m[q][k], wp = w.m[p][k];
          w.m[q][k] = c1 * wq + s1 * wp;
          w.m[p][k] = -s1 * wq + c1 * wp;
          double const uq = u.m[k][q], up = u.m[k][p];
          u.m[k][q] = c1 * uq + s1 * up;
          u.m[k][p] = -s1 * uq + c1 * up;
        }
        // Now the (q,p) block is symmetric: classic Jacobi rotation.
        double const sqq = w.m[q][q], sqp = w.m[q][p], spp = w.m[p][p];
        double c2 = 1.0, s2 = 0.0;
        if (std::fabs(sqp) > 2.2250738585072014e-308) {
          double const tau = (spp - sqq) / (2.0 * sqp);
          double const t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
          c2 = 1.0 / std::sqrt(1.0 + t * t);
          s2 = t * c2;
        }
        // W <- J^T W J with J = [c2 s2; -s2 c2] acting on (q,p)
        for (int k = 0; k < 3; ++k) {
          double const wq = w.m[q][k], wp = w.m[p][k];
          w.m[q][k] = c2 * wq - s2 * wp;
          w.m[p][k] = s2 * wq + c2 * wp;
          double const uq = u.m[k][q], up = u.m[k][p];
          u.m[k][q] = c2 * uq - s2 * up;
          u.m[k][p] = s2 * uq + c2 * up;
        }
        for (int k = 0; k < 3; ++k) {
          double const wq = w.m[k][q], wp = w.m[k][p];
          w.m[k][q] = c2 * wq - s2 * wp;
          w.m[k][p] = s2 * wq + c2 * wp;
          double const vq = v.m[k][q], vp = v.m[k][p];
          v.m[k][q] = c2 * vq - s2 * vp;
          v.m[k][p] = s2 * vq + c2 * vp;
        }
      }
    }
    if (!rotated) break;
  }
  // make singular values non-negative (flip the matching column of U), then sort descending like Eigen.
  for (int i = 0; i < 3; ++i) {
    s[i] = w.m[i][i];
    if (s[i] < 0) {
      s[i] = -s[i];
      for (int k = 0; k < 3; ++k) u.m[k][i] = -u.m[k][i];
    }
  }
  for (int i = 0; i < 2; ++i) {
    int big = i;
    for (int j = i + 1; j < 3; ++j)
      if (s[j] > s[big]) big = j;
    if (big != i) {
      std::swap(s[i], s[big]);
      for (int k = 0; k < 3; ++k) {
        std::swap(u.m[k][i], u.m[k][big]);
        std::swap(v.m[k][i], v.m[k][big]);
      }
    }
  }
}

// Transform::rotation() for Mode == Affine (lie_algebra.cpp:95): the closest proper rotation to the linear
// block, U diag(1,1,x) V^T with x = sign(det(U V^T)).
static Mat3 PolarRotation(Mat3 const& linear) {
  Mat3 u, v;
  double s[3];
  JacobiSvd3(linear, u, s, v);
  double const x = (Det(Mul(u, Transpose(v))) < 0.0) ? -1.0 : 1.0;
  for (int k = 0; k < 3; ++k) u.m[k][2] *= x;
  return Mul(u, Transpose(v));
}

// ---------- lie_algebra.cpp ---------------------------------------------------------------------------

// lie_algebra.cpp:7-18
static Mat3 Hat(Vec3 const& a) {
  Mat3 h = Zero3();
  h.m[0][1] = -a.v[2];
  h.m[0][2] = a.v[1];
  h.m[1][0] = a.v[2];
  h.m[1][2] = -a.v[0];
  h.m[2][0] = -a.v[1];
  h.m[2][1] = a.v[0];
  return h;
}

// lie_algebra.cpp:20
static Vec3 Vee(Mat3 const& a) { return Vec3{{a.m[2][1], a.m[0][2], a.m[1][0]}}; }

// lie_algebra.cpp:22-35 — Rodrigues; first-order branch below 1e-6 rad.
static Mat3 ExpSo3(Vec3 const& phi) {
  double const angle = Norm(phi);
  if (angle < 1e-6) return Add(Identity3(), Hat(phi));
  Vec3 const axis{{phi.v[0] / angle, phi.v[1] / angle, phi.v[2] / angle}};
  double const c = std::cos(angle);
  double const s = std::sin(angle);
  return Add(Add(Scale(c, Identity3()), Scale(1.0 - c, Outer(axis, axis))), Scale(s, Hat(axis)));
}

// lie_algebra.cpp:37-49
static Vec3 LogSo3(Mat3 const& R) {
  double c = 0.5 * Trace(R) - 0.5;
  c = std::clamp(c, -1.0, 1.0);
  double const angle = std::acos(c);
  if (angle < 1e-6) return Vee(Sub(R, Identity3()));
  return Vee(Scale(0.5 * angle / std::sin(angle), Sub(R, Transpose(R))));
}

// lie_algebra.cpp:51-65
static Mat3 LeftJacobian(Vec3 const& phi) {
  double const angle = Norm(phi);
  if (angle < 1e-6) return Add(Identity3(), Scale(0.5, Hat(phi)));
  Vec3 const axis{{phi.v[0] / angle, phi.v[1] / angle, phi.v[2] / angle}};
  double const c = std::cos(angle);
  double const s = std::sin(angle);
  return Add(Add(Scale(s / angle, Identity3()), Scale(1.0 - (s / angle), Outer(axis, axis))),
             Scale((1 - c) / angle, Hat(axis)));
}

// lie_algebra.cpp:67-81
static Mat3 InverseLeftJacobian(Vec3 const& phi) {
  double const angle = Norm(phi);
  if (angle < 1e-6) return Sub(Identity3(), Scale(0.5, Hat(phi)));
  Vec3 const axis{{phi.v[0] / angle, phi.v[1] / angle, phi.v[2] / angle}};
  double const half = 0.5 * angle;
  double const cot = 1.0 / std::tan(half);
  return Sub(Add(Scale(half * cot, Identity3()), Scale(1 - (half * cot), Outer(axis, axis))), Scale(half, Hat(axis)));
}

// lie_algebra.cpp:83-92
static Affine ExpSe3(Twist const& xi) {
  Vec3 const rho{{xi.v[0], xi.v[1], xi.v[2]}};
  Vec3 const phi{{xi.v[3], xi.v[4], xi.v[5]}};
  Affine T = AffineIdentity();
  T.L = Mul(T.L, ExpSo3(phi));          // T *= Exp(phi)
  T.t = Mul(LeftJacobian(phi), rho);    // T.translation() << J(phi) rho
  return T;
}

// lie_algebra.cpp:94-103
static Twist LogSe3(Affine const& T) {
  Vec3 const phi = LogSo3(PolarRotation(T.L));  // T.rotation() — SVD polar projection in Affine mode
  Vec3 const rho = Mul(InverseLeftJacobian(phi), T.t);
  return Twist{{rho.v[0], rho.v[1], rho.v[2], phi.v[0], phi.v[1], phi.v[2]}};
}

// ---------- trajectory_interpolation.cpp ----------------------------------------------------------------

struct TrajectoryInterpolator {
  double time_1;
  Affine pose_1;
  double time_2;
  Affine pose_2;

  // trajectory_interpolation.cpp:47
  bool TimeIsInRange(double t) const { return (t >= time_1) && (t <= time_2); }
  // trajectory_interpolation.cpp:49-51
  double FractionOfTrajectory(double t) const { return (t - time_1) / (time_2 - time_1); }

  // trajectory_interpolation.cpp:31-41.  The reference asserts (aborts) on out-of-range times even in
  // release builds; the oracle reports it through *ok instead so a test process survives.
  Affine GetPoseAtTime(double t, bool* ok) const {
    if (!TimeIsInRange(t)) *ok = false;
    Twist const f = LogSe3(AffineMul(AffineInverse(pose_1), pose_2));
    double const x = FractionOfTrajectory(t);
    Twist fx;
    for (int i = 0; i < 6; ++i) fx.v[i] = x * f.v[i];
    return AffineMul(pose_1, ExpSe3(fx));
  }

  // trajectory_interpolation.cpp:43-45
  Affine RelativePoseBetweenTimes(double anchor, double query, bool* ok) const {
    return AffineMul(AffineInverse(GetPoseAtTime(anchor, ok)), GetPoseAtTime(query, ok));
  }
};

// ---------- timestamp_mocking.cpp ----------------------------------------------------------------------

// timestamp_mocking.cpp:46
static double FractionOfScanCompleted(double x, double y) { return (M_PI - std::atan2(y, x)) / (2.0 * M_PI); }

// timestamp_mocking.cpp:49-54
static double GetPseudoTimeStamp(double x, double y, double scan_start, double scan_end) {
  double const position_in_scan = FractionOfScanCompleted(x, y);
  double const scan_duration = scan_end - scan_start;
  return scan_start + (position_in_scan * scan_duration);
}

// ---------- data_io.cpp (fixture building only) ------------------------------------------------------------

struct Oxts {
  double stamp, lat, lon, alt, roll, pitch, yaw;
};

static Mat3 RotX(double a) { return Mat3{{{1, 0, 0}, {0, std::cos(a), -std::sin(a)}, {0, std::sin(a), std::cos(a)}}}; }
static Mat3 RotY(double a) { return Mat3{{{std::cos(a), 0, std::sin(a)}, {0, 1, 0}, {-std::sin(a), 0, std::cos(a)}}}; }
static Mat3 RotZ(double a) { return Mat3{{{std::cos(a), -std::sin(a), 0}, {std::sin(a), std::cos(a), 0}, {0, 0, 1}}}; }

// data_io.cpp:68-88.  Eigen multiplies the three AngleAxisd through quaternions; the matrix product below is
// the same rotation to ~1e-16.
static Affine OxtsToPose(Oxts const& o, double scale) {
  double const earth_radius = 6378137.0;
  double const tx = scale * earth_radius * M_PI * o.lon / 180.0;
  double const ty = scale * earth_radius * std::log(std::tan(M_PI * (90.0 + o.lat) / 360.0));
  double const tz = o.alt;
  Affine pose;
  pose.L = Mul(Mul(RotZ(o.yaw), RotY(o.pitch)), RotX(o.roll));
  pose.t = Vec3{{tx, ty, tz}};
  return pose;
}

// trajectory_interpolation.cpp:14-25
static Affine InterpolateTrajectory(Oxts const& o1, Oxts const& o2, double t, bool* ok) {
  TrajectoryInterpolator const interp{o1.stamp, OxtsToPose(o1, 1.0), o2.stamp, OxtsToPose(o2, 1.0)};
  return interp.GetPoseAtTime(t, ok);
}

// ---------- motion_compensation.cpp ------------------------------------------------------------------------

// motion_compensation.cpp:9-14
static void MotionCompensatePoint(TrajectoryInterpolator const& interp, double point_stamp, double const point[4],
                                  double requested_time, double out[4], bool* ok) {
  Affine const correction = interp.RelativePoseBetweenTimes(requested_time, point_stamp, ok);
  AffineApply(correction, point, out);
}

// motion_compensation.cpp:16-28.  cloud/out are COLUMN-major n x 4 (Eigen::MatrixX4d).
static bool MotionCompensateFrame(double const* cloud, double const* timestamps, int64_t n, Affine const& T_start,
                                  Affine const& T_end, double stamp_start, double stamp_end, double requested_time,
                                  double* out) {
  TrajectoryInterpolator const interp{stamp_start, T_start, stamp_end, T_end};
  bool ok = true;
  for (int64_t i = 0; i < n; ++i) {
    double const p[4] = {cloud[i], cloud[n + i], cloud[2 * n + i], cloud[3 * n + i]};
    double q[4];
    MotionCompensatePoint(interp, timestamps[i], p, requested_time, q, &ok);
    out[i] = q[0];
    out[n + i] = q[1];
    out[2 * n + i] = q[2];
    out[3 * n + i] = q[3];
  }
  return ok;
}

static Affine FromColMajor16(double const m[16]) {
  Affine a;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) a.L.m[r][c] = m[c * 4 + r];
    a.t.v[r] = m[12 + r];
  }
  return a;
}
static void ToColMajor16(Affine const& a, double m[16]) {
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) m[c * 4 + r] = a.L.m[r][c];
    m[12 + r] = a.t.v[r];
  }
  m[3] = m[7] = m[11] = 0.0;
  m[15] = 1.0;
}
static Mat3 FromColMajor9(double const m[9]) {
  Mat3 a;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) a.m[r][c] = m[c * 3 + r];
  return a;
}
static void ToColMajor9(Mat3 const& a, double m[9]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m[c * 3 + r] = a.m[r][c];
}

// Whole reference pipeline for one KITTI scan held as float32 xyzi (the on-disk format):
//   KittiPclLoader::LoadPointcloud (data_io.cpp:124-135)  float xyzi -> double col-major + ones column
//   GetPseudoTimeStamps           (data_io.cpp:163, timestamp_mocking.cpp:56-63)
//   MotionCompensateFrame         (motion_compensation.cpp:16-28)
// out_xyz1 is the reference's double result, row-major n x 4 (x', y', z', 1) for easy comparison.
static bool DeskewXyziScan(float const* xyzi, int64_t n, Affine const& T_start, Affine const& T_end, double stamp_start,
                           double stamp_end, double requested_time, double* out_xyz1) {
  std::vector<double> cloud(static_cast<size_t>(4 * n)), stamps(static_cast<size_t>(n)), res(static_cast<size_t>(4 * n));
  for (int64_t i = 0; i < n; ++i) {
    cloud[i] = xyzi[4 * i + 0];
    cloud[n + i] = xyzi[4 * i + 1];
    cloud[2 * n + i] = xyzi[4 * i + 2];
    cloud[3 * n + i] = 1.0;
  }
  for (int64_t i = 0; i < n; ++i) stamps[i] = GetPseudoTimeStamp(cloud[i], cloud[n + i], stamp_start, stamp_end);
  bool const ok =
      MotionCompensateFrame(cloud.data(), stamps.data(), n, T_start, T_end, stamp_start, stamp_end, requested_time, res.data());
  for (int64_t i = 0; i < n; ++i) {
    out_xyz1[4 * i + 0] = res[i];
    out_xyz1[4 * i + 1] = res[n + i];
    out_xyz1[4 * i + 2] = res[2 * n + i];
    out_xyz1[4 * i + 3] = res[3 * n + i];
  }
  return ok;
}

// ---------- camera_model.cpp (the per-point part of the projection, without the cv::circle drawing) ----

// camera_model.cpp:38-95 (ProjectPointcloudOnFrame) followed by :5-36 (ProjectPointcloudOnImage) for ONE camera:
//   X_c00      = tf_c00_lo * (x y z 1)                      :63-75
//   X_rect     = R_rect_00 * X_c00   (rotation as an Affine) :78-81
//   pixels     = P_rect * X_rect ; pixels /= pixels.z        :9-12
//   skip when z_rect < 0.01 or z_rect > max_range or y_rect > 1.25                :16-24
//   color_scale = 255 * z_rect / (max_range - 0.01) ; cv::Point truncates u, v to int   :27-31
// cloud is column-major n x 4 (the 4th column is ignored: the reference overwrites it with ones, :63-64).
// P_rect is 3x4 column-major.  out: u, v (double, before the int truncation), valid (0/1), color_scale.
static void ProjectPointcloud(double const* cloud, int64_t n, Affine const& tf_c00_lo, Mat3 const& R_rect_00, double const P[12],
                              double max_range, double* uv, int32_t* valid, double* color_scale, double* xyz_rect) {
  Affine R_rect = AffineIdentity();
  R_rect.L = Mul(R_rect_00, R_rect.L);  // R_rect_00 = camera_00.R_rect * Identity  (:78-79)
  R_rect.t = Mul(R_rect_00, R_rect.t);
  for (int64_t i = 0; i < n; ++i) {
    double const p[4] = {cloud[i], cloud[n + i], cloud[2 * n + i], 1.0};
    double c00[4], rect[4];
    AffineApply(tf_c00_lo, p, c00);
    AffineApply(R_rect, c00, rect);
    double pix[3];
    for (int r = 0; r < 3; ++r) pix[r] = P[0 * 3 + r] * rect[0] + P[1 * 3 + r] * rect[1] + P[2 * 3 + r] * rect[2] + P[3 * 3 + r] * rect[3];
    uv[2 * i + 0] = pix[0] / pix[2];
    uv[2 * i + 1] = pix[1] / pix[2];
    double const in_front = rect[2], below = rect[1];
    valid[i] = !((in_front < 0.01) || (in_front > max_range) || (below > 1.25));
    color_scale[i] = 255.0 * (in_front / (max_range - 0.01));
    if (xyz_rect) {
      xyz_rect[3 * i + 0] = rect[0];
      xyz_rect[3 * i + 1] = rect[1];
      xyz_rect[3 * i + 2] = rect[2];
    }
  }
}

}  // namespace kmc_oracle

// ============================================================================================================
// C interface for ctypes (tests / bench only)
// ============================================================================================================
using namespace kmc_oracle;

extern "C" {

void kmc_oracle_hat(double const phi[3], double out_colmajor[9]) { ToColMajor9(Hat(Vec3{{phi[0], phi[1], phi[2]}}), out_colmajor); }
void kmc_oracle_vee(double const m_colmajor[9], double out[3]) {
  Vec3 const v = Vee(FromColMajor9(m_colmajor));
  std::memcpy(out, v.v, sizeof(v.v));
}
void kmc_oracle_so3_exp(double const phi[3], double out_colmajor[9]) {
  ToColMajor9(ExpSo3(Vec3{{phi[0], phi[1], phi[2]}}), out_colmajor);
}
void kmc_oracle_so3_log(double const R_colmajor[9], double out[3]) {
  Vec3 const v = LogSo3(FromColMajor9(R_colmajor));
  std::memcpy(out, v.v, sizeof(v.v));
}
void kmc_oracle_left_jacobian(double const phi[3], double out_colmajor[9]) {
  ToColMajor9(LeftJacobian(Vec3{{phi[0], phi[1], phi[2]}}), out_colmajor);
}
void kmc_oracle_inverse_left_jacobian(double const phi[3], double out_colmajor[9]) {
  ToColMajor9(InverseLeftJacobian(Vec3{{phi[0], phi[1], phi[2]}}), out_colmajor);
}
void kmc_oracle_se3_exp(double const xi[6], double T_colmajor[16]) {
  Twist t;
  std::memcpy(t.v, xi, sizeof(t.v));
  ToColMajor16(ExpSe3(t), T_colmajor);
}
void kmc_oracle_se3_log(double const T_colmajor[16], double xi[6]) {
  Twist const t = LogSe3(FromColMajor16(T_colmajor));
  std::memcpy(xi, t.v, sizeof(t.v));
}
void kmc_oracle_polar_rotation(double const L_colmajor[9], double out_colmajor[9]) {
  ToColMajor9(PolarRotation(FromColMajor9(L_colmajor)), out_colmajor);
}
void kmc_oracle_affine_inverse(double const T_colmajor[16], double out_colmajor[16]) {
  ToColMajor16(AffineInverse(FromColMajor16(T_colmajor)), out_colmajor);
}
void kmc_oracle_affine_mul(double const A[16], double const B[16], double out[16]) {
  ToColMajor16(AffineMul(FromColMajor16(A), FromColMajor16(B)), out);
}

// returns 0 ok, 1 if the reference would have aborted (time outside [t1, t2])
int kmc_oracle_pose_at_time(double t1, double const P1[16], double t2, double const P2[16], double t, double out[16]) {
  TrajectoryInterpolator const interp{t1, FromColMajor16(P1), t2, FromColMajor16(P2)};
  bool ok = true;
  ToColMajor16(interp.GetPoseAtTime(t, &ok), out);
  return ok ? 0 : 1;
}
int kmc_oracle_relative_pose_between_times(double t1, double const P1[16], double t2, double const P2[16], double anchor,
                                           double query, double out[16]) {
  TrajectoryInterpolator const interp{t1, FromColMajor16(P1), t2, FromColMajor16(P2)};
  bool ok = true;
  ToColMajor16(interp.RelativePoseBetweenTimes(anchor, query, &ok), out);
  return ok ? 0 : 1;
}

double kmc_oracle_fraction_of_scan_completed(double const point[4]) { return FractionOfScanCompleted(point[0], point[1]); }
double kmc_oracle_pseudo_time_stamp(double const point[4], double scan_start, double scan_end) {
  return GetPseudoTimeStamp(point[0], point[1], scan_start, scan_end);
}
// timestamp_mocking.cpp:56-63; cloud is column-major n x 4
void kmc_oracle_pseudo_time_stamps(double const* cloud_colmajor, int64_t n, double start, double end, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = GetPseudoTimeStamp(cloud_colmajor[i], cloud_colmajor[n + i], start, end);
}

// oxts = {stamp, lat, lon, alt, roll, pitch, yaw}
void kmc_oracle_oxts_to_pose(double const oxts[7], double scale, double out[16]) {
  ToColMajor16(OxtsToPose(Oxts{oxts[0], oxts[1], oxts[2], oxts[3], oxts[4], oxts[5], oxts[6]}, scale), out);
}
int kmc_oracle_interpolate_trajectory(double const o1[7], double const o2[7], double t, double out[16]) {
  bool ok = true;
  ToColMajor16(InterpolateTrajectory(Oxts{o1[0], o1[1], o1[2], o1[3], o1[4], o1[5], o1[6]},
                                     Oxts{o2[0], o2[1], o2[2], o2[3], o2[4], o2[5], o2[6]}, t, &ok),
               out);
  return ok ? 0 : 1;
}
// MakeFrame (data_io.cpp:253-269): start pose from (o[n-1], o[n]) at stamp_start, end pose from (o[n], o[n+1]) at stamp_end.
int kmc_oracle_make_frame_poses(double const o_prev[7], double const o_cur[7], double const o_next[7], double stamp_start,
                                double stamp_end, double T_start[16], double T_end[16]) {
  int const a = kmc_oracle_interpolate_trajectory(o_prev, o_cur, stamp_start, T_start);
  int const b = kmc_oracle_interpolate_trajectory(o_cur, o_next, stamp_end, T_end);
  return a | b;
}

int kmc_oracle_motion_compensate_point(double t1, double const P1[16], double t2, double const P2[16], double point_stamp,
                                       double const point[4], double requested_time, double out[4]) {
  TrajectoryInterpolator const interp{t1, FromColMajor16(P1), t2, FromColMajor16(P2)};
  bool ok = true;
  MotionCompensatePoint(interp, point_stamp, point, requested_time, out, &ok);
  return ok ? 0 : 1;
}

// The reference's MotionCompensateFrame on its own data layout (column-major double n x 4 + per-point stamps).
int kmc_oracle_motion_compensate_frame(double const* cloud_colmajor, double const* timestamps, int64_t n,
                                       double const T_start[16], double const T_end[16], double stamp_start,
                                       double stamp_end, double requested_time, double* out_colmajor) {
  return MotionCompensateFrame(cloud_colmajor, timestamps, n, FromColMajor16(T_start), FromColMajor16(T_end), stamp_start,
                               stamp_end, requested_time, out_colmajor)
             ? 0
             : 1;
}

// load-convert + pseudo stamps + deskew for one float32 xyzi scan; out is row-major n x 4 doubles (x', y', z', 1).
int kmc_oracle_deskew_xyzi_scan(float const* xyzi, int64_t n, double const T_start[16], double const T_end[16],
                                double stamp_start, double stamp_end, double requested_time, double* out_xyz1) {
  return DeskewXyziScan(xyzi, n, FromColMajor16(T_start), FromColMajor16(T_end), stamp_start, stamp_end, requested_time,
                        out_xyz1)
             ? 0
             : 1;
}

// Timed CPU baseline: n_frames scans of points_per_frame float32 xyzi each (concatenated), per-frame poses
// (16 doubles each, column-major) and stamps {start, end, requested} (3 doubles each).  Frames are handed to
// n_threads std::threads one frame at a time (n_threads == 1 is the reference's own execution model).
// Returns elapsed seconds; checksum receives the sum of all output coordinates so the work cannot be elided.
double kmc_oracle_timed_frames(float const* xyzi, int64_t points_per_frame, int32_t n_frames, double const* T_start,
                               double const* T_end, double const* stamps3, int32_t n_threads, double* checksum) {
  if (n_threads < 1) n_threads = 1;
  std::atomic<int32_t> next{0};
  std::vector<double> partial(static_cast<size_t>(n_threads), 0.0);
  auto worker = [&](int tid) {
    std::vector<double> out(static_cast<size_t>(4 * points_per_frame));
    double acc = 0.0;
    for (;;) {
      int32_t const f = next.fetch_add(1);
      if (f >= n_frames) break;
      DeskewXyziScan(xyzi + static_cast<int64_t>(f) * points_per_frame * 4, points_per_frame, FromColMajor16(T_start + 16 * f),
                     FromColMajor16(T_end + 16 * f), stamps3[3 * f + 0], stamps3[3 * f + 1], stamps3[3 * f + 2], out.data());
      for (int64_t i = 0; i < 4 * points_per_frame; ++i) acc += out[i];
    }
    partial[tid] = acc;
  };
  auto const t0 = std::chrono::steady_clock::now();
  if (n_threads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(worker, t);
    for (auto& th : pool) th.join();
  }
  auto const t1 = std::chrono::steady_clock::now();
  double total = 0.0;
  for (double p : partial) total += p;
  if (checksum) *checksum = total;
  return std::chrono::duration<double>(t1 - t0).count();
}

// cloud column-major n x 4; tf 4x4, R_rect 3x3, P_rect 3x4 all column-major; xyz_rect may be NULL
void kmc_oracle_project_pointcloud(double const* cloud_colmajor, int64_t n, double const tf_c00_lo[16], double const R_rect_00[9],
                                   double const P_rect[12], double max_range, double* uv, int32_t* valid, double* color_scale,
                                   double* xyz_rect) {
  ProjectPointcloud(cloud_colmajor, n, FromColMajor16(tf_c00_lo), FromColMajor9(R_rect_00), P_rect, max_range, uv, valid, color_scale,
                    xyz_rect);
}

int kmc_oracle_hardware_threads(void) { return static_cast<int>(std::thread::hardware_concurrency()); }

}  // extern "C"
