// kmc_host_math.cpp — see kmc_host_math.hpp.  Product code: does not include or call anything under oracle/.
#include "kmc_host_math.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace kmc_b200::host {
namespace {

constexpr double kSmallAngle = 1e-6;  // the reference's Taylor switch (lie_algebra.cpp:25,43,54,70)

inline double& At3(double* m, int r, int c) { return m[c * 3 + r]; }
inline double At3(const double* m, int r, int c) { return m[c * 3 + r]; }

inline void Identity3(double m[9]) {
  std::memset(m, 0, 9 * sizeof(double));
  m[0] = m[4] = m[8] = 1.0;
}

inline double Norm3(const double v[3]) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

inline void MatVec3(const double m[9], const double v[3], double out[3]) {
  for (int r = 0; r < 3; ++r) out[r] = At3(m, r, 0) * v[0] + At3(m, r, 1) * v[1] + At3(m, r, 2) * v[2];
}

inline void MatMul3(const double a[9], const double b[9], double out[9]) {
  double tmp[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) At3(tmp, r, c) = At3(a, r, 0) * At3(b, 0, c) + At3(a, r, 1) * At3(b, 1, c) + At3(a, r, 2) * At3(b, 2, c);
  std::memcpy(out, tmp, sizeof(tmp));
}

inline double Det3(const double m[9]) {
  return At3(m, 0, 0) * (At3(m, 1, 1) * At3(m, 2, 2) - At3(m, 1, 2) * At3(m, 2, 1)) -
         At3(m, 0, 1) * (At3(m, 1, 0) * At3(m, 2, 2) - At3(m, 1, 2) * At3(m, 2, 0)) +
         At3(m, 0, 2) * (At3(m, 1, 0) * At3(m, 2, 1) - At3(m, 1, 1) * At3(m, 2, 0));
}

// adjugate / determinant
bool Inverse3(const double m[9], double out[9]) {
  double adj[9];
  At3(adj, 0, 0) = At3(m, 1, 1) * At3(m, 2, 2) - At3(m, 1, 2) * At3(m, 2, 1);
  At3(adj, 0, 1) = At3(m, 0, 2) * At3(m, 2, 1) - At3(m, 0, 1) * At3(m, 2, 2);
  At3(adj, 0, 2) = At3(m, 0, 1) * At3(m, 1, 2) - At3(m, 0, 2) * At3(m, 1, 1);
  At3(adj, 1, 0) = At3(m, 1, 2) * At3(m, 2, 0) - At3(m, 1, 0) * At3(m, 2, 2);
  At3(adj, 1, 1) = At3(m, 0, 0) * At3(m, 2, 2) - At3(m, 0, 2) * At3(m, 2, 0);
  At3(adj, 1, 2) = At3(m, 0, 2) * At3(m, 1, 0) - At3(m, 0, 0) * At3(m, 1, 2);
  At3(adj, 2, 0) = At3(m, 1, 0) * At3(m, 2, 1) - At3(m, 1, 1) * At3(m, 2, 0);
  At3(adj, 2, 1) = At3(m, 0, 1) * At3(m, 2, 0) - At3(m, 0, 0) * At3(m, 2, 1);
  At3(adj, 2, 2) = At3(m, 0, 0) * At3(m, 1, 1) - At3(m, 0, 1) * At3(m, 1, 0);
  double const det = At3(m, 0, 0) * At3(adj, 0, 0) + At3(m, 1, 0) * At3(adj, 0, 1) + At3(m, 2, 0) * At3(adj, 0, 2);
  if (!(std::fabs(det) > 0.0) || !std::isfinite(det)) return false;
  double const inv = 1.0 / det;
  for (int i = 0; i < 9; ++i) out[i] = adj[i] * inv;
  return true;
}

// Orthogonal polar factor of a 3x3 with positive determinant by Newton's iteration X <- (X + X^-T)/2.
// For det > 0 this is the proper rotation U V^T that Eigen's Affine-mode rotation() extracts with an SVD.
bool PolarRotation(const double linear[9], double out[9]) {
  for (int i = 0; i < 9; ++i)
    if (!std::isfinite(linear[i])) return false;
  if (!(Det3(linear) > 0.0)) return false;
  double x[9];
  std::memcpy(x, linear, sizeof(x));
  for (int it = 0; it < 100; ++it) {
    double inv[9];
    if (!Inverse3(x, inv)) return false;
    double next[9], delta = 0.0, scale = 0.0;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        At3(next, r, c) = 0.5 * (At3(x, r, c) + At3(inv, c, r));
        delta = std::max(delta, std::fabs(At3(next, r, c) - At3(x, r, c)));
        scale = std::max(scale, std::fabs(At3(next, r, c)));
      }
    std::memcpy(x, next, sizeof(x));
    if (delta <= 4.0 * 2.220446049250313e-16 * scale) break;
  }
  std::memcpy(out, x, sizeof(x));
  return true;
}

inline void Split(const double T[16], double L[9], double t[3]) {
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) At3(L, r, c) = T[c * 4 + r];
  for (int r = 0; r < 3; ++r) t[r] = T[12 + r];
}
inline void Join(const double L[9], const double t[3], double T[16]) {
  for (int c = 0; c < 3; ++c) {
    for (int r = 0; r < 3; ++r) T[c * 4 + r] = At3(L, r, c);
    T[c * 4 + 3] = 0.0;
  }
  for (int r = 0; r < 3; ++r) T[12 + r] = t[r];
  T[15] = 1.0;
}

// cos I + (1-cos) a a^T + sin a^   /   (sin/th) I + (1 - sin/th) a a^T + ((1-cos)/th) a^   and friends:
// every closed form of the reference is  alpha I + beta a a^T + gamma a^ .
inline void AxisForm(double alpha, double beta, double gamma, const double a[3], double out[9]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) At3(out, r, c) = beta * a[r] * a[c] + (r == c ? alpha : 0.0);
  At3(out, 0, 1) += -gamma * a[2];
  At3(out, 0, 2) += gamma * a[1];
  At3(out, 1, 0) += gamma * a[2];
  At3(out, 1, 2) += -gamma * a[0];
  At3(out, 2, 0) += -gamma * a[1];
  At3(out, 2, 1) += gamma * a[0];
}

}  // namespace

void So3Hat(const double phi[3], double out[9]) {
  AxisForm(0.0, 0.0, 1.0, phi, out);
}

void So3Vee(const double m[9], double out[3]) {
  out[0] = At3(m, 2, 1);
  out[1] = At3(m, 0, 2);
  out[2] = At3(m, 1, 0);
}

void So3Exp(const double phi[3], double out[9]) {
  double const th = Norm3(phi);
  if (th < kSmallAngle) {
    AxisForm(1.0, 0.0, 1.0, phi, out);  // I + phi^
    return;
  }
  double const a[3] = {phi[0] / th, phi[1] / th, phi[2] / th};
  double const c = std::cos(th), s = std::sin(th);
  AxisForm(c, 1.0 - c, s, a, out);
}

void So3Log(const double R[9], double out[3]) {
  double c = 0.5 * (R[0] + R[4] + R[8]) - 0.5;
  c = std::min(1.0, std::max(-1.0, c));
  double const th = std::acos(c);
  // vee(R - I) == vee(R) off the diagonal; vee(k (R - R^T)) = k (vee(R) - vee(R^T))
  if (th < kSmallAngle) {
    So3Vee(R, out);
    return;
  }
  double const k = 0.5 * th / std::sin(th);
  out[0] = k * (At3(R, 2, 1) - At3(R, 1, 2));
  out[1] = k * (At3(R, 0, 2) - At3(R, 2, 0));
  out[2] = k * (At3(R, 1, 0) - At3(R, 0, 1));
}

void So3LeftJacobian(const double phi[3], double out[9]) {
  double const th = Norm3(phi);
  if (th < kSmallAngle) {
    AxisForm(1.0, 0.0, 0.5, phi, out);  // I + phi^/2
    return;
  }
  double const a[3] = {phi[0] / th, phi[1] / th, phi[2] / th};
  double const c = std::cos(th), s = std::sin(th);
  AxisForm(s / th, 1.0 - s / th, (1.0 - c) / th, a, out);
}

void So3InverseLeftJacobian(const double phi[3], double out[9]) {
  double const th = Norm3(phi);
  if (th < kSmallAngle) {
    AxisForm(1.0, 0.0, -0.5, phi, out);  // I - phi^/2
    return;
  }
  double const a[3] = {phi[0] / th, phi[1] / th, phi[2] / th};
  double const half = 0.5 * th;
  double const hc = half / std::tan(half);
  AxisForm(hc, 1.0 - hc, -half, a, out);
}

void Se3Exp(const double xi[6], double T[16]) {
  double R[9], J[9], t[3];
  So3Exp(xi + 3, R);
  So3LeftJacobian(xi + 3, J);
  MatVec3(J, xi, t);
  Join(R, t, T);
}

bool Se3Log(const double T[16], double xi[6]) {
  double L[9], t[3], R[9], Jinv[9];
  Split(T, L, t);
  for (int i = 0; i < 3; ++i)
    if (!std::isfinite(t[i])) return false;
  if (!PolarRotation(L, R)) return false;
  So3Log(R, xi + 3);
  So3InverseLeftJacobian(xi + 3, Jinv);
  MatVec3(Jinv, t, xi);
  return true;
}

bool AffineInverse(const double T[16], double out[16]) {
  double L[9], t[3], Li[9], ti[3];
  Split(T, L, t);
  if (!Inverse3(L, Li)) return false;
  MatVec3(Li, t, ti);
  for (int i = 0; i < 3; ++i) ti[i] = -ti[i];
  Join(Li, ti, out);
  return true;
}

void AffineMul(const double A[16], const double B[16], double out[16]) {
  double La[9], ta[3], Lb[9], tb[3], L[9], t[3];
  Split(A, La, ta);
  Split(B, Lb, tb);
  MatMul3(La, Lb, L);
  MatVec3(La, tb, t);
  for (int i = 0; i < 3; ++i) t[i] += ta[i];
  Join(L, t, out);
}

bool RelativeTwist(const double P1[16], const double P2[16], double xi[6]) {
  double L1[9], t1[3], L2[9], t2[3], L1i[9], rel[16], L[9], d[3], t[3];
  Split(P1, L1, t1);
  Split(P2, L2, t2);
  if (!Inverse3(L1, L1i)) return false;
  MatMul3(L1i, L2, L);
  for (int i = 0; i < 3; ++i) d[i] = t2[i] - t1[i];
  MatVec3(L1i, d, t);
  Join(L, t, rel);
  return Se3Log(rel, xi);
}

int PoseAtTime(double t1, const double P1[16], double t2, const double P2[16], double t, double out[16]) {
  if (!(t2 > t1)) return KMC_B200_ERR_EMPTY_INTERVAL;
  if (!(t >= t1 && t <= t2)) return KMC_B200_ERR_TIME_OUT_OF_RANGE;
  double xi[6], inv1[16], rel[16], step[16];
  // Follows the reference literally (P1^-1 P2 as a product of 4x4s) because callers compare poses, not deltas.
  if (!AffineInverse(P1, inv1)) return KMC_B200_ERR_NOT_RIGID;
  AffineMul(inv1, P2, rel);
  if (!Se3Log(rel, xi)) return KMC_B200_ERR_NOT_RIGID;
  double const x = (t - t1) / (t2 - t1);
  for (int i = 0; i < 6; ++i) xi[i] *= x;
  Se3Exp(xi, step);
  AffineMul(P1, step, out);
  return KMC_B200_OK;
}

void FrameParamsFromTwist(const double xi[6], double x_req, kmc_b200_frame_params* out) {
  const double* rho = xi;
  const double* phi = xi + 3;
  double const th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  double const th = std::sqrt(th2);
  double par[3] = {0, 0, 0};
  // Below 1e-12 rad per scan the rotation moves a point at 120 m by 1e-10 m and phi is rounding noise of the pose product
  // (a pure translation between two Mercator-magnitude poses leaves |phi| ~ 1e-17): no axis, the whole of rho goes to rho_perp,
  // so the record of a pure translation does not depend on where on the globe the frame sits.
  if (th2 > 1e-24) {
    double const k = (phi[0] * rho[0] + phi[1] * rho[1] + phi[2] * rho[2]) / th2;  // (a.rho)/theta
    for (int i = 0; i < 3; ++i) par[i] = k * phi[i];
  }
  double const cross[3] = {phi[1] * rho[2] - phi[2] * rho[1], phi[2] * rho[0] - phi[0] * rho[2],
                           phi[0] * rho[1] - phi[1] * rho[0]};
  for (int i = 0; i < 3; ++i) {
    out->phi[i] = static_cast<float>(phi[i]);
    out->rho_perp[i] = static_cast<float>(rho[i] - par[i]);
    out->rho_par[i] = static_cast<float>(par[i]);
    out->phi_x_rho[i] = static_cast<float>(cross[i]);
  }
  out->theta2 = static_cast<float>(th2);
  out->c0 = static_cast<float>(0.5 - x_req);
  out->x_req = static_cast<float>(x_req);
  out->wide = (out->theta2 > KMC_B200_SERIES_THETA2_MAX) ? 1.0f : 0.0f;
}

void OxtsToPose(double lat, double lon, double alt, double roll, double pitch, double yaw, double scale, double T[16]) {
  constexpr double kEarthRadius = 6378137.0;  // metres
  double const cr = std::cos(roll), sr = std::sin(roll), cp = std::cos(pitch), sp = std::sin(pitch), cy = std::cos(yaw),
               sy = std::sin(yaw);
  // R = Rz(yaw) Ry(pitch) Rx(roll), written out
  double R[9];
  At3(R, 0, 0) = cy * cp;
  At3(R, 0, 1) = cy * sp * sr - sy * cr;
  At3(R, 0, 2) = cy * sp * cr + sy * sr;
  At3(R, 1, 0) = sy * cp;
  At3(R, 1, 1) = sy * sp * sr + cy * cr;
  At3(R, 1, 2) = sy * sp * cr - cy * sr;
  At3(R, 2, 0) = -sp;
  At3(R, 2, 1) = cp * sr;
  At3(R, 2, 2) = cp * cr;
  double const t[3] = {scale * kEarthRadius * M_PI * lon / 180.0,
                       scale * kEarthRadius * std::log(std::tan(M_PI * (90.0 + lat) / 360.0)), alt};
  Join(R, t, T);
}

void CameraParamsFromCalibration(const double P_rect[12], const double R_rect_00[9], const double T_velo_to_cam[16],
                                 double max_range, kmc_b200_camera_params* out) {
  double L[9], t[3], rect[3][4], pix[3][4];
  Split(T_velo_to_cam, L, t);
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c)
      rect[r][c] = At3(R_rect_00, r, 0) * At3(L, 0, c) + At3(R_rect_00, r, 1) * At3(L, 1, c) + At3(R_rect_00, r, 2) * At3(L, 2, c);
    rect[r][3] = At3(R_rect_00, r, 0) * t[0] + At3(R_rect_00, r, 1) * t[1] + At3(R_rect_00, r, 2) * t[2];
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      // P_rect is 3x4 column-major: P(r, k) = P_rect[k * 3 + r]; the homogeneous row of [rect; 0 0 0 1] only feeds column 3
      double acc = P_rect[0 * 3 + r] * rect[0][c] + P_rect[1 * 3 + r] * rect[1][c] + P_rect[2 * 3 + r] * rect[2][c];
      if (c == 3) acc += P_rect[3 * 3 + r];
      pix[r][c] = acc;
    }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      out->rect[r * 4 + c] = static_cast<float>(rect[r][c]);
      out->pix[r * 4 + c] = static_cast<float>(pix[r][c]);
    }
  out->min_depth = 0.01f;
  out->max_range = static_cast<float>(max_range);
  out->max_below = 1.25f;
  out->color_gain = static_cast<float>(255.0 / (max_range - 0.01));
}

double FractionOfScanCompleted(double x, double y) { return (M_PI - std::atan2(y, x)) / (2.0 * M_PI); }

}  // namespace kmc_b200::host
