// kmc_point_math.cuh — the per-point device math of the deskew path, shared by every kernel variant.
//
// Replaces, per point, the reference's FractionOfScanCompleted (timestamp_mocking.cpp:46), the fractional
// Exp(s xi) of TrajectoryInterpolator::GetPoseAtTime / RelativePoseBetweenTimes (trajectory_interpolation.cpp:31-45,
// lie_algebra.cpp:22-35,51-65,83-92) and the transform apply of MotionCompensatePoint (motion_compensation.cpp:9-14).
#pragma once

#include <cuda_runtime.h>

#include "kmc_b200.h"

namespace kmc_b200::dev {

// ---------------------------------------------------------------------------------------------------------------
// per-point math
// ---------------------------------------------------------------------------------------------------------------

// atan2(y, x) / (2 pi) in "turns", in (-0.5, 0.5], with atan2's signed-zero conventions (the real scan contains
// y == -0.0f, x < 0, which the reference maps to fraction 1.0; SURVEY 8c edge case i).
// atan(r)/(2 pi r) on [0,1] as a degree-7 minimax polynomial in r^2 (|err| < 6e-9 turns in exact arithmetic,
// < 3e-8 turns = 1.7e-7 rad evaluated in fp32).
__device__ __forceinline__ float Atan2Turns(float y, float x) {
  float const ax = fabsf(x), ay = fabsf(y);
  float const mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  float inv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(mx));  // one MUFU.RCP, 1 ulp
  // atan2(+-0, +-0) must not produce 0 * inf; points closer than 1e-30 m to the spin axis count as on an axis.
  float const r = (mx > 1e-30f) ? mn * inv : 0.0f;
  float const t = r * r;
  float p = -6.453042151e-04f;
  p = fmaf(p, t, 3.479597159e-03f);
  p = fmaf(p, t, -8.898722008e-03f);
  p = fmaf(p, t, 1.534603257e-02f);
  p = fmaf(p, t, -2.213627100e-02f);
  p = fmaf(p, t, 3.174594417e-02f);
  p = fmaf(p, t, -5.304612219e-02f);
  p = fmaf(p, t, 1.591548324e-01f);
  float q = r * p;                                    // [0, 1/8]
  q = (ay > ax) ? (0.25f - q) : q;                    // [0, 1/4]
  q = (__float_as_uint(x) >> 31) ? (0.5f - q) : q;    // sign BIT of x: atan2(+-0, -0) = +-pi
  return copysignf(q, y);
}

// atan(r)/(2 pi r) on [0,1] as a degree-12 interpolant in t = r^2 at Chebyshev nodes: |err| < 1e-12 turns evaluated in double,
// i.e. < 1e-13 s on a 0.1 s scan (1 ulp of KITTI's stamps, ~4.7e4 s, is 7e-12 s).  libdevice's double atan2 costs ~120 FP64
// instructions per point and made the stamp kernels FP64-bound (3.5 TB/s).
__device__ __forceinline__ double AtanOverTwoPiR(double t) {
  double p = 6.67221862187784099e-05;
  p = fma(p, t, -5.08567722343806756e-04);
  p = fma(p, t, 1.81876432393576174e-03);
  p = fma(p, t, -4.13898280986435812e-03);
  p = fma(p, t, 6.94657045892680715e-03);
  p = fma(p, t, -9.58462323410399705e-03);
  p = fma(p, t, 1.19290805724528125e-02);
  p = fma(p, t, -1.44022158368946745e-02);
  p = fma(p, t, 1.76746446101705770e-02);
  p = fma(p, t, -2.27356426872162426e-02);
  p = fma(p, t, 3.18309541393976964e-02);
  p = fma(p, t, -5.30516470898420370e-02);
  p = fma(p, t, 1.59154943090102946e-01);
  return p;
}

// atan2(y, x) / (2 pi) in turns for DOUBLE coordinates (the reference's column-major cloud), for the GetPseudoTimeStamps entry
// points (timestamp_mocking.cpp:56-63).  The quotient min/max comes from the fp32 reciprocal refined by one Newton step in
// double (relative error ~1e-14); magnitudes outside the fp32 range take the exact quotient.  A variant with the octant logic
// in fp32 (as FractionOfScanF64 below) was measured and is SLOWER here — 3.57 vs 5.39 TB/s: the extra f32 <-> f64 conversions
// cost more than the FP64 compares they replace (profiles/r02_sweep_secondary.log vs the run after it).
__device__ __forceinline__ double Atan2TurnsF64(double y, double x) {
  double const ax = fabs(x), ay = fabs(y);
  double const mx = fmax(ax, ay), mn = fmin(ax, ay);
  double r;
  if (mx > 1e-30 && mx < 1e30) {
    float seed;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(seed) : "f"(static_cast<float>(mx)));
    double const x0 = static_cast<double>(seed);
    r = mn * (x0 * fma(-mx, x0, 2.0));
  } else {
    // atan2(+-0, +-0): no 0/0; inf/inf counts as the diagonal; values outside the fp32 range take the exact quotient
    r = (mx > 0.0) ? ((mn == mx) ? 1.0 : mn / mx) : 0.0;
  }
  double q = r * AtanOverTwoPiR(r * r);
  q = (ay > ax) ? (0.25 - q) : q;
  q = (__double2hiint(x) < 0) ? (0.5 - q) : q;  // sign BIT of x
  q = (x != x || y != y) ? __longlong_as_double(0x7ff8000000000000ll) : q;  // fmax / fmin drop a NaN operand, atan2 does not
  return copysign(q, y);
}

// FractionOfScanCompleted (timestamp_mocking.cpp:46) in double for a point whose coordinates ARE floats (the .bin
// layout): the octant logic — |x| vs |y|, the sign bits — is exact in fp32, so only the quotient, the polynomial and the
// final fold run in double.  The two folds collapse into frac = A + B q with A, B small exact constants
// picked in fp32, which keeps q = 0 (a point on an axis) exact: frac = 1 for y = -0, x < 0; 0 for y = +0, x < 0; 0.5 for
// x = y = 0.  ~45 instructions per point instead of ~150.  Domain: |x|, |y| < 1e30 (larger values give r = 0).
__device__ __forceinline__ double FractionOfScanF64(float y, float x) {
  float const ax = fabsf(x), ay = fabsf(y);
  float const mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  float seed;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(seed) : "f"(mx));
  seed = (mx > 1e-30f) ? seed : 0.0f;  // atan2(+-0, +-0): r = 0, no 0 * inf
  double const x0 = static_cast<double>(seed);
  double const dmx = static_cast<double>(mx);
  double const r = static_cast<double>(mn) * (x0 * fma(-dmx, x0, 2.0));  // one Newton step: relative error ~1e-14
  double const q = r * AtanOverTwoPiR(r * r);  // atan(min/max) / 2 pi in [0, 1/8]
  // turns = sy (c + m q);  frac = 0.5 - turns = (0.5 - sy c) + (-sy m) q
  bool const steep = ay > ax, back = (__float_as_uint(x) >> 31) != 0;
  float const c = steep ? 0.25f : (back ? 0.5f : 0.0f);
  float const m = (steep != back) ? -1.0f : 1.0f;
  float const sy = (__float_as_uint(y) >> 31) ? -1.0f : 1.0f;
  // fmaxf / fminf drop a NaN operand; atan2 (and the reference's stamp) does not
  float const a = (x != x || y != y) ? __int_as_float(0x7fc00000) : 0.5f - sy * c;
  return fma(static_cast<double>(-sy * m), q, static_cast<double>(a));
}

// S = sin(s th)/th and C = (1 - cos(s th))/th^2 for x2 = (s th)^2 <= 1, as s*P(x2) and s^2*Q(x2).
__device__ __forceinline__ void SeriesSC(float s, float s2, float x2, float& S, float& C) {
  float ps = 2.755731922e-06f;  // 1/9!
  ps = fmaf(ps, x2, -1.984126984e-04f);
  ps = fmaf(ps, x2, 8.333333333e-03f);
  ps = fmaf(ps, x2, -1.666666667e-01f);
  ps = fmaf(ps, x2, 1.0f);
  float pc = 2.755731922e-07f;  // 1/10!
  pc = fmaf(pc, x2, -2.480158730e-05f);
  pc = fmaf(pc, x2, 1.388888889e-03f);
  pc = fmaf(pc, x2, -4.166666667e-02f);
  pc = fmaf(pc, x2, 0.5f);
  S = s * ps;
  C = s2 * pc;
}

// The same S and C for any scan rotation up to pi: with y = s th / 2 (|y| <= pi/2),
//   hs = sin(y)/th = (s/2) Ps(y^2),  ch = cos(y) = Pc(y^2),  S = 2 hs ch,  C = 2 hs^2      (no division by th either)
// Taylor to y^13 / y^14: truncation < 7e-10 at |y| = pi/2.  Taken only when a scan rotates by more than 1 rad.
__device__ __forceinline__ void HalfAngleSC(float s, float x2, float& S, float& C) {
  float const y2 = 0.25f * x2;
  float ps = 1.605904384e-10f;  // 1/13!
  ps = fmaf(ps, y2, -2.505210839e-08f);
  ps = fmaf(ps, y2, 2.755731922e-06f);
  ps = fmaf(ps, y2, -1.984126984e-04f);
  ps = fmaf(ps, y2, 8.333333333e-03f);
  ps = fmaf(ps, y2, -1.666666667e-01f);
  ps = fmaf(ps, y2, 1.0f);
  float pc = -1.147074560e-11f;  // -1/14!
  pc = fmaf(pc, y2, 2.087675699e-09f);
  pc = fmaf(pc, y2, -2.755731922e-07f);
  pc = fmaf(pc, y2, 2.480158730e-05f);
  pc = fmaf(pc, y2, -1.388888889e-03f);
  pc = fmaf(pc, y2, 4.166666667e-02f);
  pc = fmaf(pc, y2, -0.5f);
  pc = fmaf(pc, y2, 1.0f);
  float const hs = (0.5f * s) * ps;
  S = 2.0f * hs * pc;
  C = 2.0f * hs * hs;
}

// Displacement of a point: delta = Exp(s xi) p - p, with s the signed fraction of the scan between the requested time and
// the point's capture time (everything in fp32; see DESIGN.md §3 for why the displacement form).
__device__ __forceinline__ float3 DeskewDelta(float x, float y, float z, float s, const kmc_b200_frame_params& P) {
  float const s2 = s * s;
  float S, C;
  if (P.wide == 0.0f) {  // frame-uniform branch
    SeriesSC(s, s2, s2 * P.theta2, S, C);
  } else {
    HalfAngleSC(s, s2 * P.theta2, S, C);
  }
  float const d = fmaf(P.phi[2], z, fmaf(P.phi[1], y, P.phi[0] * x));
  // u = phi (phi.p) - th^2 p + phi x rho
  float const ux = fmaf(P.phi[0], d, fmaf(-P.theta2, x, P.phi_x_rho[0]));
  float const uy = fmaf(P.phi[1], d, fmaf(-P.theta2, y, P.phi_x_rho[1]));
  float const uz = fmaf(P.phi[2], d, fmaf(-P.theta2, z, P.phi_x_rho[2]));
  // v = phi x p + rho_perp
  float const vx = fmaf(P.phi[1], z, fmaf(-P.phi[2], y, P.rho_perp[0]));
  float const vy = fmaf(P.phi[2], x, fmaf(-P.phi[0], z, P.rho_perp[1]));
  float const vz = fmaf(P.phi[0], y, fmaf(-P.phi[1], x, P.rho_perp[2]));
  return make_float3(fmaf(C, ux, fmaf(S, vx, s * P.rho_par[0])), fmaf(C, uy, fmaf(S, vy, s * P.rho_par[1])),
                     fmaf(C, uz, fmaf(S, vz, s * P.rho_par[2])));
}

// The same for a homogeneous point (x y z w) with w != 1: the reference applies the correction as an Affine3d times a
// Vector4d (motion_compensation.cpp:13), i.e. R p + t w with w passed through, so the translation terms scale with w.
// For w == 1 every operation rounds exactly as in DeskewDelta (w * c == c), so both give the same bits.
__device__ __forceinline__ float3 DeskewDeltaW(float x, float y, float z, float w, float s, const kmc_b200_frame_params& P) {
  float const s2 = s * s;
  float S, C;
  if (P.wide == 0.0f) {
    SeriesSC(s, s2, s2 * P.theta2, S, C);
  } else {
    HalfAngleSC(s, s2 * P.theta2, S, C);
  }
  float const d = fmaf(P.phi[2], z, fmaf(P.phi[1], y, P.phi[0] * x));
  float const ux = fmaf(P.phi[0], d, fmaf(-P.theta2, x, w * P.phi_x_rho[0]));
  float const uy = fmaf(P.phi[1], d, fmaf(-P.theta2, y, w * P.phi_x_rho[1]));
  float const uz = fmaf(P.phi[2], d, fmaf(-P.theta2, z, w * P.phi_x_rho[2]));
  float const vx = fmaf(P.phi[1], z, fmaf(-P.phi[2], y, w * P.rho_perp[0]));
  float const vy = fmaf(P.phi[2], x, fmaf(-P.phi[0], z, w * P.rho_perp[1]));
  float const vz = fmaf(P.phi[0], y, fmaf(-P.phi[1], x, w * P.rho_perp[2]));
  float const sw = s * w;
  return make_float3(fmaf(C, ux, fmaf(S, vx, sw * P.rho_par[0])), fmaf(C, uy, fmaf(S, vy, sw * P.rho_par[1])),
                     fmaf(C, uz, fmaf(S, vz, sw * P.rho_par[2])));
}

template <int MODE>
__device__ __forceinline__ float4 DeskewPoint(float4 p, const kmc_b200_frame_params& P) {
  float s;
  if constexpr (MODE == KMC_B200_TIME_FROM_AZIMUTH) {
    // frac = (pi - atan2(y,x)) / 2pi = 0.5 - turns ;  s = frac - x_req = c0 - turns
    s = P.c0 - Atan2Turns(p.y, p.x);
  } else {
    s = p.w - P.x_req;
  }
  float3 const delta = DeskewDelta(p.x, p.y, p.z, s, P);
  return make_float4(p.x + delta.x, p.y + delta.y, p.z + delta.z, p.w);  // one rounding at the magnitude of p
}

// ---------------------------------------------------------------------------------------------------------------
// frame table lookup (batch kernels)
// ---------------------------------------------------------------------------------------------------------------
// largest f with offsets[f] <= p (offsets[n_frames] > p is guaranteed by the caller)
__device__ __forceinline__ int LocateFrame(const int64_t* __restrict__ offsets, int n_frames, int64_t p, double frames_per_point) {
  int guess = static_cast<int>(static_cast<double>(p) * frames_per_point);
  guess = guess < 0 ? 0 : (guess > n_frames - 1 ? n_frames - 1 : guess);
  if (__ldg(offsets + guess) <= p && p < __ldg(offsets + guess + 1)) return guess;
  int lo = 0, hi = n_frames;
  while (lo < hi) {
    int const mid = (lo + hi + 1) >> 1;
    if (__ldg(offsets + mid) <= p) lo = mid; else hi = mid - 1;
  }
  return lo;
}

}  // namespace kmc_b200::dev
