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
