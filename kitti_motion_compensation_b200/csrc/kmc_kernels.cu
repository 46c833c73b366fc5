// kmc_kernels.cu — hand-written sm_100a kernels of the LiDAR deskew path.
//
// One fused streaming kernel replaces, per point, the reference's
//   GetPseudoTimeStamps            timestamp_mocking.cpp:46-63   (azimuth -> fraction of scan)
//   TrajectoryInterpolator         trajectory_interpolation.cpp:31-51 (fraction -> interpolated pose, relative pose)
//   lie::Exp                       lie_algebra.cpp:22-35,51-65,83-92  (Rodrigues + left Jacobian)
//   MotionCompensatePoint          motion_compensation.cpp:9-14  (rigid transform apply)
// Everything that does not depend on the point (Log(T_start^-1 T_end) and the constants derived from it) is computed
// once per frame on the host in double (kmc_host_math.cpp) and reaches the kernel as a 64-byte record.
//
// Math.  With xi = [rho; phi] the twist of the whole scan and s = x_i - x_req the signed fraction of
// the scan between the requested time and the point's capture time, the reference's
//   correction = GetPoseAtTime(t_req)^-1 GetPoseAtTime(t_i) = Exp(x_req xi)^-1 Exp(x_i xi) = Exp(s xi)
// and  p' = R(s phi) p + J(s phi) s rho.  Written as a DISPLACEMENT so that fp32 keeps 1e-6 m at 120 m range:
//   S = sin(s th)/th,  C = (1 - cos(s th))/th^2
//   delta = C (phi (phi.p) - th^2 p + phi x rho) + S (phi x p + rho_perp) + s rho_par
//   p'    = p + delta                                  (one rounding at the magnitude of p)
// S and C come from a power series in (s th)^2 (no division, exact limit at th -> 0, which is the reference's own
// golden test), or from half-angle polynomials valid to pi when a scan rotates by more than 1 rad.
//
// Memory: the N x 4 float32 "x y z i" array is read once and written once (32 B/point), 128-bit or 256-bit
// (sm_100 LDG.256/STG.256) accesses, fully coalesced, one CTA of 128 threads per work item — the hardware scheduler balances
// SMs of different speed better than a persistent grid does (PickConfig).  HBM-bandwidth bound; no shared memory, no tensor
// cores (nothing to contract).
#include "kmc_internal.hpp"
#include "kmc_kernels.cuh"
#include "kmc_point_math.cuh"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>

namespace kmc_b200::dev {
namespace {

std::atomic<uint64_t> g_launches{0};

// ---------------------------------------------------------------------------------------------------------------
// memory access helpers
// ---------------------------------------------------------------------------------------------------------------
struct alignas(32) Point2 {
  float4 a, b;
};

template <int HINT>
__device__ __forceinline__ float4 LoadPoint(const float4* p) {
  float4 v;
  if constexpr (HINT == 1) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
  } else if constexpr (HINT == 2) {
    asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  } else {
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  }
  return v;
}

template <int HINT>
__device__ __forceinline__ void StorePoint(float4* p, float4 v) {
  if constexpr (HINT == 1) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  } else if constexpr (HINT == 2) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  } else {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
}

// 256-bit accesses: two consecutive points per instruction (PTX ISA 8.8, sm_100+: SASS LDG.E.256 / STG.E.256).
template <int HINT>
__device__ __forceinline__ Point2 LoadPoint2(const float4* p) {
  Point2 v;
  if constexpr (HINT == 1) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w)
                 : "l"(p));
  } else if constexpr (HINT == 2) {
    asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w)
                 : "l"(p));
  } else if constexpr (HINT == 3 || HINT == 5 || HINT == 6) {  // experiment: L2 eviction priority (SASS LDG.E.EFL2.256)
    asm volatile("ld.global.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w)
                 : "l"(p));
  } else {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w)
                 : "l"(p));
  }
  return v;
}

template <int HINT>
__device__ __forceinline__ void StorePoint2(float4* p, const Point2& v) {
  if constexpr (HINT == 1) {
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v.a.x), "f"(v.a.y),
                 "f"(v.a.z), "f"(v.a.w), "f"(v.b.x), "f"(v.b.y), "f"(v.b.z), "f"(v.b.w)
                 : "memory");
  } else if constexpr (HINT == 2) {
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v.a.x), "f"(v.a.y), "f"(v.a.z),
                 "f"(v.a.w), "f"(v.b.x), "f"(v.b.y), "f"(v.b.z), "f"(v.b.w)
                 : "memory");
  } else if constexpr (HINT == 4 || HINT == 6) {  // STG.E.EFL2.256
    asm volatile("st.global.L2::evict_first.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v.a.x), "f"(v.a.y), "f"(v.a.z),
                 "f"(v.a.w), "f"(v.b.x), "f"(v.b.y), "f"(v.b.z), "f"(v.b.w)
                 : "memory");
  } else if constexpr (HINT == 5) {  // STG.E.ELL2.256
    asm volatile("st.global.L2::evict_last.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v.a.x), "f"(v.a.y), "f"(v.a.z),
                 "f"(v.a.w), "f"(v.b.x), "f"(v.b.y), "f"(v.b.z), "f"(v.b.w)
                 : "memory");
  } else {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v.a.x), "f"(v.a.y), "f"(v.a.z),
                 "f"(v.a.w), "f"(v.b.x), "f"(v.b.y), "f"(v.b.z), "f"(v.b.w)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// A contiguous range [a, b) of points that all belong to one frame, processed by the whole CTA.
// ---------------------------------------------------------------------------------------------------------------
template <int MODE, int VEC, int UNROLL, int HINT, int BLOCK>
__device__ __forceinline__ void ProcessRange(const float4* __restrict__ in, float4* __restrict__ out, int64_t a, int64_t b,
                                             const kmc_b200_frame_params& P) {
  int const tid = threadIdx.x;
  if constexpr (VEC == 2) {
    if (a & 1) {  // 256-bit accesses need an even point index (the base pointer is 32-byte aligned)
      if (tid == 0 && a < b) StorePoint<HINT>(out + a, DeskewPoint<MODE>(LoadPoint<HINT>(in + a), P));
      a += 1;
    }
  }
  constexpr int64_t kTile = static_cast<int64_t>(BLOCK) * UNROLL * VEC;
  int64_t base = a;
  for (; base + kTile <= b; base += kTile) {
    if constexpr (VEC == 2) {
      Point2 v[UNROLL];
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) v[j] = LoadPoint2<HINT>(in + base + 2 * (j * BLOCK + tid));
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        Point2 r;
        r.a = DeskewPoint<MODE>(v[j].a, P);
        r.b = DeskewPoint<MODE>(v[j].b, P);
        StorePoint2<HINT>(out + base + 2 * (j * BLOCK + tid), r);
      }
    } else {
      float4 v[UNROLL];
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) v[j] = LoadPoint<HINT>(in + base + j * BLOCK + tid);
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) StorePoint<HINT>(out + base + j * BLOCK + tid, DeskewPoint<MODE>(v[j], P));
    }
  }
  for (int64_t i = base + tid; i < b; i += BLOCK) StorePoint<HINT>(out + i, DeskewPoint<MODE>(LoadPoint<HINT>(in + i), P));
}

// ---------------------------------------------------------------------------------------------------------------
// single frame: per-frame constants as a __grid_constant__ parameter (constant bank, broadcast to every lane for free)
// ---------------------------------------------------------------------------------------------------------------
template <int MODE, int VEC, int UNROLL, int HINT, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
    DeskewFrameKernel(const float4* __restrict__ in, float4* __restrict__ out, int64_t n, int64_t item_points,
                      const __grid_constant__ kmc_b200_frame_params P) {
  // Programmatic dependent launch (LaunchFrameT): let the next kernel of the stream be scheduled while this one drains, and
  // do not touch memory before the previous one has completed and flushed.  Both are no-ops for an ordinary launch.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int64_t const n_items = (n + item_points - 1) / item_points;
  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    int64_t const p0 = item * item_points;
    int64_t const p1 = (p0 + item_points < n) ? (p0 + item_points) : n;
    ProcessRange<MODE, VEC, UNROLL, HINT, BLOCK>(in, out, p0, p1, P);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// batch of frames stored back to back: per-frame records in a global table (10 000 frames x 64 B do not fit the 64 KB
// constant bank).  A work item is cut at frame boundaries; each piece is processed with its frame's record, which one
// coalesced 64-byte load per warp fetches (lanes 0-15, one float each) and __shfl_sync broadcasts to the warp.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ kmc_b200_frame_params LoadParamsWarpBroadcast(const kmc_b200_frame_params* __restrict__ table,
                                                                        int frame) {
  int const lane = threadIdx.x & 31;
  const float* rec = reinterpret_cast<const float*>(table + frame);
  float const mine = __ldg(rec + (lane & 15));
  kmc_b200_frame_params P;
  float* dst = reinterpret_cast<float*>(&P);
#pragma unroll
  for (int k = 0; k < 16; ++k) dst[k] = __shfl_sync(0xffffffffu, mine, k);
  return P;
}

template <int MODE, int VEC, int UNROLL, int HINT, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
    DeskewBatchKernel(const float4* __restrict__ in, float4* __restrict__ out, const int64_t* __restrict__ offsets,
                      const kmc_b200_frame_params* __restrict__ table, int n_frames, int64_t n, int64_t item_points,
                      int64_t point_base, double frames_per_point) {
  int64_t const n_items = (n + item_points - 1) / item_points;
  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    int64_t p0 = item * item_points;
    int64_t const p1 = (p0 + item_points < n) ? (p0 + item_points) : n;
    // in/out address the chunk [point_base, point_base + n) of the batch; offsets count from the start of the batch
    int f = LocateFrame(offsets, n_frames, p0 + point_base, frames_per_point);
    // f < n_frames: a caller whose n exceeds offsets[n_frames] (a broken precondition, the tables live on the device and
    // cannot be checked by the host) leaves the surplus points untouched instead of walking off the tables
    while (p0 < p1 && f < n_frames) {
      int64_t const frame_end = __ldg(offsets + f + 1) - point_base;
      if (frame_end <= p0) {  // empty frame
        ++f;
        continue;
      }
      int64_t const seg_end = frame_end < p1 ? frame_end : p1;
      kmc_b200_frame_params const P = LoadParamsWarpBroadcast(table, f);
      ProcessRange<MODE, VEC, UNROLL, HINT, BLOCK>(in, out, p0, seg_end, P);
      p0 = seg_end;
      ++f;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Projection onto a rectified camera (camera_model.cpp:5-36,38-95 without the drawing), optionally fused behind the
// deskew: one pass reads a point, (deskews it,) writes the (deskewed) point and its pixel record.
// Streaming map, HBM bound: 32 B/point (project only, or deskew+project without the cloud output), 48 B/point (both).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ProjectPoint(float4 p, const kmc_b200_camera_params& K) {
  float const xr = fmaf(K.rect[0], p.x, fmaf(K.rect[1], p.y, fmaf(K.rect[2], p.z, K.rect[3])));
  float const yr = fmaf(K.rect[4], p.x, fmaf(K.rect[5], p.y, fmaf(K.rect[6], p.z, K.rect[7])));
  float const zr = fmaf(K.rect[8], p.x, fmaf(K.rect[9], p.y, fmaf(K.rect[10], p.z, K.rect[11])));
  float const pu = fmaf(K.pix[0], p.x, fmaf(K.pix[1], p.y, fmaf(K.pix[2], p.z, K.pix[3])));
  float const pv = fmaf(K.pix[4], p.x, fmaf(K.pix[5], p.y, fmaf(K.pix[6], p.z, K.pix[7])));
  float const pw = fmaf(K.pix[8], p.x, fmaf(K.pix[9], p.y, fmaf(K.pix[10], p.z, K.pix[11])));
  float const inv = 1.0f / pw;  // IEEE division: pixel coordinates are compared at the 1e-2 px level
  bool const culled = (zr < K.min_depth) || (zr > K.max_range) || (yr > K.max_below);
  (void)xr;
  return make_float4(pu * inv, pv * inv, zr, culled ? -1.0f : zr * K.color_gain);
}

template <bool DESKEW, bool WRITE_CLOUD, int MODE, bool VEC2, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
    ProjectFrameKernel(const float4* __restrict__ in, float4* __restrict__ cloud_out, float4* __restrict__ pix_out, int64_t n,
                       const __grid_constant__ kmc_b200_frame_params P, const __grid_constant__ kmc_b200_camera_params K) {
  if constexpr (!VEC2) {  // buffers only 16-byte aligned: one point per 128-bit access
    int64_t const stride1 = static_cast<int64_t>(gridDim.x) * BLOCK;
    for (int64_t j = static_cast<int64_t>(blockIdx.x) * BLOCK + threadIdx.x; j < n; j += stride1) {
      float4 p = LoadPoint<0>(in + j);
      if constexpr (DESKEW) {
        p = DeskewPoint<MODE>(p, P);
        if constexpr (WRITE_CLOUD) StorePoint<0>(cloud_out + j, p);
      }
      StorePoint<0>(pix_out + j, ProjectPoint(p, K));
    }
    return;
  }
  int64_t const stride = static_cast<int64_t>(gridDim.x) * BLOCK * 2;
  int64_t i = (static_cast<int64_t>(blockIdx.x) * BLOCK + threadIdx.x) * 2;
  for (; i + 1 < n; i += stride) {  // two points per 256-bit access
    Point2 v = LoadPoint2<0>(in + i);
    if constexpr (DESKEW) {
      v.a = DeskewPoint<MODE>(v.a, P);
      v.b = DeskewPoint<MODE>(v.b, P);
      if constexpr (WRITE_CLOUD) StorePoint2<0>(cloud_out + i, v);
    }
    Point2 r;
    r.a = ProjectPoint(v.a, K);
    r.b = ProjectPoint(v.b, K);
    StorePoint2<0>(pix_out + i, r);
  }
  if (i < n) {  // odd tail
    float4 p = LoadPoint<0>(in + i);
    if constexpr (DESKEW) {
      p = DeskewPoint<MODE>(p, P);
      if constexpr (WRITE_CLOUD) StorePoint<0>(cloud_out + i, p);
    }
    StorePoint<0>(pix_out + i, ProjectPoint(p, K));
  }
}

// All four rectified cameras in one pass (camera_model.cpp:85-92 projects the same cloud four times): the cloud is read
// once, four pixel records are written.  16 (+16 with the deskewed cloud) + 4 x 16 B/point instead of 4 x 32.
struct Cameras4 {
  kmc_b200_camera_params cam[4];
};
struct PixelPlanes4 {
  float4* plane[4];
};

template <bool DESKEW, bool WRITE_CLOUD, int MODE, bool VEC2, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
    ProjectFrame4Kernel(const float4* __restrict__ in, float4* __restrict__ cloud_out, PixelPlanes4 const pix_out, int64_t n,
                        const __grid_constant__ kmc_b200_frame_params P, const __grid_constant__ Cameras4 K) {
  if constexpr (!VEC2) {  // buffers only 16-byte aligned: one point per 128-bit access
    int64_t const stride1 = static_cast<int64_t>(gridDim.x) * BLOCK;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * BLOCK + threadIdx.x; i < n; i += stride1) {
      float4 p = LoadPoint<0>(in + i);
      if constexpr (DESKEW) {
        p = DeskewPoint<MODE>(p, P);
        if constexpr (WRITE_CLOUD) StorePoint<0>(cloud_out + i, p);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) StorePoint<0>(pix_out.plane[c] + i, ProjectPoint(p, K.cam[c]));
    }
    return;
  }
  int64_t const stride = static_cast<int64_t>(gridDim.x) * BLOCK * 2;
  int64_t i = (static_cast<int64_t>(blockIdx.x) * BLOCK + threadIdx.x) * 2;
  for (; i + 1 < n; i += stride) {  // two points per 256-bit access
    Point2 v = LoadPoint2<0>(in + i);
    if constexpr (DESKEW) {
      v.a = DeskewPoint<MODE>(v.a, P);
      v.b = DeskewPoint<MODE>(v.b, P);
      if constexpr (WRITE_CLOUD) StorePoint2<0>(cloud_out + i, v);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      Point2 r;
      r.a = ProjectPoint(v.a, K.cam[c]);
      r.b = ProjectPoint(v.b, K.cam[c]);
      StorePoint2<0>(pix_out.plane[c] + i, r);
    }
  }
  if (i < n) {  // odd tail
    float4 p = LoadPoint<0>(in + i);
    if constexpr (DESKEW) {
      p = DeskewPoint<MODE>(p, P);
      if constexpr (WRITE_CLOUD) StorePoint<0>(cloud_out + i, p);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) StorePoint<0>(pix_out.plane[c] + i, ProjectPoint(p, K.cam[c]));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Deskew + projection over a BATCH of frames — the loop body of GenerateProjectionVisualizationOfRun (handlers.cpp:67-92:
// per frame one MotionCompensateFrame and four ProjectPointcloudOnImage) for a whole run in one launch.  Frames back to
// back as for DeskewBatchKernel (global offset / record tables, work items cut at frame boundaries, one coalesced
// 64-byte record fetch per warp and piece); the cameras are the same for every frame of a run and ride in the constant
// bank.  16 B/point read, 16 B/point written per camera (+16 with the deskewed cloud).
// ---------------------------------------------------------------------------------------------------------------
template <int MODE, int NCAM, bool WRITE_CLOUD, int VEC, int BLOCK>
__device__ __forceinline__ void ProjectRange(const float4* __restrict__ in, float4* __restrict__ cloud_out, const PixelPlanes4& pix,
                                             int64_t a, int64_t b, const kmc_b200_frame_params& P, const Cameras4& K) {
  int const tid = threadIdx.x;
  auto one = [&](int64_t i) {
    float4 const p = DeskewPoint<MODE>(LoadPoint<0>(in + i), P);
    if constexpr (WRITE_CLOUD) StorePoint<0>(cloud_out + i, p);
#pragma unroll
    for (int c = 0; c < NCAM; ++c) StorePoint<0>(pix.plane[c] + i, ProjectPoint(p, K.cam[c]));
  };
  if constexpr (VEC == 2) {
    if (a & 1) {  // 256-bit accesses need an even point index
      if (tid == 0 && a < b) one(a);
      a += 1;
    }
  }
  constexpr int64_t kTile = static_cast<int64_t>(BLOCK) * VEC;
  int64_t base = a;
  for (; base + kTile <= b; base += kTile) {
    if constexpr (VEC == 2) {
      int64_t const i = base + 2 * tid;
      Point2 v = LoadPoint2<0>(in + i);
      v.a = DeskewPoint<MODE>(v.a, P);
      v.b = DeskewPoint<MODE>(v.b, P);
      if constexpr (WRITE_CLOUD) StorePoint2<0>(cloud_out + i, v);
#pragma unroll
      for (int c = 0; c < NCAM; ++c) {
        Point2 r;
        r.a = ProjectPoint(v.a, K.cam[c]);
        r.b = ProjectPoint(v.b, K.cam[c]);
        StorePoint2<0>(pix.plane[c] + i, r);
      }
    } else {
      one(base + tid);
    }
  }
  for (int64_t i = base + tid; i < b; i += BLOCK) one(i);
}

template <int MODE, int NCAM, bool WRITE_CLOUD, int VEC, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
    DeskewProjectBatchKernel(const float4* __restrict__ in, float4* __restrict__ cloud_out, PixelPlanes4 const pix,
                             const int64_t* __restrict__ offsets, const kmc_b200_frame_params* __restrict__ table, int n_frames,
                             int64_t n, int64_t item_points, double frames_per_point, const __grid_constant__ Cameras4 K) {
  int64_t const n_items = (n + item_points - 1) / item_points;
  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    int64_t p0 = item * item_points;
    int64_t const p1 = (p0 + item_points < n) ? (p0 + item_points) : n;
    int f = LocateFrame(offsets, n_frames, p0, frames_per_point);
    while (p0 < p1 && f < n_frames) {
      int64_t const frame_end = __ldg(offsets + f + 1);
      if (frame_end <= p0) {  // empty frame
        ++f;
        continue;
      }
      int64_t const seg_end = frame_end < p1 ? frame_end : p1;
      kmc_b200_frame_params const P = LoadParamsWarpBroadcast(table, f);
      ProjectRange<MODE, NCAM, WRITE_CLOUD, VEC, BLOCK>(in, cloud_out, pix, p0, seg_end, P, K);
      p0 = seg_end;
      ++f;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The reference's own memory layout (motion_compensation.cpp:16-28): cloud and result are COLUMN-major N x 4 doubles
// (Eigen::MatrixX4d), per-point stamps a separate double vector.  Used by the C++ mirror of MotionCompensateFrame so that
// no host-side layout conversion is needed.  The displacement is computed in fp32 from the rounded coordinates and the
// stamp fraction (formed in double), then added to the DOUBLE coordinate: no float32 output rounding, the result is
// within ~2.5e-7 of the displacement of the reference's.  72 B/point (32 + 8 read, 32 written), columns coalesced.
// flags: bit 0 = a stamp outside [t1, t2] (the reference asserts), bit 1 = a 4th-column entry that is not 1 (honoured as
// the reference does: R p + t w, w passed through — motion_compensation.cpp:13).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 LoadPairF64(const double* __restrict__ p, bool aligned16) {
  if (aligned16) return __ldg(reinterpret_cast<const double2*>(p));
  return make_double2(__ldg(p), __ldg(p + 1));
}
__device__ __forceinline__ void StorePairF64(double* __restrict__ p, bool aligned16, double a, double b) {
  if (aligned16) {
    // as PTX: written as a C++ double2 store, the compiler merges both branches into one STG.128 and faults on odd columns
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
  } else {
    p[0] = a;
    p[1] = b;
  }
}

// One column-major N x 4 double matrix `c` with its stamps `ts` -> `o`, the points lane, lane + lanes, ... (PAIR: the point
// PAIRS).  PAIR = two neighbouring points per thread with 128-bit accesses on every column that starts 16-byte aligned (all of
// them when N is even; x and z always; a misaligned column falls back to two 64-bit accesses, uniformly per frame): twice the
// bytes in flight per thread at the same residency, half the memory instructions — 6.6-6.8 TB/s against 5.9-6.0 for one point
// per thread (profiles/r02_sweep_f64_batch_pair_defaults.log).  Returns the flag bits seen by this thread.
template <bool PAIR, typename Params>
__device__ __forceinline__ int DeskewF64Columns(const double* __restrict__ c, const double* __restrict__ ts, double* __restrict__ o,
                                                int64_t n, int64_t lane, int64_t lanes, double t1, double t2, double duration,
                                                double x_req, const Params& P) {
  int bad = 0;
  auto one = [&](int64_t i) {
    double const x = __ldg(c + i), y = __ldg(c + n + i), z = __ldg(c + 2 * n + i), w = __ldg(c + 3 * n + i);
    double const t = __ldg(ts + i);
    if (!(t >= t1 && t <= t2)) bad |= 1;
    if (w != 1.0) bad |= 2;
    float const s = static_cast<float>((t - t1) / duration - x_req);  // FractionOfTrajectory, trajectory_interpolation.cpp:49-51
    float3 const d = DeskewDeltaW(static_cast<float>(x), static_cast<float>(y), static_cast<float>(z), static_cast<float>(w), s, P);
    o[i] = x + static_cast<double>(d.x);
    o[n + i] = y + static_cast<double>(d.y);
    o[2 * n + i] = z + static_cast<double>(d.z);
    o[3 * n + i] = w;
  };
  if constexpr (!PAIR) {
    for (int64_t i = lane; i < n; i += lanes) one(i);
  } else {
    // i is even, so a column is 16-byte aligned at i exactly when its first element is
    auto const al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    bool const ax = al(c) && al(o), ay = al(c + n) && al(o + n), az = al(c + 2 * n) && al(o + 2 * n), aw = al(c + 3 * n) && al(o + 3 * n);
    bool const at = al(ts);
    int64_t i = 2 * lane;
    for (; i + 1 < n; i += 2 * lanes) {
      double2 const x = LoadPairF64(c + i, ax), y = LoadPairF64(c + n + i, ay), z = LoadPairF64(c + 2 * n + i, az);
      double2 const w = LoadPairF64(c + 3 * n + i, aw), t = LoadPairF64(ts + i, at);
      if (!(t.x >= t1 && t.x <= t2) || !(t.y >= t1 && t.y <= t2)) bad |= 1;
      if (w.x != 1.0 || w.y != 1.0) bad |= 2;
      float const s0 = static_cast<float>((t.x - t1) / duration - x_req), s1 = static_cast<float>((t.y - t1) / duration - x_req);
      float3 const d0 = DeskewDeltaW(static_cast<float>(x.x), static_cast<float>(y.x), static_cast<float>(z.x), static_cast<float>(w.x), s0, P);
      float3 const d1 = DeskewDeltaW(static_cast<float>(x.y), static_cast<float>(y.y), static_cast<float>(z.y), static_cast<float>(w.y), s1, P);
      StorePairF64(o + i, ax, x.x + static_cast<double>(d0.x), x.y + static_cast<double>(d1.x));
      StorePairF64(o + n + i, ay, y.x + static_cast<double>(d0.y), y.y + static_cast<double>(d1.y));
      StorePairF64(o + 2 * n + i, az, z.x + static_cast<double>(d0.z), z.y + static_cast<double>(d1.z));
      StorePairF64(o + 3 * n + i, aw, w.x, w.y);
    }
    if (i < n) one(i);  // the odd last point
  }
  return bad;
}

template <bool PAIR>
__global__ void __launch_bounds__(kBlockThreads, PAIR ? 4 : 1)
    DeskewCloudF64Kernel(const double* __restrict__ cloud, const double* __restrict__ stamps, double* __restrict__ out, int64_t n,
                         double t1, double t2, double x_req, const __grid_constant__ kmc_b200_frame_params P,
                         int* __restrict__ flags) {
  int const bad = DeskewF64Columns<PAIR>(cloud, stamps, out, n, static_cast<int64_t>(blockIdx.x) * kBlockThreads + threadIdx.x,
                                         static_cast<int64_t>(gridDim.x) * kBlockThreads, t1, t2, t2 - t1, x_req, P);
  if (bad) atomicOr(flags, bad);
}

// A BATCH of frames in the reference's layout: frame f owns the 4 N_f doubles at cloud + 4 offsets[f] (its own column-major
// N_f x 4 matrix: x | y | z | w columns), the N_f stamps at stamps + offsets[f], and the same block of `out`.  Per-frame
// records in the global table as for DeskewBatchKernel, per-frame times (t_start, t_end, t_req) in a table of three
// doubles, per-frame flags (bit 0 stamp out of range, bit 1 some w != 1).  72 B/point.
// Grid: blockIdx.y (+ z) is the FRAME, blockIdx.x one of gridDim.x CTAs that stride over the frame's tiles together — inside
// a frame this is exactly DeskewCloudF64Kernel's access pattern (neighbouring CTAs read neighbouring 2 KB pieces of each of
// the nine column streams at the same time, which is what keeps DRAM rows open), no frame search, and the hardware
// scheduler hands out (frame, lane) pairs in frame order.  The first version cut the flat point range into 4096-point items
// as DeskewBatchKernel does; with nine streams per item that gave 4.3-5.5 TB/s against 6.6 for the single-frame kernel
// (profiles/r02_sweep_f64_batch.log).  The frame's 64-byte record sits in shared memory: in registers it costs 16 of them on top of
// nine 64-bit column pointers.  One point per thread this shape ran at 5.97-6.06 TB/s (64 registers, 1024 threads per SM, 40 bytes
// in flight per thread); two points per thread (DeskewF64Columns<true>, still 64 registers) 6.6-6.8 TB/s for every frame shape
// from one 39 M-point frame to 3000 frames of 13 001 points (profiles/r02_sweep_f64_batch_pair_defaults.log).  Measured and not adopted:
// loading the next tile before computing the current one — slower, the extra live values spill.
template <int BLOCK, bool PAIR, int MIN_CTAS>
__global__ void __launch_bounds__(BLOCK, MIN_CTAS)
    DeskewCloudF64BatchKernel(const double* __restrict__ cloud, const double* __restrict__ stamps, double* __restrict__ out,
                              const int64_t* __restrict__ offsets, const kmc_b200_frame_params* __restrict__ table,
                              const double* __restrict__ times, int n_frames, int* __restrict__ flags) {
  __shared__ kmc_b200_frame_params record;
  constexpr int PER = PAIR ? 2 : 1;
  int const f = static_cast<int>(blockIdx.y) + static_cast<int>(blockIdx.z) * static_cast<int>(gridDim.y);
  if (f >= n_frames) return;
  int64_t const frame_begin = __ldg(offsets + f);
  int64_t const nf = __ldg(offsets + f + 1) - frame_begin;
  if (static_cast<int64_t>(blockIdx.x) * BLOCK * PER >= nf) return;  // more CTAs per frame than this frame has tiles
  if (threadIdx.x < 16) reinterpret_cast<float*>(&record)[threadIdx.x] = __ldg(reinterpret_cast<const float*>(table + f) + threadIdx.x);
  double const t1 = __ldg(times + 3 * f), t2 = __ldg(times + 3 * f + 1), t_req = __ldg(times + 3 * f + 2);
  __syncthreads();
  double const duration = t2 - t1;
  double const x_req = (t_req - t1) / duration;
  const double* const c = cloud + 4 * frame_begin;
  double* const o = out + 4 * frame_begin;
  const double* const ts = stamps + frame_begin;
  int const bad = DeskewF64Columns<PAIR>(c, ts, o, nf, static_cast<int64_t>(blockIdx.x) * BLOCK + threadIdx.x,
                                         static_cast<int64_t>(gridDim.x) * BLOCK, t1, t2, duration, x_req, record);
  if (bad) atomicOr(flags + f, bad);
}

// ---------------------------------------------------------------------------------------------------------------
// Narrow transport of the reference-layout HOST path (kmc_b200_deskew_cloud_f64_host).  The link, not the GPU, bounds a
// host call, so the double cloud never crosses it: the host rounds x, y, z to float, forms each point's signed
// trajectory fraction s in double and rounds it, and ships four float columns (16 B/point up, a fifth column only when a
// w != 1 is present); the kernel returns the three displacement columns (12 B/point down) and the host adds them to
// its own doubles.  Same arithmetic as DeskewCloudF64Kernel operation by operation, hence the same bits.
// Columns hold stride4 float4s each (chunk length rounded up to 4 points; the pad lanes carry zeros).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDeltaBlockThreads = 128;

template <bool HAS_W>
__global__ void __launch_bounds__(kDeltaBlockThreads)
    DeskewDeltaColumnsKernel(const float4* __restrict__ in, float4* __restrict__ out, int64_t stride4,
                             const __grid_constant__ kmc_b200_frame_params P) {
  int64_t const step = static_cast<int64_t>(gridDim.x) * kDeltaBlockThreads;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kDeltaBlockThreads + threadIdx.x; i < stride4; i += step) {
    float4 const x = LoadPoint<0>(in + i), y = LoadPoint<0>(in + stride4 + i), z = LoadPoint<0>(in + 2 * stride4 + i);
    float4 const s = LoadPoint<0>(in + 3 * stride4 + i);
    float3 d0, d1, d2, d3;
    if constexpr (HAS_W) {
      float4 const w = LoadPoint<0>(in + 4 * stride4 + i);
      d0 = DeskewDeltaW(x.x, y.x, z.x, w.x, s.x, P);
      d1 = DeskewDeltaW(x.y, y.y, z.y, w.y, s.y, P);
      d2 = DeskewDeltaW(x.z, y.z, z.z, w.z, s.z, P);
      d3 = DeskewDeltaW(x.w, y.w, z.w, w.w, s.w, P);
    } else {
      d0 = DeskewDelta(x.x, y.x, z.x, s.x, P);
      d1 = DeskewDelta(x.y, y.y, z.y, s.y, P);
      d2 = DeskewDelta(x.z, y.z, z.z, s.z, P);
      d3 = DeskewDelta(x.w, y.w, z.w, s.w, P);
    }
    StorePoint<0>(out + i, make_float4(d0.x, d1.x, d2.x, d3.x));
    StorePoint<0>(out + stride4 + i, make_float4(d0.y, d1.y, d2.y, d3.y));
    StorePoint<0>(out + 2 * stride4 + i, make_float4(d0.z, d1.z, d2.z, d3.z));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GetPseudoTimeStamps (timestamp_mocking.cpp:56-63) in double, for callers that want the stamps themselves.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kStampBlockThreads = 128;

template <bool VEC2>
__global__ void __launch_bounds__(kStampBlockThreads)
    PseudoTimeStampsKernel(const float4* __restrict__ in, double* __restrict__ stamps, int64_t n, double start, double duration) {
  if constexpr (!VEC2) {
    int64_t const stride1 = static_cast<int64_t>(gridDim.x) * kStampBlockThreads;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kStampBlockThreads + threadIdx.x; i < n; i += stride1) {
      float4 const p = LoadPoint<0>(in + i);
      stamps[i] = fma(FractionOfScanF64(p.y, p.x), duration, start);
    }
    return;
  }
  // two points per 256-bit load, two stamps per 128-bit store
  int64_t const stride = static_cast<int64_t>(gridDim.x) * kStampBlockThreads * 2;
  int64_t i = (static_cast<int64_t>(blockIdx.x) * kStampBlockThreads + threadIdx.x) * 2;
  for (; i + 1 < n; i += stride) {
    Point2 const v = LoadPoint2<0>(in + i);
    double2 out;
    out.x = fma(FractionOfScanF64(v.a.y, v.a.x), duration, start);
    out.y = fma(FractionOfScanF64(v.b.y, v.b.x), duration, start);
    *reinterpret_cast<double2*>(stamps + i) = out;
  }
  if (i < n) {
    float4 const p = LoadPoint<0>(in + i);
    stamps[i] = fma(FractionOfScanF64(p.y, p.x), duration, start);
  }
}

// VEC2: two consecutive points per thread through 128-bit accesses (columns 16-byte aligned).  The kernel is latency bound —
// ncu: long_scoreboard 11.8 of 17 stalled warps per issue, DRAM 59 % busy with 16 bytes in flight per thread — so doubling
// the bytes in flight per thread is what moves it, not fewer instructions (an fp32 octant-logic variant was slower).
template <bool VEC2>
__global__ void __launch_bounds__(kBlockThreads)
    PseudoTimeStampsXyKernel(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ stamps, int64_t n,
                             double start, double duration) {
  if constexpr (VEC2) {
    int64_t const stride = static_cast<int64_t>(gridDim.x) * kBlockThreads * 2;
    int64_t i = (static_cast<int64_t>(blockIdx.x) * kBlockThreads + threadIdx.x) * 2;
    for (; i + 1 < n; i += stride) {
      double2 const xx = __ldg(reinterpret_cast<const double2*>(x + i));
      double2 const yy = __ldg(reinterpret_cast<const double2*>(y + i));
      double2 out;
      out.x = fma(0.5 - Atan2TurnsF64(yy.x, xx.x), duration, start);
      out.y = fma(0.5 - Atan2TurnsF64(yy.y, xx.y), duration, start);
      *reinterpret_cast<double2*>(stamps + i) = out;
    }
    if (i < n) stamps[i] = fma(0.5 - Atan2TurnsF64(__ldg(y + i), __ldg(x + i)), duration, start);  // odd tail
  } else {
    int64_t const stride = static_cast<int64_t>(gridDim.x) * kBlockThreads;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kBlockThreads + threadIdx.x; i < n; i += stride) {
      stamps[i] = fma(0.5 - Atan2TurnsF64(__ldg(y + i), __ldg(x + i)), duration, start);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// seeded synthetic spinning-LiDAR scans, generated in HBM
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t SplitMix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(kBlockThreads)
    SynthScansKernel(float4* __restrict__ out, int64_t points_per_scan, int64_t n_total, int n_rings, uint64_t seed,
                     int64_t first_scan_index) {
  int64_t const stride = static_cast<int64_t>(gridDim.x) * kBlockThreads;
  int64_t const steps = (points_per_scan + n_rings - 1) / n_rings;  // azimuth steps per ring
  float const el_top = (n_rings == 64) ? 2.0f : 15.0f;              // HDL-64E: +2.0 .. -24.8 deg; dense: +15 .. -25 deg
  float const el_bot = (n_rings == 64) ? -24.8f : -25.0f;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * kBlockThreads + threadIdx.x; g < n_total; g += stride) {
    int64_t const scan = g / points_per_scan;
    int64_t const i = g - scan * points_per_scan;
    int64_t const ring = i / steps;
    int64_t const step = i - ring * steps;
    uint64_t const h0 = SplitMix64(SplitMix64(seed + static_cast<uint64_t>(first_scan_index + scan)) ^ static_cast<uint64_t>(i));
    uint64_t const h1 = SplitMix64(h0);
    float const u_range = static_cast<float>(h0 >> 40) * (1.0f / 16777216.0f);            // [0,1)
    float const u_jit = static_cast<float>((h0 >> 16) & 0xFFFFFF) * (1.0f / 16777216.0f);  // [0,1)
    float const inten = static_cast<float>(h1 % 100u) * 0.01f;
    float const range = 2.0f * expf(u_range * 4.0943445622f);  // log-uniform [2, 120)
    float const el = (el_top + (el_bot - el_top) * (static_cast<float>(ring) / static_cast<float>(n_rings - 1))) * 0.01745329252f;
    float const az = 6.283185307f * ((static_cast<float>(step) + u_jit) / static_cast<float>(steps));
    float sa, ca, se, ce;
    sincosf(az, &sa, &ca);
    sincosf(el, &se, &ce);
    out[g] = make_float4(range * ce * ca, range * ce * sa, range * se, inten);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Validation pass for FROM_W buffers: the reference asserts TimeIsInRange for every point stamp
// (trajectory_interpolation.cpp:32,47); the fp32 deskew kernels do not spend a flag store per launch on it, so callers that
// cannot vouch for their fractions run this 16 B/point read-only pass first.  Bit 0: some w outside [0, 1] or NaN.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlockThreads)
    CheckFractionsKernel(const float4* __restrict__ in, int64_t n, int* __restrict__ flags) {
  int64_t const stride = static_cast<int64_t>(gridDim.x) * kBlockThreads;
  int bad = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kBlockThreads + threadIdx.x; i < n; i += stride) {
    float const w = LoadPoint<0>(in + i).w;
    bad |= !(w >= 0.0f && w <= 1.0f);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flags, 1);
}

// ---------------------------------------------------------------------------------------------------------------
// Per-frame 64-bit checksums of a batch (verification aid: "is the result of a sharded run bit-identical to the unsharded
// one?" without moving 20 GB).  Word j of a frame (its 4 n_f float bit patterns, j counted from the frame's first point)
// contributes (bits + 0x9E3779B9) * (2 j + 1) mod 2^64; a frame's checksum is the wrapping sum, so it does not depend on
// which thread, CTA or GPU added which word, but it does depend on every bit and on every word's position.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kChecksumBlockThreads = 256;
constexpr int64_t kChecksumItemPoints = 4096;

__global__ void __launch_bounds__(kChecksumBlockThreads)
    FrameChecksumsKernel(const uint4* __restrict__ points, const int64_t* __restrict__ offsets, int n_frames, int64_t n,
                         double frames_per_point, unsigned long long* __restrict__ sums) {
  __shared__ unsigned long long warp_sums[kChecksumBlockThreads / 32];
  int64_t const n_items = (n + kChecksumItemPoints - 1) / kChecksumItemPoints;
  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    int64_t p0 = item * kChecksumItemPoints;
    int64_t const p1 = (p0 + kChecksumItemPoints < n) ? (p0 + kChecksumItemPoints) : n;
    int f = LocateFrame(offsets, n_frames, p0, frames_per_point);
    while (p0 < p1 && f < n_frames) {
      int64_t const frame_begin = __ldg(offsets + f);
      int64_t const frame_end = __ldg(offsets + f + 1);
      if (frame_end <= p0) {
        ++f;
        continue;
      }
      int64_t const seg_end = frame_end < p1 ? frame_end : p1;
      unsigned long long acc = 0;
      for (int64_t i = p0 + threadIdx.x; i < seg_end; i += kChecksumBlockThreads) {
        uint4 const v = __ldg(points + i);
        unsigned long long const j = 4ull * static_cast<unsigned long long>(i - frame_begin);
        acc += (static_cast<unsigned long long>(v.x) + 0x9E3779B9ull) * (2 * j + 1);
        acc += (static_cast<unsigned long long>(v.y) + 0x9E3779B9ull) * (2 * j + 3);
        acc += (static_cast<unsigned long long>(v.z) + 0x9E3779B9ull) * (2 * j + 5);
        acc += (static_cast<unsigned long long>(v.w) + 0x9E3779B9ull) * (2 * j + 7);
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
      if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned long long total = 0;
#pragma unroll
        for (int w = 0; w < kChecksumBlockThreads / 32; ++w) total += warp_sums[w];
        atomicAdd(sums + f, total);
      }
      __syncthreads();
      p0 = seg_end;
      ++f;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------------------------------
template <int MODE, int VEC, int UNROLL, int HINT, int BLOCK>
cudaError_t LaunchFrameT(const float* in, float* out, int64_t n, const kmc_b200_frame_params& P, const LaunchConfig& cfg,
                         int sm_count, cudaStream_t stream) {
  int64_t const tile = static_cast<int64_t>(BLOCK) * UNROLL * VEC;
  int64_t const item_points = tile * cfg.item_tiles;
  int64_t const n_items = (n + item_points - 1) / item_points;
  int64_t grid = static_cast<int64_t>(sm_count) * cfg.ctas_per_sm;  // CTAs resident per SM, whatever their size
  if (grid > n_items) grid = n_items;
  if (grid < 1) grid = 1;
  if (kmc_b200::internal::TuneValue("pdl", 1) != 0) {  // 10 M-point frames back to back: 48.1 us against 49.8 (profiles/r02_config5_pdl.log)
    // back-to-back launches on one stream: the next kernel's CTAs become resident while this one's last wave drains and wait
    // (griddepcontrol.wait) for its completion — the launch gap disappears, the stream order does not
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(static_cast<unsigned>(grid));
    lc.blockDim = dim3(BLOCK);
    lc.dynamicSmemBytes = 0;
    lc.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = &attr;
    lc.numAttrs = 1;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaLaunchKernelEx(&lc, DeskewFrameKernel<MODE, VEC, UNROLL, HINT, BLOCK>, reinterpret_cast<const float4*>(in),
                              reinterpret_cast<float4*>(out), n, item_points, P);
  }
  DeskewFrameKernel<MODE, VEC, UNROLL, HINT, BLOCK><<<static_cast<unsigned>(grid), BLOCK, 0, stream>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), n, item_points, P);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

template <int MODE, int VEC, int UNROLL, int HINT, int BLOCK>
cudaError_t LaunchBatchT(const float* in, float* out, const int64_t* offsets, const kmc_b200_frame_params* table,
                         int32_t n_frames, int64_t n, int64_t point_base, int64_t n_batch_points, const LaunchConfig& cfg,
                         int sm_count, cudaStream_t stream) {
  int64_t const tile = static_cast<int64_t>(BLOCK) * UNROLL * VEC;
  int64_t const item_points = tile * cfg.item_tiles;
  int64_t const n_items = (n + item_points - 1) / item_points;
  int64_t grid = static_cast<int64_t>(sm_count) * cfg.ctas_per_sm;  // CTAs resident per SM, whatever their size
  if (grid > n_items) grid = n_items;
  if (grid < 1) grid = 1;
  double const frames_per_point = static_cast<double>(n_frames) / static_cast<double>(n_batch_points);
  DeskewBatchKernel<MODE, VEC, UNROLL, HINT, BLOCK><<<static_cast<unsigned>(grid), BLOCK, 0, stream>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), offsets, table, n_frames, n, item_points,
      point_base, frames_per_point);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

// Expands the runtime knobs into template instantiations.  (vec, unroll) in {1,2}x{1,2}, hint in {0,1}, block in {128,256,512}.
template <int MODE, int V, int U, int H>
cudaError_t DispatchFrameBlock(const float* in, float* out, int64_t n, const kmc_b200_frame_params& P, const LaunchConfig& cfg,
                               int sm, cudaStream_t st) {
  if (cfg.block == 128) return LaunchFrameT<MODE, V, U, H, 128>(in, out, n, P, cfg, sm, st);
  if (cfg.block == 512) return LaunchFrameT<MODE, V, U, H, 512>(in, out, n, P, cfg, sm, st);
  return LaunchFrameT<MODE, V, U, H, 256>(in, out, n, P, cfg, sm, st);
}

template <int MODE>
cudaError_t DispatchFrame(const float* in, float* out, int64_t n, const kmc_b200_frame_params& P, const LaunchConfig& cfg,
                          int sm, cudaStream_t st) {
#define KMC_FRAME_CALL(V, U)                                                                          \
  return cfg.hint == 1 ? DispatchFrameBlock<MODE, V, U, 1>(in, out, n, P, cfg, sm, st)                \
                  : DispatchFrameBlock<MODE, V, U, 0>(in, out, n, P, cfg, sm, st);
  if (cfg.vec == 2) {
    if (cfg.unroll >= 2) { KMC_FRAME_CALL(2, 2) }
    KMC_FRAME_CALL(2, 1)
  }
  if (cfg.unroll >= 2) { KMC_FRAME_CALL(1, 2) }
  KMC_FRAME_CALL(1, 1)
#undef KMC_FRAME_CALL
}

template <int MODE, int V, int U, int H>
cudaError_t DispatchBatchBlock(const float* in, float* out, const int64_t* offsets, const kmc_b200_frame_params* table,
                               int32_t n_frames, int64_t n, int64_t base, int64_t nb, const LaunchConfig& cfg, int sm,
                               cudaStream_t st) {
  if (cfg.block == 128) return LaunchBatchT<MODE, V, U, H, 128>(in, out, offsets, table, n_frames, n, base, nb, cfg, sm, st);
  if (cfg.block == 512) return LaunchBatchT<MODE, V, U, H, 512>(in, out, offsets, table, n_frames, n, base, nb, cfg, sm, st);
  return LaunchBatchT<MODE, V, U, H, 256>(in, out, offsets, table, n_frames, n, base, nb, cfg, sm, st);
}

template <int MODE>
cudaError_t DispatchBatch(const float* in, float* out, const int64_t* offsets, const kmc_b200_frame_params* table,
                          int32_t n_frames, int64_t n, int64_t base, int64_t nb, const LaunchConfig& cfg, int sm,
                          cudaStream_t st) {
#define KMC_BATCH_CALL(V, U)                                                                                              \
  return cfg.hint == 1 ? DispatchBatchBlock<MODE, V, U, 1>(in, out, offsets, table, n_frames, n, base, nb, cfg, sm, st)    \
                  : DispatchBatchBlock<MODE, V, U, 0>(in, out, offsets, table, n_frames, n, base, nb, cfg, sm, st);
  if (cfg.vec == 2 && cfg.hint >= 3) {  // L2 eviction-priority experiments (256-bit accesses only)
    switch (cfg.hint) {
      case 3: return DispatchBatchBlock<MODE, 2, 1, 3>(in, out, offsets, table, n_frames, n, base, nb, cfg, sm, st);
      case 4: return DispatchBatchBlock<MODE, 2, 1, 4>(in, out, offsets, table, n_frames, n, base, nb, cfg, sm, st);
      case 5: return DispatchBatchBlock<MODE, 2, 1, 5>(in, out, offsets, table, n_frames, n, base, nb, cfg, sm, st);
      default: return DispatchBatchBlock<MODE, 2, 1, 6>(in, out, offsets, table, n_frames, n, base, nb, cfg, sm, st);
    }
  }
  if (cfg.vec == 2) {
    if (cfg.unroll >= 2) { KMC_BATCH_CALL(2, 2) }
    KMC_BATCH_CALL(2, 1)
  }
  if (cfg.unroll >= 2) { KMC_BATCH_CALL(1, 2) }
  KMC_BATCH_CALL(1, 1)
#undef KMC_BATCH_CALL
}

// KMC_B200_TUNE="key=value,key=value,...": experiment knob for tools/sweep.py.  Parsed without strtok (re-entrant: the
// multi-GPU entry point calls this from several host threads).
void ApplyTuneEnv(LaunchConfig& cfg) {
  const char* env = std::getenv("KMC_B200_TUNE");
  if (!env || !*env) return;
  const char* p = env;
  while (*p) {
    const char* key = p;
    while (*p && *p != '=' && *p != ',') ++p;
    size_t const key_len = static_cast<size_t>(p - key);
    int value = 0;
    bool has_value = false;
    if (*p == '=') {
      ++p;
      value = std::atoi(p);
      has_value = true;
      while (*p && *p != ',') ++p;
    }
    if (*p == ',') ++p;
    if (!has_value) continue;
    auto is = [&](const char* name) { return std::strlen(name) == key_len && std::strncmp(key, name, key_len) == 0; };
    if (is("vec")) cfg.vec = value;
    else if (is("unroll")) cfg.unroll = value;
    else if (is("hint")) cfg.hint = value;
    else if (is("ctas")) cfg.ctas_per_sm = value;
    else if (is("item_tiles")) cfg.item_tiles = value;
    else if (is("block")) cfg.block = value;
    else if (is("bulk")) cfg.bulk = value;
    else if (is("stages")) cfg.stages = value;
  }
}

}  // namespace

uint64_t LaunchCount() { return g_launches.load(std::memory_order_relaxed); }

LaunchConfig PickConfig(int64_t n_points, bool aligned32, bool in_place, int sm_count, bool batch) {
  // Defaults from the B200 sweeps.  Per thread: one 256-bit load + one 256-bit store per tile, plain ld/st, 128-thread CTAs
  // (40 registers, up to 16 CTAs resident per SM).  Round 1 shipped a PERSISTENT grid of 9 CTAs per SM taking work items
  // round-robin (6.49-6.55 TB/s on the headline batch); ncu then showed the SMs' active cycles spread by 8 % under that
  // static split (sm__cycles_active min 84.5 K / max 92 K on the 10 M-point frame: SMs differ in their distance to the
  // memory controllers), i.e. the launch waits for its slowest SM.  Handing work items to CTAs through the hardware
  // scheduler instead — ONE CTA PER WORK ITEM, grid = number of items — removes that tail:
  //   batch 4000 x 130 000 points, items of 8 tiles (2048 points)     6 826 GB/s  (persistent 9 x 128: 6 558; torch copy_: 6 668)
  //   one 10 M-point frame, items of 1 tile                           6 415 GB/s back to back (persistent: 5 949)
  //   one 100 M-point frame                                           6 855 GB/s (persistent: 6 272)
  // (profiles/r02_frame_sizes.log, r02_batch_nonpersistent.log, r02_batch_items.log).  The batch kernel takes items of 8 tiles
  // because every item pays a frame lookup and a record fetch (1 / 2 / 4 / 8 / 16 / 32 / 64 tiles: 5 811 / 6 314 / 6 511 / 6 826 /
  // 6 724 / 6 650 / 6 609 GB/s); the single-frame
  // kernel has no per-item cost and takes one tile per CTA.  ctas_per_sm is the cap of the grid in CTAs per SM: 4096 means
  // "never persistent below 600 K items", 9 restores round 1's shape (KMC_B200_TUNE ctas=9,item_tiles=16).
  LaunchConfig cfg;
  cfg.vec = 2;
  cfg.unroll = 1;
  cfg.hint = 0;
  cfg.block = 128;
  cfg.ctas_per_sm = 4096;
  cfg.item_tiles = batch ? 8 : 1;
  cfg.bulk = 0;
  cfg.stages = 4;
  // Mid-size batches: keep at least ~16 items per resident CTA slot (148 x 9) so that the scheduler has something to balance;
  // tiny inputs (latency bound): 128-bit accesses to spread over as many CTAs as possible.
  int64_t const tile = static_cast<int64_t>(cfg.block) * cfg.unroll * cfg.vec;
  int64_t const slots = static_cast<int64_t>(sm_count) * 9;
  int64_t const tiles_per_slot = n_points / (slots * tile);
  if (tiles_per_slot < 32 * cfg.item_tiles) {
    int64_t const t = tiles_per_slot / 32;
    cfg.item_tiles = static_cast<int>(t < 1 ? 1 : t);
    if (tiles_per_slot < 1) cfg.vec = 1;
  }
  ApplyTuneEnv(cfg);
  if (!aligned32) cfg.vec = 1;
  if (in_place && cfg.hint == 1) cfg.hint = 0;  // the .nc path assumes the input is read-only for the whole kernel
  if (cfg.vec != 2) cfg.vec = 1;
  if (cfg.unroll < 1) cfg.unroll = 1;
  if (cfg.hint < 0 || cfg.hint > 6 || cfg.hint == 2) cfg.hint = 0;  // 3-6: L2 eviction-priority experiments (batch kernel)
  if (cfg.block != 128 && cfg.block != 512) cfg.block = 256;
  if (cfg.ctas_per_sm < 1) cfg.ctas_per_sm = 1;
  if (cfg.ctas_per_sm > 4096) cfg.ctas_per_sm = 4096;  // beyond the residency limit the grid stops being persistent (experiment)
  if (cfg.item_tiles < 1) cfg.item_tiles = 1;
  return cfg;
}

cudaError_t LaunchDeskewFrame(const float* in, float* out, int64_t n, const kmc_b200_frame_params& params, int mode,
                              const LaunchConfig& cfg, int sm_count, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (cfg.bulk) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return LaunchDeskewFrameBulk(in, out, n, params, mode, cfg.block, cfg.unroll, cfg.stages, cfg.ctas_per_sm, sm_count, stream);
  }
  if (mode == KMC_B200_TIME_FROM_AZIMUTH) return DispatchFrame<KMC_B200_TIME_FROM_AZIMUTH>(in, out, n, params, cfg, sm_count, stream);
  return DispatchFrame<KMC_B200_TIME_FROM_W>(in, out, n, params, cfg, sm_count, stream);
}

cudaError_t LaunchDeskewBatch(const float* in, float* out, const int64_t* offsets_dev,
                              const kmc_b200_frame_params* params_dev, int32_t n_frames, int64_t n_points,
                              int64_t point_base, int64_t n_batch_points, int mode, const LaunchConfig& cfg, int sm_count,
                              cudaStream_t stream) {
  if (n_points <= 0 || n_frames <= 0) return cudaSuccess;
  if (cfg.bulk) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return LaunchDeskewBatchBulk(in, out, offsets_dev, params_dev, n_frames, n_points, point_base, n_batch_points, mode, cfg.block,
                                 cfg.unroll, cfg.stages, cfg.ctas_per_sm, sm_count, stream);
  }
  if (mode == KMC_B200_TIME_FROM_AZIMUTH)
    return DispatchBatch<KMC_B200_TIME_FROM_AZIMUTH>(in, out, offsets_dev, params_dev, n_frames, n_points, point_base,
                                                     n_batch_points, cfg, sm_count, stream);
  return DispatchBatch<KMC_B200_TIME_FROM_W>(in, out, offsets_dev, params_dev, n_frames, n_points, point_base,
                                             n_batch_points, cfg, sm_count, stream);
}

namespace {
// Launch shape of the projection kernels: threads per CTA, cap of the grid in CTAs per SM and 128- vs 256-bit accesses;
// KMC_B200_TUNE="pblock=128,pctas=9,pvec=2" overrides the defaults for tools/sweep_projection.py.
// Round 1 shipped persistent grids whose best residency fell as the stores per thread rose (8 / 5 / 6 / 5 / 5 CTAs per SM).
// As for the deskew kernels, ONE CTA PER TILE — the hardware scheduler balancing SMs of different speed — beats every
// persistent shape (profiles/r02_sweep_projection.log, 100 M points, torch copy_ 6 550, fill_ 7 166 GB/s on that box):
//   1 load + 1 store  (project only)             128 threads   6 939 GB/s   (persistent 8 x 128: 6 611)
//   1 load + 1 store  (deskew + project)         128           6 686        (persistent 5 x 256: 6 371)
//   1 load + 2 stores (deskew + project + cloud) 128           6 854        (persistent 6 x 128: 6 335)
//   1 load + 4 stores (four cameras)             128 / 256     6 714 / 6 723 (persistent 5 x 128: 6 244)
//   1 load + 5 stores (four cameras + cloud)     128 / 256     6 669 / 6 678 (persistent 5 x 128: 6 171)
struct ProjectShape {
  int block = 128;
  int ctas_per_sm = 65536;
  int vec = 2;
};

ProjectShape PickProjectShape(bool aligned32, int stores_per_thread, bool deskew) {
  ProjectShape shape;
  (void)stores_per_thread;
  (void)deskew;
  if (const char* env = std::getenv("KMC_B200_TUNE")) {
    auto find = [&](const char* key, int* out) {
      size_t const len = std::strlen(key);
      for (const char* p = env; *p;) {
        if (std::strncmp(p, key, len) == 0 && p[len] == '=') *out = std::atoi(p + len + 1);
        while (*p && *p != ',') ++p;
        if (*p == ',') ++p;
      }
    };
    find("pblock", &shape.block);
    find("pctas", &shape.ctas_per_sm);
    find("pvec", &shape.vec);
  }
  if (shape.block != 128) shape.block = 256;
  shape.ctas_per_sm = std::min(std::max(shape.ctas_per_sm, 1), 65536);  // > 16: one CTA per tile, not persistent
  if (!aligned32 || shape.vec != 2) shape.vec = 1;
  return shape;
}

unsigned ProjectGrid(int64_t n, const ProjectShape& shape, int sm_count) {
  int64_t const per_cta = static_cast<int64_t>(shape.block) * shape.vec;
  int64_t const grid = (n + per_cta - 1) / per_cta;
  return static_cast<unsigned>(std::min<int64_t>(grid, static_cast<int64_t>(sm_count) * shape.ctas_per_sm));
}

template <bool DESKEW, bool WRITE_CLOUD, int MODE>
cudaError_t LaunchProjectT(const float* in, float* cloud_out, float* pix_out, int64_t n, const kmc_b200_frame_params& P,
                           const kmc_b200_camera_params& K, bool vec2, int sm_count, cudaStream_t stream) {
  ProjectShape const shape = PickProjectShape(vec2, WRITE_CLOUD ? 2 : 1, DESKEW);
  unsigned const grid = ProjectGrid(n, shape, sm_count);
  auto const* in4 = reinterpret_cast<const float4*>(in);
  auto* cloud4 = reinterpret_cast<float4*>(cloud_out);
  auto* pix4 = reinterpret_cast<float4*>(pix_out);
#define KMC_PROJECT_CALL(V, B) ProjectFrameKernel<DESKEW, WRITE_CLOUD, MODE, V, B><<<grid, B, 0, stream>>>(in4, cloud4, pix4, n, P, K)
  if (shape.vec == 2) {
    if (shape.block == 128) KMC_PROJECT_CALL(true, 128);
    else KMC_PROJECT_CALL(true, 256);
  } else {
    if (shape.block == 128) KMC_PROJECT_CALL(false, 128);
    else KMC_PROJECT_CALL(false, 256);
  }
#undef KMC_PROJECT_CALL
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

template <bool DESKEW, bool WRITE_CLOUD, int MODE>
void LaunchProject4T(const float4* in4, float4* cloud4, const PixelPlanes4& planes, int64_t n, const kmc_b200_frame_params& P,
                     const Cameras4& K, const ProjectShape& shape, unsigned grid, cudaStream_t stream) {
#define KMC_PROJECT4_CALL(V, B) ProjectFrame4Kernel<DESKEW, WRITE_CLOUD, MODE, V, B><<<grid, B, 0, stream>>>(in4, cloud4, planes, n, P, K)
  if (shape.vec == 2) {
    if (shape.block == 128) KMC_PROJECT4_CALL(true, 128);
    else KMC_PROJECT4_CALL(true, 256);
  } else {
    if (shape.block == 128) KMC_PROJECT4_CALL(false, 128);
    else KMC_PROJECT4_CALL(false, 256);
  }
#undef KMC_PROJECT4_CALL
}
}  // namespace

cudaError_t LaunchProject(const float* in, float* cloud_out, float* pix_out, int64_t n, const kmc_b200_frame_params* params,
                          const kmc_b200_camera_params& camera, int mode, bool vec2, int sm_count, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (!params) {
    kmc_b200_frame_params const none{};
    return LaunchProjectT<false, false, 0>(in, nullptr, pix_out, n, none, camera, vec2, sm_count, stream);
  }
  bool const az = (mode == KMC_B200_TIME_FROM_AZIMUTH);
  if (cloud_out) {
    return az ? LaunchProjectT<true, true, KMC_B200_TIME_FROM_AZIMUTH>(in, cloud_out, pix_out, n, *params, camera, vec2, sm_count, stream)
              : LaunchProjectT<true, true, KMC_B200_TIME_FROM_W>(in, cloud_out, pix_out, n, *params, camera, vec2, sm_count, stream);
  }
  return az ? LaunchProjectT<true, false, KMC_B200_TIME_FROM_AZIMUTH>(in, nullptr, pix_out, n, *params, camera, vec2, sm_count, stream)
            : LaunchProjectT<true, false, KMC_B200_TIME_FROM_W>(in, nullptr, pix_out, n, *params, camera, vec2, sm_count, stream);
}

cudaError_t LaunchProject4(const float* in, float* cloud_out, float* const pix_out[4], int64_t n, const kmc_b200_frame_params* params,
                           const kmc_b200_camera_params cameras[4], int mode, int sm_count, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  Cameras4 K;
  PixelPlanes4 planes;
  bool aligned32 = (reinterpret_cast<uintptr_t>(in) % 32 == 0) && (!cloud_out || reinterpret_cast<uintptr_t>(cloud_out) % 32 == 0);
  for (int c = 0; c < 4; ++c) {
    K.cam[c] = cameras[c];
    planes.plane[c] = reinterpret_cast<float4*>(pix_out[c]);
    aligned32 = aligned32 && (reinterpret_cast<uintptr_t>(pix_out[c]) % 32 == 0);
  }
  ProjectShape const shape = PickProjectShape(aligned32, cloud_out ? 5 : 4, params != nullptr);
  unsigned const grid = ProjectGrid(n, shape, sm_count);
  auto const* in4 = reinterpret_cast<const float4*>(in);
  auto* cloud4 = reinterpret_cast<float4*>(cloud_out);
  kmc_b200_frame_params const none{};
  if (!params) {
    LaunchProject4T<false, false, 0>(in4, nullptr, planes, n, none, K, shape, grid, stream);
  } else if (mode == KMC_B200_TIME_FROM_AZIMUTH) {
    if (cloud_out) LaunchProject4T<true, true, KMC_B200_TIME_FROM_AZIMUTH>(in4, cloud4, planes, n, *params, K, shape, grid, stream);
    else LaunchProject4T<true, false, KMC_B200_TIME_FROM_AZIMUTH>(in4, nullptr, planes, n, *params, K, shape, grid, stream);
  } else {
    if (cloud_out) LaunchProject4T<true, true, KMC_B200_TIME_FROM_W>(in4, cloud4, planes, n, *params, K, shape, grid, stream);
    else LaunchProject4T<true, false, KMC_B200_TIME_FROM_W>(in4, nullptr, planes, n, *params, K, shape, grid, stream);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

namespace {
template <int MODE, int NCAM, bool WRITE_CLOUD>
void LaunchProjectBatchT(const float4* in4, float4* cloud4, const PixelPlanes4& planes, const int64_t* offsets, const kmc_b200_frame_params* table,
                         int32_t n_frames, int64_t n, const Cameras4& K, const ProjectShape& shape, int sm_count, cudaStream_t stream) {
  int const item_tiles = std::max(1, kmc_b200::internal::TuneValue("pitem_tiles", 8));
  int64_t const item_points = static_cast<int64_t>(shape.block) * shape.vec * item_tiles;
  int64_t const n_items = (n + item_points - 1) / item_points;
  unsigned const grid = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(n_items, static_cast<int64_t>(sm_count) * shape.ctas_per_sm)));
  double const fpp = static_cast<double>(n_frames) / static_cast<double>(n);
#define KMC_PB_CALL(V, B) \
  DeskewProjectBatchKernel<MODE, NCAM, WRITE_CLOUD, V, B><<<grid, B, 0, stream>>>(in4, cloud4, planes, offsets, table, n_frames, n, item_points, fpp, K)
  if (shape.vec == 2) {
    if (shape.block == 128) KMC_PB_CALL(2, 128);
    else KMC_PB_CALL(2, 256);
  } else {
    if (shape.block == 128) KMC_PB_CALL(1, 128);
    else KMC_PB_CALL(1, 256);
  }
#undef KMC_PB_CALL
}
}  // namespace

cudaError_t LaunchDeskewProjectBatch(const float* in, float* cloud_out, float* const pix_out[], int n_cameras, const int64_t* offsets_dev,
                                     const kmc_b200_frame_params* params_dev, int32_t n_frames, int64_t n_points,
                                     const kmc_b200_camera_params cameras[], int mode, int sm_count, cudaStream_t stream) {
  if (n_points <= 0 || n_frames <= 0) return cudaSuccess;
  Cameras4 K{};
  PixelPlanes4 planes{};
  bool aligned32 = (reinterpret_cast<uintptr_t>(in) % 32 == 0) && (!cloud_out || reinterpret_cast<uintptr_t>(cloud_out) % 32 == 0);
  for (int c = 0; c < n_cameras; ++c) {
    K.cam[c] = cameras[c];
    planes.plane[c] = reinterpret_cast<float4*>(pix_out[c]);
    aligned32 = aligned32 && (reinterpret_cast<uintptr_t>(pix_out[c]) % 32 == 0);
  }
  ProjectShape const shape = PickProjectShape(aligned32, n_cameras + (cloud_out ? 1 : 0), true);
  auto const* in4 = reinterpret_cast<const float4*>(in);
  auto* cloud4 = reinterpret_cast<float4*>(cloud_out);
  bool const az = (mode == KMC_B200_TIME_FROM_AZIMUTH);
#define KMC_PB_DISPATCH(NCAM)                                                                                                              \
  if (cloud_out) {                                                                                                                         \
    if (az) LaunchProjectBatchT<KMC_B200_TIME_FROM_AZIMUTH, NCAM, true>(in4, cloud4, planes, offsets_dev, params_dev, n_frames, n_points, K, shape, sm_count, stream); \
    else LaunchProjectBatchT<KMC_B200_TIME_FROM_W, NCAM, true>(in4, cloud4, planes, offsets_dev, params_dev, n_frames, n_points, K, shape, sm_count, stream);          \
  } else {                                                                                                                                 \
    if (az) LaunchProjectBatchT<KMC_B200_TIME_FROM_AZIMUTH, NCAM, false>(in4, nullptr, planes, offsets_dev, params_dev, n_frames, n_points, K, shape, sm_count, stream); \
    else LaunchProjectBatchT<KMC_B200_TIME_FROM_W, NCAM, false>(in4, nullptr, planes, offsets_dev, params_dev, n_frames, n_points, K, shape, sm_count, stream);          \
  }
  if (n_cameras == 1) { KMC_PB_DISPATCH(1) }
  else { KMC_PB_DISPATCH(4) }
#undef KMC_PB_DISPATCH
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t LaunchDeskewCloudF64Batch(const double* cloud, const double* stamps, double* out, const int64_t* offsets_dev,
                                      const kmc_b200_frame_params* params_dev, const double* times_dev, int32_t n_frames, int64_t n_points,
                                      int* flags_dev, int /*sm_count: the grid is (lanes per frame) x (frames)*/, cudaStream_t stream) {
  if (n_frames <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(flags_dev, 0, static_cast<size_t>(n_frames) * sizeof(int), stream);
  if (e != cudaSuccess || n_points <= 0) return e;
  // gridDim.x CTAs share a frame: enough of them that each strides over ~f64_item_tiles tiles of an average frame
  // (KITTI-size and larger frames: 2 tiles of 512 points, 6.70-6.81 TB/s; small frames want 4: 6.65 against 6.30 at 13 001 points)
  int const block = kmc_b200::internal::TuneValue("f64_block", 256) == 128 ? 128 : 256;
  int64_t const avg = (n_points + n_frames - 1) / n_frames;
  int64_t const tiles = std::max(1, kmc_b200::internal::TuneValue("f64_item_tiles", avg >= 65536 ? 2 : 4));
  bool const pair = kmc_b200::internal::TuneValue("f64_pair", 1) != 0;
  int const min_ctas = kmc_b200::internal::TuneValue("f64_min_ctas", 4);
  int64_t const per_cta = static_cast<int64_t>(block) * (pair ? 2 : 1) * tiles;
  unsigned const gx = static_cast<unsigned>(std::min<int64_t>(std::max<int64_t>((avg + per_cta - 1) / per_cta, 1), 65535));
  unsigned const gy = static_cast<unsigned>(std::min<int32_t>(n_frames, 32768));
  unsigned const gz = static_cast<unsigned>((n_frames + static_cast<int32_t>(gy) - 1) / static_cast<int32_t>(gy));
  dim3 const grid(gx, gy, gz);
#define KMC_F64B(BLOCK, PAIR, MIN) DeskewCloudF64BatchKernel<BLOCK, PAIR, MIN><<<grid, BLOCK, 0, stream>>>(cloud, stamps, out, offsets_dev, params_dev, times_dev, n_frames, flags_dev)
  if (pair) {
    if (block == 128) { if (min_ctas == 8) KMC_F64B(128, true, 8); else KMC_F64B(128, true, 6); }
    else { if (min_ctas == 4) KMC_F64B(256, true, 4); else KMC_F64B(256, true, 3); }
  } else {
    if (block == 128) KMC_F64B(128, false, 4); else KMC_F64B(256, false, 4);
  }
#undef KMC_F64B
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t LaunchPseudoTimeStamps(const float* in, double* stamps, int64_t n, double start, double end, int sm_count,
                                   cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  bool const vec2 = (reinterpret_cast<uintptr_t>(in) % 32 == 0) && (reinterpret_cast<uintptr_t>(stamps) % 16 == 0);
  int64_t const per_cta = static_cast<int64_t>(kStampBlockThreads) * (vec2 ? 2 : 1);
  int64_t grid = (n + per_cta - 1) / per_cta;
  int64_t const cap = static_cast<int64_t>(sm_count) * std::max(1, kmc_b200::internal::TuneValue("stamp_ctas", 4096));  // one CTA per tile: 7 170 GB/s vs 6 289 persistent (24 B/point, profiles/r02_sweep_secondary.log)
  if (grid > cap) grid = cap;
  auto const* in4 = reinterpret_cast<const float4*>(in);
  if (vec2) PseudoTimeStampsKernel<true><<<static_cast<unsigned>(grid), kStampBlockThreads, 0, stream>>>(in4, stamps, n, start, end - start);
  else PseudoTimeStampsKernel<false><<<static_cast<unsigned>(grid), kStampBlockThreads, 0, stream>>>(in4, stamps, n, start, end - start);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t LaunchDeskewCloudF64(const double* cloud, const double* stamps, double* out, int64_t n, double t1, double t2, double x_req,
                                 const kmc_b200_frame_params& params, int* flags_dev, int sm_count, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  // two points per thread, one CTA per tile of 512 points: 6.90 TB/s (2 tiles 6.78, 4 tiles 6.67, persistent 8 CTAs per SM 6.18;
  // the one-point kernel, f64_pair=0, persistent at 8 x 256 threads per SM: 6.43-6.59 — profiles/r02_sweep_f64_batch_pair_defaults.log)
  bool const pair = kmc_b200::internal::TuneValue("f64_pair", 1) != 0;
  if (pair) {
    int64_t const per_cta = static_cast<int64_t>(kBlockThreads) * 2 * std::max(1, kmc_b200::internal::TuneValue("f64_item_tiles", 1));
    int64_t grid = (n + per_cta - 1) / per_cta;
    int64_t const cap = static_cast<int64_t>(sm_count) * std::max(1, kmc_b200::internal::TuneValue("f64_ctas", 4096));
    if (grid > cap) grid = cap;
    DeskewCloudF64Kernel<true><<<static_cast<unsigned>(grid), kBlockThreads, 0, stream>>>(cloud, stamps, out, n, t1, t2, x_req, params, flags_dev);
  } else {
    int64_t grid = (n + kBlockThreads - 1) / kBlockThreads;
    int64_t const cap = static_cast<int64_t>(sm_count) * std::max(1, kmc_b200::internal::TuneValue("f64_ctas", 8));
    if (grid > cap) grid = cap;
    DeskewCloudF64Kernel<false><<<static_cast<unsigned>(grid), kBlockThreads, 0, stream>>>(cloud, stamps, out, n, t1, t2, x_req, params, flags_dev);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t LaunchDeskewDeltaColumns(const float* columns_in, float* columns_out, int64_t stride_points, bool has_w,
                                     const kmc_b200_frame_params& params, int sm_count, cudaStream_t stream, bool over_pcie) {
  if (stride_points <= 0) return cudaSuccess;
  int64_t const stride4 = stride_points / 4;
  int64_t grid = (stride4 + kDeltaBlockThreads - 1) / kDeltaBlockThreads;
  // columns in pinned host memory: a small grid keeps the number of PCIe read requests in flight low (profiles/r02_latency_probe.log)
  int64_t const cap = static_cast<int64_t>(sm_count) * (over_pcie ? std::max(1, kmc_b200::internal::TuneValue("f64_zc_ctas", 2)) : 8);
  if (grid > cap) grid = cap;
  auto const* in4 = reinterpret_cast<const float4*>(columns_in);
  auto* out4 = reinterpret_cast<float4*>(columns_out);
  if (has_w) DeskewDeltaColumnsKernel<true><<<static_cast<unsigned>(grid), kDeltaBlockThreads, 0, stream>>>(in4, out4, stride4, params);
  else DeskewDeltaColumnsKernel<false><<<static_cast<unsigned>(grid), kDeltaBlockThreads, 0, stream>>>(in4, out4, stride4, params);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t LaunchPseudoTimeStampsXy(const double* x, const double* y, double* stamps, int64_t n, double start, double end,
                                     int sm_count, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  bool const vec2 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(stamps)) & 15) == 0 &&
                    kmc_b200::internal::TuneValue("stamp_vec", 2) == 2;
  int64_t const per_cta = static_cast<int64_t>(kBlockThreads) * (vec2 ? 2 : 1);
  int64_t grid = (n + per_cta - 1) / per_cta;
  int64_t const cap = static_cast<int64_t>(sm_count) * std::max(1, kmc_b200::internal::TuneValue("stamp_ctas", 6));  // FP64-heavy: stays persistent
  if (grid > cap) grid = cap;
  if (vec2) PseudoTimeStampsXyKernel<true><<<static_cast<unsigned>(grid), kBlockThreads, 0, stream>>>(x, y, stamps, n, start, end - start);
  else PseudoTimeStampsXyKernel<false><<<static_cast<unsigned>(grid), kBlockThreads, 0, stream>>>(x, y, stamps, n, start, end - start);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t LaunchCheckFractions(const float* xyzi, int64_t n, int* flags_dev, int sm_count, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(flags_dev, 0, sizeof(int), stream);
  if (e != cudaSuccess || n <= 0) return e;
  int64_t grid = (n + 4 * kBlockThreads - 1) / (4 * kBlockThreads);  // four points per thread, one CTA per 1024 points
  int64_t const cap = static_cast<int64_t>(sm_count) * 4096;
  if (grid > cap) grid = cap;
  CheckFractionsKernel<<<static_cast<unsigned>(grid), kBlockThreads, 0, stream>>>(reinterpret_cast<const float4*>(xyzi), n, flags_dev);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t LaunchFrameChecksums(const float* xyzi, const int64_t* offsets_dev, int32_t n_frames, int64_t n_points, uint64_t* sums_dev,
                                 int sm_count, cudaStream_t stream) {
  if (n_frames <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(sums_dev, 0, static_cast<size_t>(n_frames) * sizeof(uint64_t), stream);
  if (e != cudaSuccess || n_points <= 0) return e;
  int64_t grid = (n_points + kChecksumItemPoints - 1) / kChecksumItemPoints;  // one CTA per item (see PickConfig)
  int64_t const cap = static_cast<int64_t>(sm_count) * 4096;
  if (grid > cap) grid = cap;
  FrameChecksumsKernel<<<static_cast<unsigned>(grid), kChecksumBlockThreads, 0, stream>>>(
      reinterpret_cast<const uint4*>(xyzi), offsets_dev, n_frames, n_points, static_cast<double>(n_frames) / static_cast<double>(n_points),
      reinterpret_cast<unsigned long long*>(sums_dev));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t LaunchSynthScans(float* out, int64_t points_per_scan, int32_t n_scans, int32_t n_rings, uint64_t seed,
                             int64_t first_scan_index, int sm_count, cudaStream_t stream) {
  int64_t const n_total = points_per_scan * n_scans;
  if (n_total <= 0) return cudaSuccess;
  int64_t grid = (n_total + kBlockThreads - 1) / kBlockThreads;
  int64_t const cap = static_cast<int64_t>(sm_count) * 8;
  if (grid > cap) grid = cap;
  SynthScansKernel<<<static_cast<unsigned>(grid), kBlockThreads, 0, stream>>>(reinterpret_cast<float4*>(out), points_per_scan,
                                                                              n_total, n_rings, seed, first_scan_index);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

}  // namespace kmc_b200::dev
