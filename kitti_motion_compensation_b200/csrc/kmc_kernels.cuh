// kmc_kernels.cuh — launch interface of the sm_100a kernels (internal; the public boundary is include/kmc_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "kmc_b200.h"

namespace kmc_b200::dev {

// Kernel shape knobs.  Defaults are chosen by PickConfig(); a sweep (tools/sweep.py) can override them through the
// KMC_B200_TUNE environment variable ("vec=2,unroll=1,hint=0,block=256,ctas=4,item_tiles=16") without recompiling.
struct LaunchConfig {
  int vec;         // points per memory instruction: 1 = 128-bit LDG/STG, 2 = 256-bit LDG/STG (sm_100 only)
  int unroll;      // memory instructions in flight per thread per tile (1, 2)
  int hint;        // 0 = plain ld/st, 1 = L1::no_allocate (+ .nc load)
  int block;       // threads per CTA (128, 256, 512)
  int ctas_per_sm; // grid = min(work items, SMs * ctas_per_sm): <= 16 makes the grid persistent, the default (4096) one CTA per item
  int item_tiles;  // tiles per work item (a work item is the unit a CTA takes per scheduling step)
  int bulk;        // 1 = single-frame kernel variant staged through shared memory by the TMA engine (kmc_kernels_bulk.cu);
                   //     `unroll` then means points per thread per stage (2, 4, 8)
  int stages;      // shared-memory slots of the bulk variant (3, 4)
};

constexpr int kBlockThreads = 256;

// batch: the launch is LaunchDeskewBatch (work items pay a frame lookup) rather than LaunchDeskewFrame.
LaunchConfig PickConfig(int64_t n_points, bool aligned32, bool in_place, int sm_count, bool batch);

cudaError_t LaunchDeskewFrame(const float* in, float* out, int64_t n, const kmc_b200_frame_params& params, int mode,
                              const LaunchConfig& cfg, int sm_count, cudaStream_t stream);

// in/out address the points [point_base, point_base + n_points) of a batch of n_batch_points points whose frame table
// (offsets_dev, n_frames + 1 entries counted from the start of the batch) and records live in device memory.
cudaError_t LaunchDeskewBatch(const float* in, float* out, const int64_t* offsets_dev,
                              const kmc_b200_frame_params* params_dev, int32_t n_frames, int64_t n_points,
                              int64_t point_base, int64_t n_batch_points, int mode, const LaunchConfig& cfg, int sm_count,
                              cudaStream_t stream);

// Single-frame deskew staged through shared memory with cp.async.bulk (measured alternative, see kmc_kernels_bulk.cu).
cudaError_t LaunchDeskewFrameBulk(const float* in, float* out, int64_t n, const kmc_b200_frame_params& P, int mode, int block,
                                  int pts_per_thread, int stages, int ctas_per_sm, int sm_count, cudaStream_t stream);

cudaError_t LaunchDeskewBatchBulk(const float* in, float* out, const int64_t* offsets, const kmc_b200_frame_params* table,
                                  int32_t n_frames, int64_t n, int64_t point_base, int64_t n_batch_points, int mode, int block,
                                  int pts_per_thread, int stages, int ctas_per_sm, int sm_count, cudaStream_t stream);

// Projection onto one rectified camera, optionally fused behind the deskew (params != nullptr) and optionally also
// writing the deskewed cloud (cloud_out != nullptr).
cudaError_t LaunchProject(const float* in, float* cloud_out, float* pix_out, int64_t n, const kmc_b200_frame_params* params,
                          const kmc_b200_camera_params& camera, int mode, bool vec2, int sm_count, cudaStream_t stream);

cudaError_t LaunchPseudoTimeStamps(const float* in, double* stamps, int64_t n, double start, double end, int sm_count,
                                   cudaStream_t stream);

// The same for four cameras in one pass: pix_out[c] receives camera c's records.
cudaError_t LaunchProject4(const float* in, float* cloud_out, float* const pix_out[4], int64_t n, const kmc_b200_frame_params* params,
                           const kmc_b200_camera_params cameras[4], int mode, int sm_count, cudaStream_t stream);

// Deskew in the reference's own layout: column-major N x 4 double cloud + per-point double stamps -> column-major double.
cudaError_t LaunchDeskewCloudF64(const double* cloud, const double* stamps, double* out, int64_t n, double t1, double t2, double x_req,
                                 const kmc_b200_frame_params& params, int* flags_dev, int sm_count, cudaStream_t stream);

// Deskew + projection onto n_cameras (1 or 4) cameras over a batch of frames (tables in device memory as for LaunchDeskewBatch);
// cloud_out may be nullptr.
cudaError_t LaunchDeskewProjectBatch(const float* in, float* cloud_out, float* const pix_out[], int n_cameras, const int64_t* offsets_dev,
                                     const kmc_b200_frame_params* params_dev, int32_t n_frames, int64_t n_points,
                                     const kmc_b200_camera_params cameras[], int mode, int sm_count, cudaStream_t stream);

// A batch of frames in the reference's column-major double layout (frame f: 4 N_f doubles at cloud + 4 offsets[f]); times_dev holds
// (t_start, t_end, t_req) per frame, flags_dev one int per frame (zeroed by the call).
cudaError_t LaunchDeskewCloudF64Batch(const double* cloud, const double* stamps, double* out, const int64_t* offsets_dev,
                                      const kmc_b200_frame_params* params_dev, const double* times_dev, int32_t n_frames, int64_t n_points,
                                      int* flags_dev, int sm_count, cudaStream_t stream);

// Narrow transport of the reference-layout host path: float columns x | y | z | s [| w] of stride_points entries each
// (a multiple of 4, 16-byte aligned) -> float columns dx | dy | dz.
// over_pcie: the columns are pinned HOST memory (zero copy); the grid is then kept small.
cudaError_t LaunchDeskewDeltaColumns(const float* columns_in, float* columns_out, int64_t stride_points, bool has_w,
                                     const kmc_b200_frame_params& params, int sm_count, cudaStream_t stream, bool over_pcie);

cudaError_t LaunchPseudoTimeStampsXy(const double* x, const double* y, double* stamps, int64_t n, double start, double end,
                                     int sm_count, cudaStream_t stream);

// FROM_W validation: *flags_dev (zeroed by the call) gets bit 0 if any w lies outside [0, 1] or is NaN.
cudaError_t LaunchCheckFractions(const float* xyzi, int64_t n, int* flags_dev, int sm_count, cudaStream_t stream);

// Position-weighted 64-bit checksum of every frame of a batch (sums_dev: n_frames entries, zeroed by the call).
cudaError_t LaunchFrameChecksums(const float* xyzi, const int64_t* offsets_dev, int32_t n_frames, int64_t n_points, uint64_t* sums_dev,
                                 int sm_count, cudaStream_t stream);

cudaError_t LaunchSynthScans(float* out, int64_t points_per_scan, int32_t n_scans, int32_t n_rings, uint64_t seed,
                             int64_t first_scan_index, int sm_count, cudaStream_t stream);

// Number of kernel launches issued through this translation unit since process start (bench.py reports it).
uint64_t LaunchCount();

}  // namespace kmc_b200::dev
