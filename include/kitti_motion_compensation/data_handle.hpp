// data_handle.hpp — kmc::b200::DataHandle: RAII owner of a device, its streams and the pinned/device staging buffers.
//
// The reference has no class of this name.  Its buffer-owning object on the path is KittiPclLoader (reference
// include/kitti_motion_compensation/data_io.hpp:19-83): constructed once, non-copyable, owns a fixed 250 000-point x
// 4-float staging buffer that every scan is read into.  DataHandle keeps those ownership rules and moves the buffer to
// where the kernel needs it: three pinned host slots + three device slots per direction, so that file/host data can be
// streamed through the GPU (H2D, fused deskew kernel, D2H overlapped).  It is a thin header-only wrapper over the C
// ABI (kmc_b200_handle_* / kmc_b200_deskew_*_host in kmc_b200.h); errors surface as std::runtime_error.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "kmc_b200.h"

namespace kmc::b200 {

inline void ThrowOnError(int status, const char* what) {
  if (status < 0)  // positive statuses are warnings: the outputs are valid
    throw std::runtime_error(std::string(what) + ": " + kmc_b200_status_string(status) + " — " + kmc_b200_last_error());
}

// Per-frame constants from the scan's start/end pose (column-major 4x4 doubles, Eigen::Affine3d::matrix().data()).
inline kmc_b200_frame_params FrameParamsFromPoses(const double* T_start_colmajor, const double* T_end_colmajor, double t_start,
                                                  double t_end, double t_requested) {
  kmc_b200_frame_params p{};
  ThrowOnError(kmc_b200_frame_params_from_poses(T_start_colmajor, T_end_colmajor, t_start, t_end, t_requested, &p),
               "kmc_b200_frame_params_from_poses");
  return p;
}

class DataHandle {
 public:
  static constexpr std::int64_t kDefaultCapacityPoints = 250000;  // the reference loader's buffer (data_io.hpp:17)

  explicit DataHandle(int device = 0, std::int64_t capacity_points = kDefaultCapacityPoints) {
    ThrowOnError(kmc_b200_handle_create(device, capacity_points, &handle_), "kmc_b200_handle_create");
  }
  ~DataHandle() { kmc_b200_handle_destroy(handle_); }

  DataHandle(DataHandle const&) = delete;
  DataHandle& operator=(DataHandle const&) = delete;

  int device() const { return kmc_b200_handle_device(handle_); }
  std::int64_t capacity() const { return kmc_b200_handle_capacity(handle_); }
  kmc_b200_handle* raw() const { return handle_; }

  // One scan held as interleaved float32 "x y z i" (the KITTI .bin layout); the azimuth -> time mapping is fused.
  void DeskewScan(const float* xyzi_in, float* xyzi_out, std::int64_t n_points, kmc_b200_frame_params const& params) {
    ThrowOnError(kmc_b200_deskew_frame_host(handle_, xyzi_in, xyzi_out, n_points, &params, KMC_B200_TIME_FROM_AZIMUTH),
                 "kmc_b200_deskew_frame_host");
  }

  // Same, but w carries each point's fraction of the trajectory (explicit per-point time stamps).
  void DeskewScanWithFractions(const float* xyzw_in, float* xyzw_out, std::int64_t n_points, kmc_b200_frame_params const& params) {
    ThrowOnError(kmc_b200_deskew_frame_host(handle_, xyzw_in, xyzw_out, n_points, &params, KMC_B200_TIME_FROM_W),
                 "kmc_b200_deskew_frame_host");
  }

  // A batch of scans stored back to back; frame_offsets has params.size() + 1 entries (points).
  void DeskewBatch(const float* xyzi_in, float* xyzi_out, std::vector<std::int64_t> const& frame_offsets,
                   std::vector<kmc_b200_frame_params> const& params) {
    if (frame_offsets.size() != params.size() + 1) throw std::invalid_argument("DeskewBatch: frame_offsets must have params.size() + 1 entries");
    ThrowOnError(kmc_b200_deskew_batch_host(handle_, xyzi_in, xyzi_out, frame_offsets.data(), params.data(),
                                            static_cast<std::int32_t>(params.size()), KMC_B200_TIME_FROM_AZIMUTH),
                 "kmc_b200_deskew_batch_host");
  }

  // KITTI .bin file in, motion-compensated .bin file out; returns the number of points.
  std::int64_t DeskewBinFile(std::string const& path_in, std::string const& path_out, kmc_b200_frame_params const& params) {
    std::int64_t n = 0;
    ThrowOnError(kmc_b200_deskew_bin_file(handle_, path_in.c_str(), path_out.c_str(), &params, &n), "kmc_b200_deskew_bin_file");
    return n;
  }

  // Many .bin files as one overlapped read -> H2D -> kernel -> D2H -> write pipeline (params[k] belongs to paths_in[k]);
  // returns the points per file.  Give the handle a capacity of several scans so that a staging slot holds a group.
  std::vector<std::int64_t> DeskewBinFiles(std::vector<std::string> const& paths_in, std::vector<std::string> const& paths_out,
                                           std::vector<kmc_b200_frame_params> const& params, int io_threads = 0) {
    if (paths_in.size() != params.size() || paths_out.size() != params.size())
      throw std::invalid_argument("DeskewBinFiles: paths_in, paths_out and params must have the same length");
    std::vector<const char*> in, out;
    for (auto const& p : paths_in) in.push_back(p.c_str());
    for (auto const& p : paths_out) out.push_back(p.c_str());
    std::vector<std::int64_t> points(params.size(), 0);
    ThrowOnError(kmc_b200_deskew_bin_files(handle_, static_cast<std::int32_t>(params.size()), in.data(), out.data(), params.data(),
                                           KMC_B200_TIME_FROM_AZIMUTH, io_threads, points.data()),
                 "kmc_b200_deskew_bin_files");
    return points;
  }

  // A whole KITTI raw run folder (the reference's MotionCompensateRun, handlers.cpp:41-65) with this handle's buffers.
  kmc_b200_run_stats MotionCompensateRun(std::string const& run_folder, int io_threads = 0) {
    kmc_b200_run_stats stats{};
    ThrowOnError(kmc_b200_motion_compensate_run(handle_, run_folder.c_str(), io_threads, &stats), "kmc_b200_motion_compensate_run");
    return stats;
  }

 private:
  kmc_b200_handle* handle_ = nullptr;
};

// Splits one batch over several handles (one per GPU), contiguous frame ranges, no collective.
inline void DeskewBatchMultiGpu(std::vector<DataHandle*> const& handles, const float* xyzi_in, float* xyzi_out,
                                std::vector<std::int64_t> const& frame_offsets, std::vector<kmc_b200_frame_params> const& params) {
  std::vector<kmc_b200_handle*> raw;
  for (DataHandle* h : handles) raw.push_back(h->raw());
  ThrowOnError(kmc_b200_deskew_batch_multi_gpu(raw.data(), static_cast<std::int32_t>(raw.size()), xyzi_in, xyzi_out, frame_offsets.data(),
                                               params.data(), static_cast<std::int32_t>(params.size()), KMC_B200_TIME_FROM_AZIMUTH),
               "kmc_b200_deskew_batch_multi_gpu");
}

}  // namespace kmc::b200
