// kmc_capi.cu — the C ABI declared in include/kmc_b200.h, part 1: status / error text, per-frame host prep (double),
// and the DEVICE entry points (argument checking, then a kernel launch from kmc_kernels.cu).  The handle, the
// host<->device pipelines, the file pipelines and multi-GPU sharding are in kmc_pipeline.cu.  No CPU fallback lives
// here: every compute entry point ends in a kernel launch or an error status.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <exception>
#include <mutex>
#include <new>
#include <string>

#include "kmc_b200.h"
#include "kmc_host_math.hpp"
#include "kmc_internal.hpp"
#include "kmc_kernels.cuh"

namespace kmc_b200::internal {

namespace {
thread_local std::string t_last_error;
int g_sm_count[kMaxDevices];
std::once_flag g_sm_once[kMaxDevices];
}  // namespace

std::string& LastError() { return t_last_error; }

int SetError(int status, const std::string& what) {
  t_last_error = what;
  return status;
}

int FailCuda(cudaError_t e, const char* where) {
  t_last_error = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? KMC_B200_ERR_NO_DEVICE : KMC_B200_ERR_CUDA;
}

int FailException(const char* where) noexcept {
  try {
    try {
      throw;
    } catch (std::bad_alloc const&) {
      t_last_error = std::string(where) + ": out of host memory (std::bad_alloc)";
    } catch (std::exception const& e) {
      t_last_error = std::string(where) + ": C++ exception: " + e.what();
    } catch (...) {
      t_last_error = std::string(where) + ": unknown C++ exception";
    }
  } catch (...) {  // building the message failed as well
  }
  return KMC_B200_ERR_INTERNAL;
}

int SmCount(int device, int* out) {
  if (device < 0 || device >= kMaxDevices) return SetError(KMC_B200_ERR_NO_DEVICE, "device ordinal out of range");
  cudaError_t err = cudaSuccess;
  std::call_once(g_sm_once[device], [&] {
    int v = 0;
    err = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    g_sm_count[device] = (err == cudaSuccess) ? v : 0;
  });
  if (err != cudaSuccess) return FailCuda(err, "cudaDeviceGetAttribute(MultiProcessorCount)");
  if (g_sm_count[device] <= 0) return SetError(KMC_B200_ERR_NO_DEVICE, "device reports no multiprocessors");
  *out = g_sm_count[device];
  return KMC_B200_OK;
}

}  // namespace kmc_b200::internal

namespace {

using kmc_b200::internal::Aligned;
using kmc_b200::internal::FailCuda;
using kmc_b200::internal::SmCount;
using kmc_b200::internal::ValidMode;

int Fail(int status, const std::string& what) { return kmc_b200::internal::SetError(status, what); }

// counter-based generator of the synthetic frame twists (kmc_b200_synth_frame_params)
uint64_t Mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
double Uniform01(uint64_t& state) {
  state = Mix(state);
  return static_cast<double>(state >> 11) * (1.0 / 9007199254740992.0);
}
double Normal(uint64_t& state) {
  double u1 = Uniform01(state), u2 = Uniform01(state);
  if (u1 < 1e-300) u1 = 1e-300;
  return std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2);
}

}  // namespace

// =============================================================================================================
// C ABI
// =============================================================================================================
extern "C" {

int kmc_b200_version(void) { return KMC_B200_VERSION; }

const char* kmc_b200_status_string(int status) {
  switch (status) {
    case KMC_B200_OK: return "ok";
    case KMC_B200_ERR_NULL_POINTER: return "null pointer";
    case KMC_B200_ERR_BAD_SIZE: return "bad size, offsets or alignment";
    case KMC_B200_ERR_TIME_OUT_OF_RANGE: return "time outside the interpolation interval";
    case KMC_B200_ERR_EMPTY_INTERVAL: return "t_end <= t_start";
    case KMC_B200_ERR_NOT_RIGID: return "pose is not a finite rigid transform";
    case KMC_B200_ERR_CUDA: return "CUDA error";
    case KMC_B200_ERR_NO_DEVICE: return "no usable CUDA device";
    case KMC_B200_ERR_BAD_MODE: return "unknown time mode";
    case KMC_B200_ERR_CAPACITY: return "handle capacity exceeded";
    case KMC_B200_ERR_IO: return "file I/O error";
    case KMC_B200_ERR_INTERNAL: return "internal error (C++ exception caught at the C boundary)";
    case KMC_B200_WARN_ACCURACY: return "warning: frame outside the 1e-5 m accuracy domain of the fp32 kernels";
    default: return "unknown status";
  }
}

const char* kmc_b200_last_error(void) { return kmc_b200::internal::LastError().c_str(); }

int kmc_b200_device_count(void) {
  int n = 0;
  cudaError_t const e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return FailCuda(e, "cudaGetDeviceCount");
  return n;
}

uint64_t kmc_b200_launch_count(void) { return kmc_b200::dev::LaunchCount(); }

// ---- host prep ------------------------------------------------------------------------------------------------
namespace {
// include/kmc_b200.h "ACCURACY DOMAIN": ulp32(|p'|)/2 + 2.5e-7 |delta| + 5e-8 (|rho| + theta |p|) at the far end of the range.
double AccuracyBound(const kmc_b200_frame_params& P, double max_range) {
  double const theta = std::sqrt(static_cast<double>(P.theta2));
  double rho2 = 0.0;
  for (int i = 0; i < 3; ++i) rho2 += static_cast<double>(P.rho_par[i]) * P.rho_par[i] + static_cast<double>(P.rho_perp[i]) * P.rho_perp[i];
  double const rho = std::sqrt(rho2);
  double const s_max = std::max(static_cast<double>(P.x_req), 1.0 - static_cast<double>(P.x_req));
  double const angle = std::min(s_max * theta, M_PI);
  double const delta = s_max * rho + 2.0 * max_range * std::sin(0.5 * angle);  // translation + chord of the rotation
  double const reach = max_range + s_max * rho;  // a rotation does not move a point away from the sensor
  double const half_ulp = std::ldexp(1.0, std::ilogb(std::max(reach, 1e-30)) - 24);
  return half_ulp + 2.5e-7 * delta + 5e-8 * (rho + theta * max_range);
}
constexpr double kContractMetres = 1e-5;  // BASELINE.json north_star
constexpr double kContractRange = 120.0;  // the HDL-64E's reach; every coordinate then stays below 128 m (3.8e-6 m rounding)

int AccuracyStatus(const kmc_b200_frame_params& P) {
  if (AccuracyBound(P, kContractRange) <= kContractMetres) return KMC_B200_OK;
  return Fail(KMC_B200_WARN_ACCURACY, "frame constants are valid, but the motion per scan puts max |dxyz| < 1e-5 m out of reach of the fp32 "
                                      "kernels for points out to 120 m (kmc_b200_frame_accuracy_bound gives the bound)");
}
}  // namespace

int kmc_b200_frame_accuracy_bound(const kmc_b200_frame_params* params, double max_range_m, double* bound_m) {
  if (!params || !bound_m) return Fail(KMC_B200_ERR_NULL_POINTER, "frame_accuracy_bound: null argument");
  if (!(max_range_m >= 0.0) || !std::isfinite(max_range_m)) return Fail(KMC_B200_ERR_BAD_SIZE, "frame_accuracy_bound: max_range_m must be finite and >= 0");
  *bound_m = AccuracyBound(*params, max_range_m);
  return *bound_m <= kContractMetres ? KMC_B200_OK : KMC_B200_WARN_ACCURACY;
}

int kmc_b200_frame_params_from_twist(const double xi[6], double x_req, kmc_b200_frame_params* out) {
  if (!xi || !out) return Fail(KMC_B200_ERR_NULL_POINTER, "frame_params_from_twist: null argument");
  for (int i = 0; i < 6; ++i)
    if (!std::isfinite(xi[i])) return Fail(KMC_B200_ERR_NOT_RIGID, "frame_params_from_twist: non-finite twist");
  if (!(x_req >= 0.0 && x_req <= 1.0))
    return Fail(KMC_B200_ERR_TIME_OUT_OF_RANGE, "requested fraction outside [0, 1] (reference asserts, trajectory_interpolation.cpp:32)");
  kmc_b200::host::FrameParamsFromTwist(xi, x_req, out);
  return AccuracyStatus(*out);
}

int kmc_b200_frame_params_from_poses(const double T_start[16], const double T_end[16], double t_start, double t_end,
                                     double t_req, kmc_b200_frame_params* out) {
  if (!T_start || !T_end || !out) return Fail(KMC_B200_ERR_NULL_POINTER, "frame_params_from_poses: null argument");
  if (!(t_end > t_start)) return Fail(KMC_B200_ERR_EMPTY_INTERVAL, "t_end <= t_start (or NaN)");
  if (!(t_req >= t_start && t_req <= t_end))
    return Fail(KMC_B200_ERR_TIME_OUT_OF_RANGE, "requested time outside [t_start, t_end] (reference asserts, trajectory_interpolation.cpp:32)");
  double xi[6];
  if (!kmc_b200::host::RelativeTwist(T_start, T_end, xi))
    return Fail(KMC_B200_ERR_NOT_RIGID, "T_start^-1 T_end has no proper-rotation polar factor");
  kmc_b200::host::FrameParamsFromTwist(xi, (t_req - t_start) / (t_end - t_start), out);
  return AccuracyStatus(*out);
}

#define KMC_NULLCHECK2(a, b) \
  if (!(a) || !(b)) return Fail(KMC_B200_ERR_NULL_POINTER, std::string(__func__) + ": null argument")

int kmc_b200_so3_hat(const double phi[3], double out[9]) { KMC_NULLCHECK2(phi, out); kmc_b200::host::So3Hat(phi, out); return KMC_B200_OK; }
int kmc_b200_so3_vee(const double m[9], double out[3]) { KMC_NULLCHECK2(m, out); kmc_b200::host::So3Vee(m, out); return KMC_B200_OK; }
int kmc_b200_so3_exp(const double phi[3], double out[9]) { KMC_NULLCHECK2(phi, out); kmc_b200::host::So3Exp(phi, out); return KMC_B200_OK; }
int kmc_b200_so3_log(const double R[9], double out[3]) { KMC_NULLCHECK2(R, out); kmc_b200::host::So3Log(R, out); return KMC_B200_OK; }
int kmc_b200_so3_left_jacobian(const double phi[3], double out[9]) { KMC_NULLCHECK2(phi, out); kmc_b200::host::So3LeftJacobian(phi, out); return KMC_B200_OK; }
int kmc_b200_so3_inverse_left_jacobian(const double phi[3], double out[9]) { KMC_NULLCHECK2(phi, out); kmc_b200::host::So3InverseLeftJacobian(phi, out); return KMC_B200_OK; }
int kmc_b200_se3_exp(const double xi[6], double T[16]) { KMC_NULLCHECK2(xi, T); kmc_b200::host::Se3Exp(xi, T); return KMC_B200_OK; }
int kmc_b200_se3_log(const double T[16], double xi[6]) {
  KMC_NULLCHECK2(T, xi);
  if (!kmc_b200::host::Se3Log(T, xi)) return Fail(KMC_B200_ERR_NOT_RIGID, "se3_log: linear block has no proper-rotation polar factor");
  return KMC_B200_OK;
}

int kmc_b200_pose_at_time(double t1, const double P1[16], double t2, const double P2[16], double t, double out[16]) {
  if (!P1 || !P2 || !out) return Fail(KMC_B200_ERR_NULL_POINTER, "pose_at_time: null argument");
  int const rc = kmc_b200::host::PoseAtTime(t1, P1, t2, P2, t, out);
  if (rc != KMC_B200_OK) return Fail(rc, std::string("pose_at_time: ") + kmc_b200_status_string(rc));
  return rc;
}

int kmc_b200_relative_pose_between_times(double t1, const double P1[16], double t2, const double P2[16], double anchor,
                                         double query, double out[16]) {
  if (!P1 || !P2 || !out) return Fail(KMC_B200_ERR_NULL_POINTER, "relative_pose_between_times: null argument");
  double a[16], q[16], ainv[16];
  int rc = kmc_b200::host::PoseAtTime(t1, P1, t2, P2, anchor, a);
  if (rc == KMC_B200_OK) rc = kmc_b200::host::PoseAtTime(t1, P1, t2, P2, query, q);
  if (rc != KMC_B200_OK) return Fail(rc, std::string("relative_pose_between_times: ") + kmc_b200_status_string(rc));
  if (!kmc_b200::host::AffineInverse(a, ainv)) return Fail(KMC_B200_ERR_NOT_RIGID, "anchor pose is singular");
  kmc_b200::host::AffineMul(ainv, q, out);
  return KMC_B200_OK;
}

double kmc_b200_fraction_of_scan_completed(double x, double y) { return kmc_b200::host::FractionOfScanCompleted(x, y); }
double kmc_b200_pseudo_time_stamp(double x, double y, double scan_start, double scan_end) {
  return scan_start + kmc_b200::host::FractionOfScanCompleted(x, y) * (scan_end - scan_start);
}

int kmc_b200_oxts_to_pose(double lat, double lon, double alt, double roll, double pitch, double yaw, double scale, double T[16]) {
  if (!T) return Fail(KMC_B200_ERR_NULL_POINTER, "oxts_to_pose: null output");
  kmc_b200::host::OxtsToPose(lat, lon, alt, roll, pitch, yaw, scale, T);
  return KMC_B200_OK;
}

int kmc_b200_camera_params_from_calibration(const double P_rect[12], const double R_rect_00[9], const double T_velo_to_cam[16],
                                            double max_range, kmc_b200_camera_params* out) {
  if (!P_rect || !R_rect_00 || !T_velo_to_cam || !out) return Fail(KMC_B200_ERR_NULL_POINTER, "camera_params_from_calibration: null argument");
  if (!(max_range > 0.01)) return Fail(KMC_B200_ERR_BAD_SIZE, "camera_params_from_calibration: max_range must exceed 0.01 m");
  kmc_b200::host::CameraParamsFromCalibration(P_rect, R_rect_00, T_velo_to_cam, max_range, out);
  return KMC_B200_OK;
}

int kmc_b200_shard_range(int64_t n_items, int32_t n_parts, int32_t index, int64_t* begin, int64_t* end) {
  if (!begin || !end) return Fail(KMC_B200_ERR_NULL_POINTER, "shard_range: null argument");
  if (n_items < 0 || n_parts <= 0 || index < 0 || index >= n_parts) return Fail(KMC_B200_ERR_BAD_SIZE, "shard_range: bad arguments");
  // the first (n_items % n_parts) parts get one extra item
  int64_t const base = n_items / n_parts, extra = n_items % n_parts;
  *begin = index * base + std::min<int64_t>(index, extra);
  *end = *begin + base + (index < extra ? 1 : 0);
  return KMC_B200_OK;
}

// ---- device entry points ------------------------------------------------------------------------------------------
int kmc_b200_deskew_frame_device(const float* in, float* out, int64_t n, const kmc_b200_frame_params* params, int mode,
                                 void* stream) {
  if (n < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_frame_device: negative n_points");
  if (!ValidMode(mode)) return Fail(KMC_B200_ERR_BAD_MODE, "deskew_frame_device: unknown time mode");
  if (!params) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_frame_device: null params");
  if (n == 0) return KMC_B200_OK;
  if (!in || !out) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_frame_device: null point buffer");
  if (!Aligned(in, 16) || !Aligned(out, 16)) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_frame_device: buffers must be 16-byte aligned");
  int device = 0, sm = 0;
  KMC_CUDA_TRY(cudaGetDevice(&device));
  if (int rc = SmCount(device, &sm)) return rc;
  auto const cfg = kmc_b200::dev::PickConfig(n, Aligned(in, 32) && Aligned(out, 32), in == out, sm, false);
  KMC_CUDA_TRY(kmc_b200::dev::LaunchDeskewFrame(in, out, n, *params, mode, cfg, sm, static_cast<cudaStream_t>(stream)));
  return KMC_B200_OK;
}

int kmc_b200_deskew_batch_device(const float* in, float* out, const int64_t* offsets_dev, const kmc_b200_frame_params* params_dev,
                                 int32_t n_frames, int64_t n_total, int mode, void* stream) {
  if (n_frames < 0 || n_total < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_batch_device: negative size");
  if (!ValidMode(mode)) return Fail(KMC_B200_ERR_BAD_MODE, "deskew_batch_device: unknown time mode");
  if (n_frames == 0 || n_total == 0) return KMC_B200_OK;
  if (!in || !out || !offsets_dev || !params_dev) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_batch_device: null argument");
  if (!Aligned(in, 16) || !Aligned(out, 16) || !Aligned(params_dev, 16) || !Aligned(offsets_dev, 8))
    return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_batch_device: misaligned buffer");
  int device = 0, sm = 0;
  KMC_CUDA_TRY(cudaGetDevice(&device));
  if (int rc = SmCount(device, &sm)) return rc;
  auto const cfg = kmc_b200::dev::PickConfig(n_total, Aligned(in, 32) && Aligned(out, 32), in == out, sm, true);
  KMC_CUDA_TRY(kmc_b200::dev::LaunchDeskewBatch(in, out, offsets_dev, params_dev, n_frames, n_total, 0, n_total, mode, cfg, sm,
                                                static_cast<cudaStream_t>(stream)));
  return KMC_B200_OK;
}

namespace {
int ProjectCommon(const char* who, const float* in, float* cloud_out, float* pix_out, int64_t n, const kmc_b200_frame_params* params,
                  const kmc_b200_camera_params* camera, int mode, void* stream) {
  if (n < 0) return Fail(KMC_B200_ERR_BAD_SIZE, std::string(who) + ": negative n_points");
  if (!ValidMode(mode)) return Fail(KMC_B200_ERR_BAD_MODE, std::string(who) + ": unknown time mode");
  if (!camera) return Fail(KMC_B200_ERR_NULL_POINTER, std::string(who) + ": null camera params");
  if (n == 0) return KMC_B200_OK;
  if (!in || !pix_out) return Fail(KMC_B200_ERR_NULL_POINTER, std::string(who) + ": null point buffer");
  if (!Aligned(in, 16) || !Aligned(pix_out, 16) || (cloud_out && !Aligned(cloud_out, 16)))
    return Fail(KMC_B200_ERR_BAD_SIZE, std::string(who) + ": buffers must be 16-byte aligned");
  if (pix_out == in || pix_out == cloud_out) return Fail(KMC_B200_ERR_BAD_SIZE, std::string(who) + ": the pixel buffer must not alias the clouds");
  int device = 0, sm = 0;
  KMC_CUDA_TRY(cudaGetDevice(&device));
  if (int rc = SmCount(device, &sm)) return rc;
  bool const vec2 = Aligned(in, 32) && Aligned(pix_out, 32) && (!cloud_out || Aligned(cloud_out, 32));
  KMC_CUDA_TRY(kmc_b200::dev::LaunchProject(in, cloud_out, pix_out, n, params, *camera, mode, vec2, sm, static_cast<cudaStream_t>(stream)));
  return KMC_B200_OK;
}
}  // namespace

int kmc_b200_project_frame_device(const float* in, float* uvzc_out, int64_t n, const kmc_b200_camera_params* camera, void* stream) {
  return ProjectCommon("project_frame_device", in, nullptr, uvzc_out, n, nullptr, camera, KMC_B200_TIME_FROM_AZIMUTH, stream);
}

int kmc_b200_deskew_project_frame_device(const float* in, float* xyzi_out, float* uvzc_out, int64_t n, const kmc_b200_frame_params* params,
                                         const kmc_b200_camera_params* camera, int mode, void* stream) {
  if (!params) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_project_frame_device: null frame params");
  return ProjectCommon("deskew_project_frame_device", in, xyzi_out, uvzc_out, n, params, camera, mode, stream);
}

int kmc_b200_deskew_cloud_f64_device(const double* cloud, const double* stamps, double* out, int64_t n, double t_start, double t_end,
                                     double t_req, const kmc_b200_frame_params* params, int* flags_dev, void* stream) {
  if (n < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_cloud_f64_device: negative n_points");
  if (!params) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_cloud_f64_device: null params");
  if (!(t_end > t_start)) return Fail(KMC_B200_ERR_EMPTY_INTERVAL, "deskew_cloud_f64_device: t_end <= t_start");
  if (!(t_req >= t_start && t_req <= t_end)) return Fail(KMC_B200_ERR_TIME_OUT_OF_RANGE, "deskew_cloud_f64_device: requested time outside [t_start, t_end]");
  if (n == 0) return KMC_B200_OK;
  if (!cloud || !stamps || !out || !flags_dev) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_cloud_f64_device: null buffer");
  if (!Aligned(cloud, 8) || !Aligned(stamps, 8) || !Aligned(out, 8) || !Aligned(flags_dev, 4))
    return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_cloud_f64_device: misaligned buffer");
  int device = 0, sm = 0;
  KMC_CUDA_TRY(cudaGetDevice(&device));
  if (int rc = SmCount(device, &sm)) return rc;
  KMC_CUDA_TRY(kmc_b200::dev::LaunchDeskewCloudF64(cloud, stamps, out, n, t_start, t_end, (t_req - t_start) / (t_end - t_start), *params,
                                                   flags_dev, sm, static_cast<cudaStream_t>(stream)));
  return KMC_B200_OK;
}

int kmc_b200_deskew_project_frame4_device(const float* in, float* xyzi_out, float* const uvzc_out[4], int64_t n,
                                          const kmc_b200_frame_params* params, const kmc_b200_camera_params cameras[4], int mode,
                                          void* stream) {
  if (n < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_project_frame4_device: negative n_points");
  if (!ValidMode(mode)) return Fail(KMC_B200_ERR_BAD_MODE, "deskew_project_frame4_device: unknown time mode");
  if (!cameras || !uvzc_out) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_project_frame4_device: null camera / output table");
  if (n == 0) return KMC_B200_OK;
  if (!in) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_project_frame4_device: null point buffer");
  if (!Aligned(in, 16) || (xyzi_out && !Aligned(xyzi_out, 16))) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_project_frame4_device: misaligned cloud");
  for (int c = 0; c < 4; ++c) {
    if (!uvzc_out[c]) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_project_frame4_device: null pixel buffer");
    if (!Aligned(uvzc_out[c], 16)) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_project_frame4_device: misaligned pixel buffer");
    if (uvzc_out[c] == in || uvzc_out[c] == xyzi_out) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_project_frame4_device: a pixel buffer aliases a cloud");
    for (int d = 0; d < c; ++d)
      if (uvzc_out[c] == uvzc_out[d]) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_project_frame4_device: pixel buffers must be distinct");
  }
  int device = 0, sm = 0;
  KMC_CUDA_TRY(cudaGetDevice(&device));
  if (int rc = SmCount(device, &sm)) return rc;
  KMC_CUDA_TRY(kmc_b200::dev::LaunchProject4(in, xyzi_out, uvzc_out, n, params, cameras, mode, sm, static_cast<cudaStream_t>(stream)));
  return KMC_B200_OK;
}

int kmc_b200_deskew_project_batch_device(const float* in, float* xyzi_out, float* const uvzc_out[], int32_t n_cameras,
                                         const int64_t* offsets_dev, const kmc_b200_frame_params* params_dev, int32_t n_frames,
                                         int64_t n_total, const kmc_b200_camera_params* cameras, int mode, void* stream) {
  if (n_frames < 0 || n_total < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_project_batch_device: negative size");
  if (n_cameras != 1 && n_cameras != 4) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_project_batch_device: n_cameras must be 1 or 4");
  if (!ValidMode(mode)) return Fail(KMC_B200_ERR_BAD_MODE, "deskew_project_batch_device: unknown time mode");
  if (!cameras || !uvzc_out) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_project_batch_device: null camera / output table");
  if (n_frames == 0 || n_total == 0) return KMC_B200_OK;
  if (!in || !offsets_dev || !params_dev) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_project_batch_device: null argument");
  if (!Aligned(in, 16) || (xyzi_out && !Aligned(xyzi_out, 16)) || !Aligned(params_dev, 16) || !Aligned(offsets_dev, 8))
    return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_project_batch_device: misaligned buffer");
  for (int c = 0; c < n_cameras; ++c) {
    if (!uvzc_out[c]) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_project_batch_device: null pixel buffer");
    if (!Aligned(uvzc_out[c], 16)) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_project_batch_device: misaligned pixel buffer");
    if (uvzc_out[c] == in || uvzc_out[c] == xyzi_out) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_project_batch_device: a pixel buffer aliases a cloud");
    for (int d = 0; d < c; ++d)
      if (uvzc_out[c] == uvzc_out[d]) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_project_batch_device: pixel buffers must be distinct");
  }
  int device = 0, sm = 0;
  KMC_CUDA_TRY(cudaGetDevice(&device));
  if (int rc = SmCount(device, &sm)) return rc;
  KMC_CUDA_TRY(kmc_b200::dev::LaunchDeskewProjectBatch(in, xyzi_out, uvzc_out, n_cameras, offsets_dev, params_dev, n_frames, n_total, cameras, mode,
                                                       sm, static_cast<cudaStream_t>(stream)));
  return KMC_B200_OK;
}

int kmc_b200_deskew_cloud_f64_batch_device(const double* cloud, const double* stamps, double* out, const int64_t* offsets_dev,
                                           const kmc_b200_frame_params* params_dev, const double* times_dev, int32_t n_frames,
                                           int64_t n_total, int* flags_dev, void* stream) {
  if (n_frames < 0 || n_total < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_cloud_f64_batch_device: negative size");
  if (n_frames == 0) return KMC_B200_OK;
  if (!offsets_dev || !params_dev || !times_dev || !flags_dev) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_cloud_f64_batch_device: null table");
  if (n_total > 0 && (!cloud || !stamps || !out)) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_cloud_f64_batch_device: null buffer");
  if (!Aligned(cloud, 8) || !Aligned(stamps, 8) || !Aligned(out, 8) || !Aligned(times_dev, 8) || !Aligned(offsets_dev, 8) ||
      !Aligned(params_dev, 16) || !Aligned(flags_dev, 4))
    return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_cloud_f64_batch_device: misaligned buffer");
  int device = 0, sm = 0;
  KMC_CUDA_TRY(cudaGetDevice(&device));
  if (int rc = SmCount(device, &sm)) return rc;
  KMC_CUDA_TRY(kmc_b200::dev::LaunchDeskewCloudF64Batch(cloud, stamps, out, offsets_dev, params_dev, times_dev, n_frames, n_total, flags_dev, sm,
                                                        static_cast<cudaStream_t>(stream)));
  return KMC_B200_OK;
}

int kmc_b200_pseudo_time_stamps_device(const float* in, double* stamps, int64_t n, double start, double end, void* stream) {
  if (n < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "pseudo_time_stamps_device: negative n_points");
  if (n == 0) return KMC_B200_OK;
  if (!in || !stamps) return Fail(KMC_B200_ERR_NULL_POINTER, "pseudo_time_stamps_device: null argument");
  if (!Aligned(in, 16) || !Aligned(stamps, 8)) return Fail(KMC_B200_ERR_BAD_SIZE, "pseudo_time_stamps_device: misaligned buffer");
  int device = 0, sm = 0;
  KMC_CUDA_TRY(cudaGetDevice(&device));
  if (int rc = SmCount(device, &sm)) return rc;
  KMC_CUDA_TRY(kmc_b200::dev::LaunchPseudoTimeStamps(in, stamps, n, start, end, sm, static_cast<cudaStream_t>(stream)));
  return KMC_B200_OK;
}

int kmc_b200_pseudo_time_stamps_xy_device(const double* x, const double* y, double* stamps, int64_t n, double start, double end,
                                          void* stream) {
  if (n < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "pseudo_time_stamps_xy_device: negative n_points");
  if (n == 0) return KMC_B200_OK;
  if (!x || !y || !stamps) return Fail(KMC_B200_ERR_NULL_POINTER, "pseudo_time_stamps_xy_device: null argument");
  if (!Aligned(x, 8) || !Aligned(y, 8) || !Aligned(stamps, 8)) return Fail(KMC_B200_ERR_BAD_SIZE, "pseudo_time_stamps_xy_device: misaligned buffer");
  int device = 0, sm = 0;
  KMC_CUDA_TRY(cudaGetDevice(&device));
  if (int rc = SmCount(device, &sm)) return rc;
  KMC_CUDA_TRY(kmc_b200::dev::LaunchPseudoTimeStampsXy(x, y, stamps, n, start, end, sm, static_cast<cudaStream_t>(stream)));
  return KMC_B200_OK;
}

int kmc_b200_check_fractions_device(const float* xyzi, int64_t n, int* flags_dev, void* stream) {
  if (n < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "check_fractions_device: negative n_points");
  if (!flags_dev || (n > 0 && !xyzi)) return Fail(KMC_B200_ERR_NULL_POINTER, "check_fractions_device: null argument");
  if (!Aligned(xyzi, 16) || !Aligned(flags_dev, 4)) return Fail(KMC_B200_ERR_BAD_SIZE, "check_fractions_device: misaligned buffer");
  int device = 0, sm = 0;
  KMC_CUDA_TRY(cudaGetDevice(&device));
  if (int rc = SmCount(device, &sm)) return rc;
  KMC_CUDA_TRY(kmc_b200::dev::LaunchCheckFractions(xyzi, n, flags_dev, sm, static_cast<cudaStream_t>(stream)));
  return KMC_B200_OK;
}

int kmc_b200_frame_checksums_device(const float* xyzi, const int64_t* offsets_dev, int32_t n_frames, int64_t n_total, uint64_t* sums_dev,
                                    void* stream) {
  if (n_frames < 0 || n_total < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "frame_checksums_device: negative size");
  if (n_frames == 0) return KMC_B200_OK;
  if (!offsets_dev || !sums_dev || (n_total > 0 && !xyzi)) return Fail(KMC_B200_ERR_NULL_POINTER, "frame_checksums_device: null argument");
  if (!Aligned(xyzi, 16) || !Aligned(offsets_dev, 8) || !Aligned(sums_dev, 8)) return Fail(KMC_B200_ERR_BAD_SIZE, "frame_checksums_device: misaligned buffer");
  int device = 0, sm = 0;
  KMC_CUDA_TRY(cudaGetDevice(&device));
  if (int rc = SmCount(device, &sm)) return rc;
  KMC_CUDA_TRY(kmc_b200::dev::LaunchFrameChecksums(xyzi, offsets_dev, n_frames, n_total, sums_dev, sm, static_cast<cudaStream_t>(stream)));
  return KMC_B200_OK;
}

int kmc_b200_synth_scans_device(float* out, int64_t points_per_scan, int32_t n_scans, int32_t n_rings, uint64_t seed,
                                int64_t first_scan_index, void* stream) {
  if (points_per_scan < 0 || n_scans < 0 || n_rings < 2) return Fail(KMC_B200_ERR_BAD_SIZE, "synth_scans_device: bad size");
  if (points_per_scan == 0 || n_scans == 0) return KMC_B200_OK;
  if (!out) return Fail(KMC_B200_ERR_NULL_POINTER, "synth_scans_device: null output");
  if (!Aligned(out, 16)) return Fail(KMC_B200_ERR_BAD_SIZE, "synth_scans_device: misaligned buffer");
  int device = 0, sm = 0;
  KMC_CUDA_TRY(cudaGetDevice(&device));
  if (int rc = SmCount(device, &sm)) return rc;
  KMC_CUDA_TRY(kmc_b200::dev::LaunchSynthScans(out, points_per_scan, n_scans, n_rings, seed, first_scan_index, sm,
                                               static_cast<cudaStream_t>(stream)));
  return KMC_B200_OK;
}

int kmc_b200_synth_frame_params(int32_t n_frames, uint64_t seed, int64_t first_scan_index, double x_req,
                                kmc_b200_frame_params* params_out, double* xi_out) {
  if (n_frames < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "synth_frame_params: negative n_frames");
  if (!params_out && n_frames > 0) return Fail(KMC_B200_ERR_NULL_POINTER, "synth_frame_params: null output");
  if (!(x_req >= 0.0 && x_req <= 1.0)) return Fail(KMC_B200_ERR_TIME_OUT_OF_RANGE, "synth_frame_params: x_req outside [0,1]");
  for (int32_t k = 0; k < n_frames; ++k) {
    uint64_t state = Mix(seed + static_cast<uint64_t>(first_scan_index + k)) ^ 0x5DEECE66Dull;
    double xi[6];
    xi[0] = 3.0 * Uniform01(state);   // forward motion per scan, up to 30 m/s
    xi[1] = 0.05 * Normal(state);
    xi[2] = 0.02 * Normal(state);
    xi[3] = 0.003 * Normal(state);    // roll
    xi[4] = 0.004 * Normal(state);    // pitch
    xi[5] = 0.05 * Normal(state);     // yaw per scan (0.5 rad/s sigma)
    kmc_b200::host::FrameParamsFromTwist(xi, x_req, params_out + k);
    if (xi_out) std::memcpy(xi_out + 6 * k, xi, sizeof(xi));
  }
  return KMC_B200_OK;
}

}  // extern "C"
