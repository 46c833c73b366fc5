// kmc_internal.hpp — what the translation units of libkmc_b200.so share that is not part of the C ABI: error
// recording, CUDA status plumbing, device selection, NVTX ranges.  Nothing here is visible outside the library.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <cstring>
#include <string>

#include "kmc_b200.h"

// NVTX ranges around the host entry points (visible in Nsight Systems / Compute timelines); header-only, no cost when
// no tool is attached.
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
#define KMC_HAVE_NVTX 1
#else
#define KMC_HAVE_NVTX 0
#endif

namespace kmc_b200::internal {

// The calling thread's kmc_b200_last_error() text.
std::string& LastError();

// Records `what` as the calling thread's last error and returns `status`.
int SetError(int status, const std::string& what);

// Records the CUDA error (name + text, prefixed by `where`) and maps it to KMC_B200_ERR_NO_DEVICE / KMC_B200_ERR_CUDA.
int FailCuda(cudaError_t e, const char* where);

// Inside a catch (...) at the C boundary: records what was caught and returns KMC_B200_ERR_INTERNAL.  No exception may
// cross an extern "C" function (std::bad_alloc from a staging vector, std::system_error from thread creation ...).
int FailException(const char* where) noexcept;

constexpr int kMaxDevices = 64;  // device ordinals the per-device tables cover

// Multiprocessor count of `device`, cached.
int SmCount(int device, int* out);

inline bool ValidMode(int mode) { return mode == KMC_B200_TIME_FROM_AZIMUTH || mode == KMC_B200_TIME_FROM_W; }
inline bool Aligned(const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// Integer knob from KMC_B200_TUNE="key=value,key=value,..." (experiment switch of the sweep tools), else `fallback`.
inline int TuneValue(const char* key, int fallback) {
  const char* env = std::getenv("KMC_B200_TUNE");
  if (!env) return fallback;
  size_t const len = std::strlen(key);
  for (const char* p = env; *p;) {
    if (std::strncmp(p, key, len) == 0 && p[len] == '=') return std::atoi(p + len + 1);
    while (*p && *p != ',') ++p;
    if (*p == ',') ++p;
  }
  return fallback;
}

struct TraceRange {
  explicit TraceRange(const char* name) {
#if KMC_HAVE_NVTX
    nvtxRangePushA(name);
#else
    (void)name;
#endif
  }
  ~TraceRange() {
#if KMC_HAVE_NVTX
    nvtxRangePop();
#endif
  }
  TraceRange(TraceRange const&) = delete;
  TraceRange& operator=(TraceRange const&) = delete;
};

// Makes `device` current for the scope and restores the caller's device afterwards (handles may live on any device).
class DeviceGuard {
 public:
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&previous_) != cudaSuccess) previous_ = -1;
    status_ = (previous_ == device) ? cudaSuccess : cudaSetDevice(device);
    changed_ = (status_ == cudaSuccess && previous_ != device);
  }
  ~DeviceGuard() {
    if (changed_ && previous_ >= 0) cudaSetDevice(previous_);
  }
  DeviceGuard(DeviceGuard const&) = delete;
  DeviceGuard& operator=(DeviceGuard const&) = delete;
  cudaError_t status() const { return status_; }

 private:
  int previous_ = -1;
  bool changed_ = false;
  cudaError_t status_ = cudaSuccess;
};

}  // namespace kmc_b200::internal

// Closes a function-try-block of an extern "C" entry point:  int kmc_b200_x(...) try { ... } KMC_CATCH_AT_BOUNDARY("x")
#define KMC_CATCH_AT_BOUNDARY(name) \
  catch (...) { return ::kmc_b200::internal::FailException(name); }

#define KMC_CUDA_TRY(expr)                                                      \
  do {                                                                          \
    cudaError_t const e_ = (expr);                                              \
    if (e_ != cudaSuccess) return ::kmc_b200::internal::FailCuda(e_, #expr);    \
  } while (0)
