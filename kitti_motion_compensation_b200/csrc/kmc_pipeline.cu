// kmc_pipeline.cu — the C ABI declared in include/kmc_b200.h, part 2: the handle (device, streams, pinned + device
// staging slots) and everything that moves data through it — host-buffer entry points (three-slot H2D / kernel / D2H
// pipeline), multi-GPU frame sharding, and the KITTI .bin file pipelines (overlapped pread / copies / kernel / pwrite).
// Every compute entry point ends in a kernel launch from kmc_kernels.cu or an error status; there is no CPU fallback.
#include <cuda_runtime.h>

#include <fcntl.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kmc_b200.h"
#include "kmc_host_pool.hpp"
#include "kmc_internal.hpp"
#include "kmc_kernels.cuh"

namespace {

using kmc_b200::internal::Aligned;
using kmc_b200::internal::DeviceGuard;
using kmc_b200::internal::FailCuda;
using kmc_b200::internal::HostPool;
using kmc_b200::internal::kMaxDevices;
using kmc_b200::internal::SmCount;
using kmc_b200::internal::TraceRange;
using kmc_b200::internal::TuneValue;
using kmc_b200::internal::ValidMode;

int Fail(int status, const std::string& what) { return kmc_b200::internal::SetError(status, what); }
using kmc_b200::internal::LastError;

bool IsPinnedHost(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();  // clear
    return false;
  }
  return attr.type == cudaMemoryTypeHost;
}

}  // namespace

// =============================================================================================================
// handle
// =============================================================================================================
struct kmc_b200_handle {
  static constexpr int kSlots = 3;
  int device = -1;
  int sm_count = 0;
  int64_t capacity = 0;  // points per staging slot
  cudaStream_t stream[kSlots] = {};
  cudaEvent_t done[kSlots] = {};
  float* d_in[kSlots] = {};
  float* d_out[kSlots] = {};
  float* h_in[kSlots] = {};   // pinned
  float* h_out[kSlots] = {};  // pinned
  // device copies of the batch tables
  int64_t* d_offsets = nullptr;
  kmc_b200_frame_params* d_params = nullptr;
  int64_t table_capacity = 0;  // frames
  // narrow-transport staging of kmc_b200_deskew_cloud_f64_host: per slot 5 float columns up and 3 down, pinned + device
  float* f64_pinned = nullptr;
  float* f64_device = nullptr;
  int64_t f64_chunk = 0;  // points per column (multiple of 4)
  // kmc_b200_handle_set_file_callback: called after every output file of kmc_b200_deskew_bin_files has been written
  kmc_b200_file_done_fn file_done = nullptr;
  void* file_done_user = nullptr;
  std::mutex mu;
};

namespace {

void FreeHandle(kmc_b200_handle* h) {
  if (!h) return;
  DeviceGuard const guard(h->device >= 0 ? h->device : 0);
  for (int s = 0; s < kmc_b200_handle::kSlots; ++s) {
    if (h->d_in[s]) cudaFree(h->d_in[s]);
    if (h->d_out[s]) cudaFree(h->d_out[s]);
    if (h->h_in[s]) cudaFreeHost(h->h_in[s]);
    if (h->h_out[s]) cudaFreeHost(h->h_out[s]);
    if (h->done[s]) cudaEventDestroy(h->done[s]);
    if (h->stream[s]) cudaStreamDestroy(h->stream[s]);
  }
  if (h->d_offsets) cudaFree(h->d_offsets);
  if (h->d_params) cudaFree(h->d_params);
  if (h->f64_pinned) cudaFreeHost(h->f64_pinned);
  if (h->f64_device) cudaFree(h->f64_device);
  delete h;
}

int EnsureTables(kmc_b200_handle* h, int64_t n_frames) {
  if (n_frames <= h->table_capacity) return KMC_B200_OK;
  if (h->d_offsets) cudaFree(h->d_offsets);
  if (h->d_params) cudaFree(h->d_params);
  h->d_offsets = nullptr;
  h->d_params = nullptr;
  h->table_capacity = 0;
  int64_t const cap = std::max<int64_t>(n_frames, 1024);
  KMC_CUDA_TRY(cudaMalloc(&h->d_offsets, static_cast<size_t>(cap + 1) * sizeof(int64_t)));
  KMC_CUDA_TRY(cudaMalloc(&h->d_params, static_cast<size_t>(cap) * sizeof(kmc_b200_frame_params)));
  h->table_capacity = cap;
  return KMC_B200_OK;
}

// Core of every host pipeline: chunks of the point range [0, n) rotate through the handle's three stream / buffer slots
//   fetch(slot, first, count) -> pinned source of the chunk (the caller's pinned memory, or h_in[slot] after a staging
//                                copy or a pread)                                                        [host, this thread]
//   H2D -> launch(slot, first, count, d_in, d_out) -> D2H into sink(slot, first) (caller's pinned memory or h_out[slot])
//                                                                                                        [slot's stream]
//   deliver(slot, first, count) once the chunk's event has fired (unstage / pwrite; nothing for pinned)  [host, this thread]
// so the copy engines, the SMs and the host work of neighbouring chunks overlap.
// zero_copy: no copy engines — the kernel reads the pinned source and writes the pinned sink directly over PCIe (one launch
// per chunk).  For a transfer of a few MB the two copies' fixed latencies (~10 us each) and their serialisation behind the
// kernel cost more than SM-issued PCIe reads lose: one 130 000-point scan takes 67 us instead of 96 us
// (profiles/r02_latency_probe.log), a 2 M-point frame 881 us instead of 978 us (profiles/r02_sweep_single_scan.log); for long
// streams the copy engines are at least as fast (36-47 vs 36-40 GB/s each way, box dependent) and leave the SMs free.
template <class Fetch, class Launch, class Sink, class Deliver>
int StreamChunksImpl(kmc_b200_handle* h, int64_t n, bool zero_copy, Fetch&& fetch, Launch&& launch, Sink&& sink, Deliver&& deliver);

// On any failure the slots' streams are drained before returning, so that no copy is still reading or writing the
// caller's buffers after the call has reported an error.
template <class Fetch, class Launch, class Sink, class Deliver>
int StreamChunks(kmc_b200_handle* h, int64_t n, bool zero_copy, Fetch&& fetch, Launch&& launch, Sink&& sink, Deliver&& deliver) {
  int const rc = StreamChunksImpl(h, n, zero_copy, fetch, launch, sink, deliver);
  if (rc != KMC_B200_OK) {
    std::string const keep = LastError();
    for (int s = 0; s < kmc_b200_handle::kSlots; ++s) cudaStreamSynchronize(h->stream[s]);
    cudaGetLastError();
    LastError() = keep;
  }
  return rc;
}

// The library's host threads: ONE pool per process, min(12, 3/4 of the hardware threads) - 1 workers beside the calling thread
// (KMC_B200_TUNE host_threads=N), created on first use and never torn down.  Handles share it and take turns: the host
// passes of concurrent callers run one after the other at full width while their copies and kernels overlap on the
// devices — a pool per handle oversubscribes the cores as soon as a few threads call at once (8 callers x 8 threads on a
// 16-thread box: 3x slower in aggregate than one caller, profiles/r02_sweep_dropin.log).
class SharedPool {
 public:
  template <class F>
  static void Run(int64_t n_blocks, F&& fn) {
    SharedPool& self = Instance();
    // a turn lasts tens of microseconds (one pass over one chunk): spin for it before sleeping on the mutex — a futex
    // sleep / wake pair costs about as much as the turn itself, and four concurrent MotionCompensateFrame callers ran at
    // half the single caller's rate with a plain lock (profiles/r02_dropin_concurrent_callers_shared_pool.log)
    std::unique_lock<std::mutex> lock(self.turn_, std::try_to_lock);
    if (!lock.owns_lock()) {
      auto const give_up = std::chrono::steady_clock::now() + std::chrono::microseconds(300);
      for (int spins = 1; !lock.try_lock(); ++spins) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((spins & 63) == 0) {
          if (std::chrono::steady_clock::now() > give_up) {
            lock.lock();
            break;
          }
          std::this_thread::yield();  // with as many callers as cores the pool's workers need the cores more than the spinners do
        }
      }
    }
    self.pool_.Run(n_blocks, std::forward<F>(fn));
  }

 private:
  static SharedPool& Instance() {
    static SharedPool* const self = new SharedPool;  // leaked on purpose: worker threads must not be joined during static teardown
    return *self;
  }
  static int Workers() {
    unsigned const hw = std::max(1u, std::thread::hardware_concurrency());
    int const dflt = static_cast<int>(std::min(12u, std::max(2u, 3 * hw / 4)));  // 1 / 2 / 4 / 8 / 12 threads: 570 / 328 / 169 / 137 / 128 us per KITTI frame
    return std::min(std::max(TuneValue("host_threads", dflt) - 1, 0), 63);
  }
  SharedPool() : pool_(Workers()) {}
  HostPool pool_;
  std::mutex turn_;
};

// A copy whose destination will not be read by this core again (a staging slot the DMA engine reads next, or the caller's
// result buffer): non-temporal stores skip the read-for-ownership of the destination lines, i.e. 2 instead of 3 bytes of
// memory traffic per byte copied.  A long host stream through pageable buffers is bound by exactly that traffic (stage +
// DMA + DMA + unstage); glibc's memcpy only switches to streaming stores far above the 256 KB blocks used here.
void StreamingCopy(void* dst, const void* src, size_t bytes) {
#if defined(__SSE2__)
  auto* d = static_cast<char*>(dst);
  auto const* s = static_cast<const char*>(src);
  size_t const head = std::min(bytes, (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15);
  if (head) std::memcpy(d, s, head);
  size_t i = head;
  for (; i + 64 <= bytes; i += 64) {
    __m128i const a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
    __m128i const b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
    __m128i const c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
    __m128i const e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
  }
  _mm_sfence();
  if (i < bytes) std::memcpy(d + i, s + i, bytes - i);
#else
  std::memcpy(dst, src, bytes);
#endif
}

// Pageable caller memory is staged through the handle's pinned slots.  One thread copies about 10-15 GB/s, a quarter of
// what the PCIe link moves, so copies of 512 KB and more are cut into 256 KB blocks for the library's host threads; long
// transfers (>= 8 MB, which no cache will hold until they are read again) use streaming stores.
void StagingCopy(kmc_b200_handle*, void* dst, const void* src, size_t bytes) {
  constexpr size_t kBlock = size_t{256} << 10;
  if (bytes < 2 * kBlock) {
    std::memcpy(dst, src, bytes);
    return;
  }
  bool const streaming = bytes >= (size_t{8} << 20) && TuneValue("nt_copy", 1) != 0;
  int64_t const blocks = static_cast<int64_t>((bytes + kBlock - 1) / kBlock);
  SharedPool::Run(blocks, [=](int64_t b) {
    size_t const begin = static_cast<size_t>(b) * kBlock;
    size_t const len = std::min(kBlock, bytes - begin);
    if (streaming) StreamingCopy(static_cast<char*>(dst) + begin, static_cast<const char*>(src) + begin, len);
    else std::memcpy(static_cast<char*>(dst) + begin, static_cast<const char*>(src) + begin, len);
  });
}

// Chunk sizes ramp up at the start and down at the end of a transfer.  The first device->host copy cannot start before
// the first host->device copy and kernel have finished, and while the last results drain nothing flows host->device, so
// with full-size chunks the link runs in one direction only for two chunk times per call; short first and last chunks
// shrink that to two short-chunk times (profiles/r01_e2e_chunk_sweep.log).  A transfer that fits one slot — a single
// KITTI scan, the 10 Hz sensor case — is cut into `frame_parts` equal pieces for the same reason: with one piece H2D,
// kernel and D2H run strictly one after the other (99 us per 130 000-point scan); with four the copies of neighbouring
// pieces overlap.  Below ~512 KB per copy the per-copy latency dominates.
int64_t NextChunkPoints(int64_t chunk_index, int64_t remaining, int64_t capacity, int64_t total) {
  int64_t const min_chunk = std::max(1024, TuneValue("min_chunk", 1 << 15));  // 32 768 points = 512 KiB
  if (total <= capacity) {
    int64_t const parts = std::max(1, TuneValue("frame_parts", 4));
    int64_t const c = std::max(min_chunk, ((total + parts - 1) / parts + 1023) & ~int64_t{1023});
    return std::min(c, remaining);
  }
  int64_t c = capacity;
  if (chunk_index < 3) c = std::max(min_chunk, capacity >> (3 - chunk_index));          // capacity / 8, / 4, / 2, then full
  if (remaining < 2 * capacity) c = std::min(c, std::max(min_chunk, remaining / 2));  // halve what is left, down to the floor
  if (c >= capacity) return std::min(capacity, remaining);
  if (c >= remaining) return remaining;
  return std::max<int64_t>(c & ~int64_t{1023}, 1024);  // ramp sizes only: whole multiples of 1024 points
}

// Chunks of a zero-copy transfer: one launch per slot-sized piece (`zc_parts` = 1).  Cutting a scan into 2-3 pieces so that the
// staging copies of pageable memory overlap the neighbouring piece's kernel did not pay (131 vs 136-147 us per 130 000-point scan,
// profiles/r02_sweep_single_scan.log): every extra launch + event costs what the overlap wins.
int64_t ZeroCopyChunkPoints(int64_t remaining, int64_t capacity, int64_t total) {
  int64_t const parts = std::max(1, TuneValue("zc_parts", 1));
  int64_t const c = std::max<int64_t>(8192, ((total + parts - 1) / parts + 1023) & ~int64_t{1023});
  return std::min({c, remaining, capacity});
}

// Device-visible alias of a pinned host pointer (identical under UVA for cudaHostAlloc memory; may differ for registered memory).
template <class T>
bool DeviceAlias(T* host, T** dev) {
  void* d = nullptr;
  if (cudaHostGetDevicePointer(&d, const_cast<void*>(static_cast<const void*>(host)), 0) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  *dev = static_cast<T*>(d);
  return true;
}

template <class Fetch, class Launch, class Sink, class Deliver>
int StreamChunksImpl(kmc_b200_handle* h, int64_t n, bool zero_copy, Fetch&& fetch, Launch&& launch, Sink&& sink, Deliver&& deliver) {
  constexpr int kSlots = kmc_b200_handle::kSlots;
  struct Pending {
    int64_t first = 0, count = 0;
    bool active = false;
  } pending[kSlots];

  auto retire = [&](int slot) -> int {
    if (!pending[slot].active) return KMC_B200_OK;
    KMC_CUDA_TRY(cudaEventSynchronize(h->done[slot]));
    pending[slot].active = false;
    return deliver(slot, pending[slot].first, pending[slot].count);
  };

  int64_t chunk_index = 0;
  int64_t count = 0;
  for (int64_t first = 0; first < n; first += count, ++chunk_index) {
    int const slot = static_cast<int>(chunk_index % kSlots);
    count = zero_copy ? ZeroCopyChunkPoints(n - first, h->capacity, n) : NextChunkPoints(chunk_index, n - first, h->capacity, n);
    size_t const bytes = static_cast<size_t>(count) * 16;
    if (int rc = retire(slot)) return rc;
    const float* src = nullptr;
    if (int rc = fetch(slot, first, count, &src)) return rc;
    float* const dst = sink(slot, first);
    if (zero_copy) {
      const float* d_src = nullptr;
      float* d_dst = nullptr;
      if (!DeviceAlias(src, &d_src) || !DeviceAlias(dst, &d_dst)) return Fail(KMC_B200_ERR_CUDA, "pinned host buffer is not mapped into the device address space");
      if (int rc = launch(slot, first, count, d_src, d_dst)) return rc;
    } else {
      KMC_CUDA_TRY(cudaMemcpyAsync(h->d_in[slot], src, bytes, cudaMemcpyHostToDevice, h->stream[slot]));
      if (int rc = launch(slot, first, count, h->d_in[slot], h->d_out[slot])) return rc;
      KMC_CUDA_TRY(cudaMemcpyAsync(dst, h->d_out[slot], bytes, cudaMemcpyDeviceToHost, h->stream[slot]));
    }
    KMC_CUDA_TRY(cudaEventRecord(h->done[slot], h->stream[slot]));
    pending[slot] = {first, count, true};
  }
  for (int k = 0; k < kSlots; ++k) {
    int const slot = static_cast<int>((chunk_index + k) % kSlots);  // oldest first
    if (int rc = retire(slot)) return rc;
  }
  return KMC_B200_OK;
}

// A host ARRAY through the device: pinned caller memory is used directly, pageable memory is staged through the slots;
// transfers of up to zc_points (default 2 000 000 points = 32 MB each way) skip the copy engines (zero_copy above).
template <class Launch>
int StreamThroughDevice(kmc_b200_handle* h, const float* in, float* out, int64_t n, Launch&& launch) {
  bool const in_pinned = IsPinnedHost(in);
  bool const out_pinned = IsPinnedHost(out);
  bool zero_copy = n <= TuneValue("zc_points", 2000000);
  if (zero_copy && in_pinned) {
    const float* alias = nullptr;
    zero_copy = DeviceAlias(in, &alias);
  }
  if (zero_copy && out_pinned) {
    float* alias = nullptr;
    zero_copy = DeviceAlias(out, &alias);
  }
  return StreamChunks(
      h, n, zero_copy,
      [&](int slot, int64_t first, int64_t count, const float** src) -> int {
        *src = in + 4 * first;
        if (!in_pinned) {
          StagingCopy(h, h->h_in[slot], in + 4 * first, static_cast<size_t>(count) * 16);
          *src = h->h_in[slot];
        }
        return KMC_B200_OK;
      },
      launch, [&](int slot, int64_t first) -> float* { return out_pinned ? out + 4 * first : h->h_out[slot]; },
      [&](int slot, int64_t first, int64_t count) -> int {
        if (!out_pinned) StagingCopy(h, out + 4 * first, h->h_out[slot], static_cast<size_t>(count) * 16);
        return KMC_B200_OK;
      });
}

// Launch shape of a chunk of a host call: the HBM-tuned shape when the chunk sits in device memory; when the kernel reads
// pinned host memory over PCIe (zero copy) what matters is the number of read requests in flight on the link, not HBM
// queues: one CTA of 256 threads per SM, 128-bit accesses (KMC_B200_TUNE zc_block / zc_ctas / zc_vec / zc_tiles).
kmc_b200::dev::LaunchConfig PickHostCallConfig(const kmc_b200_handle* h, int64_t count, const float* src, const float* dst, bool batch) {
  bool const aligned32 = Aligned(src, 32) && Aligned(dst, 32);
  auto cfg = kmc_b200::dev::PickConfig(count, aligned32, false, h->sm_count, batch);
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
    cfg.bulk = 0;
    cfg.hint = TuneValue("zc_hint", 0) == 1 ? 1 : 0;
    cfg.unroll = TuneValue("zc_unroll", 1) == 2 ? 2 : 1;
    cfg.block = TuneValue("zc_block", 256);
    cfg.ctas_per_sm = std::max(1, TuneValue("zc_ctas", 1));
    cfg.vec = (TuneValue("zc_vec", 1) == 2 && aligned32) ? 2 : 1;
    cfg.item_tiles = std::max(1, TuneValue("zc_tiles", 4));  // 70.9-71.3 us vs 72.5-73.3 with 1 tile; every other shape 71-86 (profiles/r02_sweep_single_scan_fine.log)
    if (cfg.block != 128 && cfg.block != 512) cfg.block = 256;
    if (TuneValue("zc_bulk", 0) == 1) {  // experiment: TMA bulk copies straight from / to the pinned host buffers
      cfg.bulk = 1;
      cfg.block = cfg.block == 128 ? 128 : 256;
      cfg.unroll = TuneValue("zc_unroll", 2);
      cfg.stages = TuneValue("zc_stages", 3);
    }
  } else {
    cudaGetLastError();
  }
  return cfg;
}

int CheckOffsets(const int64_t* offsets, int32_t n_frames) {
  if (offsets[0] != 0) return Fail(KMC_B200_ERR_BAD_SIZE, "frame_offsets[0] must be 0");
  for (int32_t f = 0; f < n_frames; ++f)
    if (offsets[f + 1] < offsets[f]) return Fail(KMC_B200_ERR_BAD_SIZE, "frame_offsets must be non-decreasing");
  return KMC_B200_OK;
}

}  // namespace

// =============================================================================================================
// C ABI
// =============================================================================================================
extern "C" {

// ---- handle ---------------------------------------------------------------------------------------------------------
int kmc_b200_handle_create(int device, int64_t capacity_points, kmc_b200_handle** out) try {
  if (!out) return Fail(KMC_B200_ERR_NULL_POINTER, "handle_create: null output");
  *out = nullptr;
  if (capacity_points <= 0) return Fail(KMC_B200_ERR_BAD_SIZE, "handle_create: capacity must be positive");
  int n_dev = kmc_b200_device_count();
  if (n_dev < 0) return n_dev;
  if (device < 0 || device >= n_dev) return Fail(KMC_B200_ERR_NO_DEVICE, "handle_create: no such device");
  auto* h = new kmc_b200_handle;
  h->device = device;
  h->capacity = (capacity_points + 7) & ~int64_t{7};  // even point count per chunk keeps 256-bit accesses aligned
  DeviceGuard const guard(device);
  cudaError_t e = guard.status();
  int rc = (e == cudaSuccess) ? SmCount(device, &h->sm_count) : FailCuda(e, "cudaSetDevice");
  size_t const bytes = static_cast<size_t>(h->capacity) * 16;
  for (int s = 0; s < kmc_b200_handle::kSlots && rc == KMC_B200_OK; ++s) {
    if ((e = cudaStreamCreateWithFlags(&h->stream[s], cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&h->done[s], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaMalloc(&h->d_in[s], bytes)) != cudaSuccess || (e = cudaMalloc(&h->d_out[s], bytes)) != cudaSuccess ||
        (e = cudaMallocHost(&h->h_in[s], bytes)) != cudaSuccess || (e = cudaMallocHost(&h->h_out[s], bytes)) != cudaSuccess)
      rc = FailCuda(e, "handle_create: allocating streams/buffers");
  }
  if (rc != KMC_B200_OK) {
    std::string const keep = LastError();
    FreeHandle(h);
    LastError() = keep;
    return rc;
  }
  *out = h;
  return KMC_B200_OK;
}
KMC_CATCH_AT_BOUNDARY("handle_create")

int kmc_b200_handle_destroy(kmc_b200_handle* h) {
  FreeHandle(h);
  return KMC_B200_OK;
}

int kmc_b200_default_handle(int device, kmc_b200_handle** out) try {
  if (!out) return Fail(KMC_B200_ERR_NULL_POINTER, "default_handle: null output");
  *out = nullptr;
  if (device < 0 || device >= kMaxDevices) return Fail(KMC_B200_ERR_NO_DEVICE, "default_handle: device ordinal out of range");
  static std::mutex mu;
  static kmc_b200_handle* table[kMaxDevices] = {};
  std::lock_guard<std::mutex> lock(mu);
  if (!table[device]) {
    if (int rc = kmc_b200_handle_create(device, 250000, &table[device])) return rc;  // data_io.hpp:17 in the reference
  }
  *out = table[device];
  return KMC_B200_OK;
}
KMC_CATCH_AT_BOUNDARY("default_handle")

int kmc_b200_handle_set_file_callback(kmc_b200_handle* h, kmc_b200_file_done_fn fn, void* user) {
  if (!h) return Fail(KMC_B200_ERR_NULL_POINTER, "handle_set_file_callback: null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  h->file_done = fn;
  h->file_done_user = user;
  return KMC_B200_OK;
}

int kmc_b200_handle_device(const kmc_b200_handle* h) { return h ? h->device : KMC_B200_ERR_NULL_POINTER; }
int64_t kmc_b200_handle_capacity(const kmc_b200_handle* h) { return h ? h->capacity : static_cast<int64_t>(KMC_B200_ERR_NULL_POINTER); }

// ---- host entry points -------------------------------------------------------------------------------------------------
int kmc_b200_deskew_frame_host(kmc_b200_handle* h, const float* in, float* out, int64_t n, const kmc_b200_frame_params* params,
                               int mode) try {
  TraceRange const trace("kmc_b200_deskew_frame_host");
  if (!h) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_frame_host: null handle");
  if (n < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_frame_host: negative n_points");
  if (!ValidMode(mode)) return Fail(KMC_B200_ERR_BAD_MODE, "deskew_frame_host: unknown time mode");
  if (!params) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_frame_host: null params");
  if (n == 0) return KMC_B200_OK;
  if (!in || !out) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_frame_host: null point buffer");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard const guard(h->device);
  KMC_CUDA_TRY(guard.status());
  kmc_b200_frame_params const P = *params;
  return StreamThroughDevice(h, in, out, n, [&](int slot, int64_t, int64_t count, const float* src, float* dst) -> int {
    auto const cfg = PickHostCallConfig(h, count, src, dst, false);
    KMC_CUDA_TRY(kmc_b200::dev::LaunchDeskewFrame(src, dst, count, P, mode, cfg, h->sm_count, h->stream[slot]));
    return KMC_B200_OK;
  });
}
KMC_CATCH_AT_BOUNDARY("deskew_frame_host")

int kmc_b200_deskew_batch_host(kmc_b200_handle* h, const float* in, float* out, const int64_t* offsets,
                               const kmc_b200_frame_params* params, int32_t n_frames, int mode) try {
  TraceRange const trace("kmc_b200_deskew_batch_host");
  if (!h) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_batch_host: null handle");
  if (n_frames < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_batch_host: negative n_frames");
  if (!ValidMode(mode)) return Fail(KMC_B200_ERR_BAD_MODE, "deskew_batch_host: unknown time mode");
  if (n_frames == 0) return KMC_B200_OK;
  if (!offsets || !params) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_batch_host: null table");
  if (int rc = CheckOffsets(offsets, n_frames)) return rc;
  int64_t const n_total = offsets[n_frames];
  if (n_total == 0) return KMC_B200_OK;
  if (!in || !out) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_batch_host: null point buffer");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard const guard(h->device);
  KMC_CUDA_TRY(guard.status());
  if (int rc = EnsureTables(h, n_frames)) return rc;
  // tables go up once, on slot 0's stream; the other slots wait for them through an event
  KMC_CUDA_TRY(cudaMemcpyAsync(h->d_offsets, offsets, static_cast<size_t>(n_frames + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream[0]));
  KMC_CUDA_TRY(cudaMemcpyAsync(h->d_params, params, static_cast<size_t>(n_frames) * sizeof(kmc_b200_frame_params), cudaMemcpyHostToDevice, h->stream[0]));
  KMC_CUDA_TRY(cudaStreamSynchronize(h->stream[0]));
  return StreamThroughDevice(h, in, out, n_total, [&](int slot, int64_t first, int64_t count, const float* src, float* dst) -> int {
    auto const cfg = PickHostCallConfig(h, count, src, dst, true);
    KMC_CUDA_TRY(kmc_b200::dev::LaunchDeskewBatch(src, dst, h->d_offsets, h->d_params, n_frames, count, first, n_total, mode, cfg,
                                                  h->sm_count, h->stream[slot]));
    return KMC_B200_OK;
  });
}
KMC_CATCH_AT_BOUNDARY("deskew_batch_host")

int kmc_b200_deskew_batch_multi_gpu(kmc_b200_handle* const* handles, int32_t n_handles, const float* in, float* out,
                                    const int64_t* offsets, const kmc_b200_frame_params* params, int32_t n_frames, int mode) try {
  TraceRange const trace("kmc_b200_deskew_batch_multi_gpu");
  if (!handles || n_handles <= 0) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_batch_multi_gpu: no handles");
  if (n_frames < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_batch_multi_gpu: negative n_frames");
  if (!ValidMode(mode)) return Fail(KMC_B200_ERR_BAD_MODE, "deskew_batch_multi_gpu: unknown time mode");
  if (n_frames == 0) return KMC_B200_OK;
  if (!offsets || !params) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_batch_multi_gpu: null table");
  for (int32_t i = 0; i < n_handles; ++i) {
    if (!handles[i]) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_batch_multi_gpu: null handle");
    for (int32_t j = 0; j < i; ++j)
      if (handles[j]->device == handles[i]->device)
        return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_batch_multi_gpu: handles must live on distinct devices");
  }
  if (int rc = CheckOffsets(offsets, n_frames)) return rc;
  std::vector<int> status(static_cast<size_t>(n_handles), KMC_B200_OK);
  std::vector<std::string> message(static_cast<size_t>(n_handles));
  std::vector<std::thread> workers;
  struct JoinAll {
    std::vector<std::thread>& pool;
    ~JoinAll() {
      for (auto& w : pool)
        if (w.joinable()) w.join();
    }
  } join_all{workers};
  workers.reserve(static_cast<size_t>(n_handles));
  for (int32_t i = 0; i < n_handles; ++i) {
    workers.emplace_back([&, i] {
      try {
        int64_t fb = 0, fe = 0;
        kmc_b200_shard_range(n_frames, n_handles, i, &fb, &fe);
        if (fe <= fb) return;
        std::vector<int64_t> local(static_cast<size_t>(fe - fb + 1));
        for (int64_t f = fb; f <= fe; ++f) local[static_cast<size_t>(f - fb)] = offsets[f] - offsets[fb];
        status[i] = kmc_b200_deskew_batch_host(handles[i], in ? in + 4 * offsets[fb] : nullptr, out ? out + 4 * offsets[fb] : nullptr,
                                               local.data(), params + fb, static_cast<int32_t>(fe - fb), mode);
      } catch (...) {
        status[i] = kmc_b200::internal::FailException("deskew_batch_multi_gpu (device thread)");
      }
      if (status[i] != KMC_B200_OK) message[i] = LastError();
    });
  }
  for (auto& w : workers) w.join();
  for (int32_t i = 0; i < n_handles; ++i)
    if (status[i] != KMC_B200_OK) return Fail(status[i], "device " + std::to_string(handles[i]->device) + ": " + message[i]);
  return KMC_B200_OK;
}
KMC_CATCH_AT_BOUNDARY("deskew_batch_multi_gpu")

namespace {

// Staging of the narrow transport: three slots of [5 float columns up | 3 float columns down], pinned and on the device.
int EnsureF64Staging(kmc_b200_handle* h, int64_t chunk_points) {
  if (chunk_points <= h->f64_chunk) return KMC_B200_OK;
  if (h->f64_pinned) cudaFreeHost(h->f64_pinned);
  if (h->f64_device) cudaFree(h->f64_device);
  h->f64_pinned = nullptr;
  h->f64_device = nullptr;
  h->f64_chunk = 0;
  size_t const bytes = static_cast<size_t>(kmc_b200_handle::kSlots) * 8 * static_cast<size_t>(chunk_points) * sizeof(float);
  KMC_CUDA_TRY(cudaMallocHost(&h->f64_pinned, bytes));
  KMC_CUDA_TRY(cudaMalloc(&h->f64_device, bytes));
  h->f64_chunk = chunk_points;
  return KMC_B200_OK;
}

// Points per chunk of a reference-layout host call: a KITTI-size cloud is cut in two so that conversion, both copy
// directions and the final add overlap (1 / 2 / 3 / 4 / 6 / 8 pieces: 180 / 160 / 189 / 193 / 288 / 304 us per 123 397-point
// frame, profiles/r02_sweep_dropin.log — every piece costs four API calls and an event wait); large clouds move in
// 64 Ki-point chunks (1 MB up, 0.75 MB down).
int64_t F64ChunkPoints(int64_t n) {
  int64_t const cap = std::max(4096, TuneValue("f64_chunk", 65536));
  int64_t const parts = std::max(1, TuneValue("f64_parts", 2));
  int64_t c = (n + parts - 1) / parts;
  c = std::min(std::max<int64_t>(c, 8192), cap);
  return (c + 3) & ~int64_t{3};
}

}  // namespace

// MotionCompensateFrame on the reference's layout from HOST memory (motion_compensation.cpp:16-28), narrow transport:
//   pass 1 (host threads)  x, y, z -> float columns, s_i = (t_i - t_start)/(t_end - t_start) - x_req in double -> float column,
//                          w_i copied through to the result and checked (a float w column is shipped only if some w != 1),
//                          stamps checked against [t_start, t_end]                       -> pinned staging
//   H2D 16 B/point -> DeskewDeltaColumnsKernel -> D2H 12 B/point                          (slot's stream)
//   pass 2 (host threads)  out = x + double(dx), ...                                      <- pinned staging
// in chunks over three slots, so the host passes of one chunk run while the link and the GPU work on the others.  The
// chunk sequence runs across ALL frames of a call without draining in between.
extern "C++" {
namespace {

struct F64Frame {
  const double* cloud;
  const double* stamps;
  double* out;
  int64_t n;
  double t_start, t_end, t_req;
  kmc_b200_frame_params P;
  int flags;
};

int F64HostPipeline(kmc_b200_handle* h, F64Frame* frames, int32_t n_frames) {
  constexpr int kSlots = kmc_b200_handle::kSlots;
  constexpr int64_t kBlock = 4096;  // points per host task
  int64_t longest = 0;
  for (int32_t f = 0; f < n_frames; ++f) longest = std::max(longest, frames[f].n);
  if (longest == 0) return KMC_B200_OK;
  if (int rc = EnsureF64Staging(h, F64ChunkPoints(longest))) return rc;
  size_t const slot_floats = 8 * static_cast<size_t>(h->f64_chunk);
  struct Pending {
    int32_t frame = 0;
    int64_t first = 0, count = 0;
    bool active = false;
  } pending[kSlots];

  auto retire = [&](int slot) -> int {
    if (!pending[slot].active) return KMC_B200_OK;
    KMC_CUDA_TRY(cudaEventSynchronize(h->done[slot]));
    F64Frame const& fr = frames[pending[slot].frame];
    int64_t const n = fr.n, first = pending[slot].first, count = pending[slot].count, stride = (count + 3) & ~int64_t{3};
    const double* const X = fr.cloud;
    const double* const Y = fr.cloud + n;
    const double* const Z = fr.cloud + 2 * n;
    double* const out = fr.out;
    const float* const d = h->f64_pinned + slot * slot_floats + 5 * static_cast<size_t>(h->f64_chunk);
    SharedPool::Run((count + kBlock - 1) / kBlock, [=](int64_t b) {
      int64_t const i1 = std::min(count, (b + 1) * kBlock);
      for (int64_t i = b * kBlock; i < i1; ++i) out[first + i] = X[first + i] + static_cast<double>(d[i]);
      for (int64_t i = b * kBlock; i < i1; ++i) out[n + first + i] = Y[first + i] + static_cast<double>(d[stride + i]);
      for (int64_t i = b * kBlock; i < i1; ++i) out[2 * n + first + i] = Z[first + i] + static_cast<double>(d[2 * stride + i]);
    });
    pending[slot].active = false;
    return KMC_B200_OK;
  };

  int64_t total = 0;
  for (int32_t f = 0; f < n_frames; ++f) total += frames[f].n;
  bool const zero_copy = total <= TuneValue("f64_zc_points", 1000000);
  int64_t k = 0;  // chunk counter across all frames
  for (int32_t f = 0; f < n_frames; ++f) {
    F64Frame& fr = frames[f];
    int64_t const n = fr.n;
    if (n == 0) continue;
    int64_t const chunk = std::min(F64ChunkPoints(n), h->f64_chunk);
    double const t_start = fr.t_start, t_end = fr.t_end, duration = fr.t_end - fr.t_start;
    double const x_req = (fr.t_req - fr.t_start) / duration;
    const double* const X = fr.cloud;
    const double* const Y = fr.cloud + n;
    const double* const Z = fr.cloud + 2 * n;
    const double* const W = fr.cloud + 3 * n;
    const double* const stamps = fr.stamps;
    double* const out = fr.out;
    for (int64_t first = 0; first < n; first += chunk, ++k) {
      int const slot = static_cast<int>(k % kSlots);
      if (int rc = retire(slot)) return rc;
      int64_t const count = std::min(chunk, n - first), stride = (count + 3) & ~int64_t{3};
      float* const up = h->f64_pinned + slot * slot_floats;
      float* const d_up = h->f64_device + slot * slot_floats;
      float* const d_down = d_up + 5 * static_cast<size_t>(h->f64_chunk);
      std::atomic<int> chunk_flags{0};
      SharedPool::Run((count + kBlock - 1) / kBlock, [&, first, count, stride, up](int64_t b) {
        int64_t const i0 = b * kBlock, i1 = std::min(count, (b + 1) * kBlock);
        for (int64_t i = i0; i < i1; ++i) up[i] = static_cast<float>(X[first + i]);
        for (int64_t i = i0; i < i1; ++i) up[stride + i] = static_cast<float>(Y[first + i]);
        for (int64_t i = i0; i < i1; ++i) up[2 * stride + i] = static_cast<float>(Z[first + i]);
        int outside = 0, not_one = 0;
        for (int64_t i = i0; i < i1; ++i) {
          double const t = stamps[first + i];
          outside |= !(t >= t_start && t <= t_end);
          up[3 * stride + i] = static_cast<float>((t - t_start) / duration - x_req);  // FractionOfTrajectory, trajectory_interpolation.cpp:49-51
        }
        for (int64_t i = i0; i < i1; ++i) {
          double const w = W[first + i];
          not_one |= (w != 1.0);
          out[3 * n + first + i] = w;
        }
        if (i1 == count)
          for (int c = 0; c < 4; ++c)
            for (int64_t i = count; i < stride; ++i) up[c * stride + i] = 0.0f;
        if (outside | not_one) chunk_flags.fetch_or((outside ? 1 : 0) | (not_one ? 2 : 0), std::memory_order_relaxed);
      });
      int const cf = chunk_flags.load(std::memory_order_relaxed);
      fr.flags |= cf;
      bool const has_w = (cf & 2) != 0;
      if (has_w) {  // rare: a non-homogeneous 4th column travels as a fifth float column
        SharedPool::Run((stride + kBlock - 1) / kBlock, [=](int64_t b) {
          int64_t const i1 = std::min(stride, (b + 1) * kBlock);
          for (int64_t i = b * kBlock; i < i1; ++i) up[4 * stride + i] = i < count ? static_cast<float>(W[first + i]) : 1.0f;
        });
      }
      cudaStream_t const st = h->stream[slot];
      float* const down = up + 5 * static_cast<size_t>(h->f64_chunk);
      if (zero_copy) {
        // KITTI-size frames: the kernel reads the pinned columns and writes the pinned displacement columns itself — two
        // copy-engine operations of ~1 MB cost more in fixed latency than SM-issued PCIe accesses lose (see StreamChunksImpl)
        KMC_CUDA_TRY(kmc_b200::dev::LaunchDeskewDeltaColumns(up, down, stride, has_w, fr.P, h->sm_count, st, /*over_pcie=*/true));
      } else {
        KMC_CUDA_TRY(cudaMemcpyAsync(d_up, up, static_cast<size_t>(has_w ? 5 : 4) * stride * sizeof(float), cudaMemcpyHostToDevice, st));
        KMC_CUDA_TRY(kmc_b200::dev::LaunchDeskewDeltaColumns(d_up, d_down, stride, has_w, fr.P, h->sm_count, st, false));
        KMC_CUDA_TRY(cudaMemcpyAsync(down, d_down, static_cast<size_t>(3) * stride * sizeof(float), cudaMemcpyDeviceToHost, st));
      }
      KMC_CUDA_TRY(cudaEventRecord(h->done[slot], st));
      pending[slot] = {f, first, count, true};
    }
  }
  for (int j = 0; j < kSlots; ++j)
    if (int rc = retire(static_cast<int>((k + j) % kSlots))) return rc;  // oldest first
  return KMC_B200_OK;
}

// Validates one frame's arguments (status as the single-frame call) — everything is checked before anything is launched.
int F64CheckFrame(const char* who, const F64Frame& fr) {
  if (fr.n < 0) return Fail(KMC_B200_ERR_BAD_SIZE, std::string(who) + ": negative n_points");
  if (!(fr.t_end > fr.t_start)) return Fail(KMC_B200_ERR_EMPTY_INTERVAL, std::string(who) + ": t_end <= t_start");
  if (!(fr.t_req >= fr.t_start && fr.t_req <= fr.t_end))
    return Fail(KMC_B200_ERR_TIME_OUT_OF_RANGE, std::string(who) + ": requested time outside [t_start, t_end]");
  if (fr.n > 0 && (!fr.cloud || !fr.stamps || !fr.out)) return Fail(KMC_B200_ERR_NULL_POINTER, std::string(who) + ": null buffer");
  return KMC_B200_OK;
}

int F64RunLocked(kmc_b200_handle* h, F64Frame* frames, int32_t n_frames) {
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard const guard(h->device);
  KMC_CUDA_TRY(guard.status());
  int const rc = F64HostPipeline(h, frames, n_frames);
  if (rc != KMC_B200_OK) {  // no copy may still touch the staging buffers once the call has reported an error
    std::string const keep = LastError();
    for (int s = 0; s < kmc_b200_handle::kSlots; ++s) cudaStreamSynchronize(h->stream[s]);
    cudaGetLastError();
    LastError() = keep;
  }
  return rc;
}

}  // namespace
}  // extern "C++"

int kmc_b200_deskew_cloud_f64_host(kmc_b200_handle* h, const double* cloud, const double* stamps, double* out, int64_t n, double t_start,
                                   double t_end, double t_req, const kmc_b200_frame_params* params, int* flags_out) try {
  TraceRange const trace("kmc_b200_deskew_cloud_f64_host");
  if (flags_out) *flags_out = 0;
  if (!h) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_cloud_f64_host: null handle");
  if (!params) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_cloud_f64_host: null params");
  F64Frame frame{cloud, stamps, out, n, t_start, t_end, t_req, *params, 0};
  if (int rc = F64CheckFrame("deskew_cloud_f64_host", frame)) return rc;
  if (n == 0) return KMC_B200_OK;
  if (int rc = F64RunLocked(h, &frame, 1)) return rc;
  if (flags_out) *flags_out = frame.flags;
  if (frame.flags & 1) return Fail(KMC_B200_ERR_TIME_OUT_OF_RANGE, "a point stamp lies outside [t_start, t_end] (reference asserts, trajectory_interpolation.cpp:32)");
  return KMC_B200_OK;
}
KMC_CATCH_AT_BOUNDARY("deskew_cloud_f64_host")

int kmc_b200_deskew_cloud_f64_batch_host(kmc_b200_handle* h, const double* const* clouds, const double* const* stamps, double* const* outs,
                                         const int64_t* n_points, const double* times, const kmc_b200_frame_params* params,
                                         int32_t n_frames, int* flags_out) try {
  TraceRange const trace("kmc_b200_deskew_cloud_f64_batch_host");
  if (!h) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_cloud_f64_batch_host: null handle");
  if (n_frames < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_cloud_f64_batch_host: negative n_frames");
  if (n_frames == 0) return KMC_B200_OK;
  if (!clouds || !stamps || !outs || !n_points || !times || !params) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_cloud_f64_batch_host: null table");
  std::vector<F64Frame> frames(static_cast<size_t>(n_frames));
  for (int32_t f = 0; f < n_frames; ++f) {
    frames[static_cast<size_t>(f)] = F64Frame{clouds[f], stamps[f], outs[f], n_points[f], times[3 * f], times[3 * f + 1], times[3 * f + 2], params[f], 0};
    if (flags_out) flags_out[f] = 0;
    if (int rc = F64CheckFrame("deskew_cloud_f64_batch_host", frames[static_cast<size_t>(f)]))
      return Fail(rc, "frame " + std::to_string(f) + ": " + LastError());
  }
  if (int rc = F64RunLocked(h, frames.data(), n_frames)) return rc;
  int status = KMC_B200_OK;
  for (int32_t f = 0; f < n_frames; ++f) {
    if (flags_out) flags_out[f] = frames[static_cast<size_t>(f)].flags;
    if ((frames[static_cast<size_t>(f)].flags & 1) && status == KMC_B200_OK)
      status = Fail(KMC_B200_ERR_TIME_OUT_OF_RANGE, "frame " + std::to_string(f) + ": a point stamp lies outside [t_start, t_end] (reference asserts, trajectory_interpolation.cpp:32)");
  }
  return status;
}
KMC_CATCH_AT_BOUNDARY("deskew_cloud_f64_batch_host")

// GetPseudoTimeStamps (timestamp_mocking.cpp:56-63) for HOST columns.  Round 1 issued three cudaMemcpyAsync on pageable Eigen
// memory (the driver stages those through its own bounce buffers at ~10 GB/s).  Now the columns are staged into the slot's
// pinned buffer by the host pool, and for KITTI-size clouds the kernel reads them — and writes the stamps — in pinned memory
// directly (zero copy, see StreamChunksImpl); large clouds go through the copy engines in slot-sized chunks over three slots.
int kmc_b200_pseudo_time_stamps_xy_host(kmc_b200_handle* h, const double* x, const double* y, int64_t n, double start, double end,
                                        double* stamps) try {
  TraceRange const trace("kmc_b200_pseudo_time_stamps_xy_host");
  if (!h) return Fail(KMC_B200_ERR_NULL_POINTER, "pseudo_time_stamps_xy_host: null handle");
  if (n < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "pseudo_time_stamps_xy_host: negative n_points");
  if (n == 0) return KMC_B200_OK;
  if (!x || !y || !stamps) return Fail(KMC_B200_ERR_NULL_POINTER, "pseudo_time_stamps_xy_host: null argument");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard const guard(h->device);
  KMC_CUDA_TRY(guard.status());
  constexpr int kSlots = kmc_b200_handle::kSlots;
  bool const zero_copy = n <= TuneValue("zc_points", 2000000);
  // a slot holds capacity x 16 bytes: x | y of `capacity` points in h_in / d_in, their stamps in the first half of h_out / d_out
  // one piece for a KITTI-size cloud: 133 us per 123 397 points against 204 us in two or four pieces and 171 us through the copy
  // engines (profiles/r02_get_pseudo_time_stamps_host_call.log; kmc::GetPseudoTimeStamps incl. the result allocation)
  int64_t const parts = zero_copy ? std::max(1, TuneValue("stamp_parts", 1)) : 1;
  int64_t const chunk = std::min(h->capacity, std::max<int64_t>(8192, ((n + parts - 1) / parts + 1023) & ~int64_t{1023}));
  struct Pending {
    int64_t first = 0, count = 0;
    bool active = false;
  } pending[kSlots];
  auto retire = [&](int slot) -> int {
    if (!pending[slot].active) return KMC_B200_OK;
    KMC_CUDA_TRY(cudaEventSynchronize(h->done[slot]));
    StagingCopy(h, stamps + pending[slot].first, h->h_out[slot], static_cast<size_t>(pending[slot].count) * sizeof(double));
    pending[slot].active = false;
    return KMC_B200_OK;
  };
  auto body = [&]() -> int {
    int64_t k = 0;
    for (int64_t first = 0; first < n; first += chunk, ++k) {
      int const slot = static_cast<int>(k % kSlots);
      if (int rc = retire(slot)) return rc;
      int64_t const count = std::min(chunk, n - first);
      size_t const col = static_cast<size_t>(count) * sizeof(double);
      auto* const hx = reinterpret_cast<double*>(h->h_in[slot]);
      auto* const hs = reinterpret_cast<double*>(h->h_out[slot]);
      StagingCopy(h, hx, x + first, col);
      StagingCopy(h, hx + count, y + first, col);
      cudaStream_t const st = h->stream[slot];
      if (zero_copy) {
        KMC_CUDA_TRY(kmc_b200::dev::LaunchPseudoTimeStampsXy(hx, hx + count, hs, count, start, end, h->sm_count, st));
      } else {
        auto* const dx = reinterpret_cast<double*>(h->d_in[slot]);
        auto* const ds = reinterpret_cast<double*>(h->d_out[slot]);
        KMC_CUDA_TRY(cudaMemcpyAsync(dx, hx, 2 * col, cudaMemcpyHostToDevice, st));
        KMC_CUDA_TRY(kmc_b200::dev::LaunchPseudoTimeStampsXy(dx, dx + count, ds, count, start, end, h->sm_count, st));
        KMC_CUDA_TRY(cudaMemcpyAsync(hs, ds, col, cudaMemcpyDeviceToHost, st));
      }
      KMC_CUDA_TRY(cudaEventRecord(h->done[slot], st));
      pending[slot] = {first, count, true};
    }
    for (int j = 0; j < kSlots; ++j)
      if (int rc = retire(static_cast<int>((k + j) % kSlots))) return rc;  // oldest first
    return KMC_B200_OK;
  };
  int const rc = body();
  if (rc != KMC_B200_OK) {
    std::string const keep = LastError();
    for (int s2 = 0; s2 < kSlots; ++s2) cudaStreamSynchronize(h->stream[s2]);
    cudaGetLastError();
    LastError() = keep;
  }
  return rc;
}
KMC_CATCH_AT_BOUNDARY("pseudo_time_stamps_xy_host")

int kmc_b200_project_frame_host(kmc_b200_handle* h, const float* in, float* uvzc_out, int64_t n, const kmc_b200_camera_params* camera) try {
  TraceRange const trace("kmc_b200_project_frame_host");
  if (!h) return Fail(KMC_B200_ERR_NULL_POINTER, "project_frame_host: null handle");
  if (n < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "project_frame_host: negative n_points");
  if (!camera) return Fail(KMC_B200_ERR_NULL_POINTER, "project_frame_host: null camera params");
  if (n == 0) return KMC_B200_OK;
  if (!in || !uvzc_out) return Fail(KMC_B200_ERR_NULL_POINTER, "project_frame_host: null point buffer");
  if (in == uvzc_out) return Fail(KMC_B200_ERR_BAD_SIZE, "project_frame_host: the pixel buffer must not alias the cloud");
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard const guard(h->device);
  KMC_CUDA_TRY(guard.status());
  kmc_b200_camera_params const K = *camera;
  return StreamThroughDevice(h, in, uvzc_out, n, [&](int slot, int64_t, int64_t count, const float* src, float* dst) -> int {
    bool const vec2 = Aligned(src, 32) && Aligned(dst, 32);
    KMC_CUDA_TRY(kmc_b200::dev::LaunchProject(src, nullptr, dst, count, nullptr, K, KMC_B200_TIME_FROM_AZIMUTH, vec2, h->sm_count, h->stream[slot]));
    return KMC_B200_OK;
  });
}
KMC_CATCH_AT_BOUNDARY("project_frame_host")



// ---- many KITTI .bin files through one overlapped pipeline ------------------------------------------------------------------
extern "C++" {  // templates below: no C linkage inside the extern "C" region
namespace {

// Runs fn(0..n_items-1) on up to n_threads short-lived host threads; returns the first non-zero status.
int ParallelFor(int64_t n_items, int n_threads, const std::function<int(int64_t)>& fn) {
  if (n_items <= 0) return KMC_B200_OK;
  n_threads = static_cast<int>(std::min<int64_t>(std::max(n_threads, 1), n_items));
  std::atomic<int64_t> next{0};
  std::atomic<int> status{KMC_B200_OK};
  std::string message;
  std::mutex message_mu;
  auto body = [&] {
    for (;;) {
      int64_t const i = next.fetch_add(1);
      if (i >= n_items || status.load() != KMC_B200_OK) return;
      int const rc = fn(i);
      if (rc != KMC_B200_OK) {
        std::lock_guard<std::mutex> lock(message_mu);
        if (status.load() == KMC_B200_OK) {
          message = LastError();  // Fail() wrote it on this worker thread
          status.store(rc);
        }
      }
    }
  };
  if (n_threads == 1) {
    body();
  } else {
    std::vector<std::thread> pool;
    struct JoinAll {
      std::vector<std::thread>& pool;
      ~JoinAll() {
        for (auto& th : pool)
          if (th.joinable()) th.join();
      }
    } join_all{pool};
    pool.reserve(static_cast<size_t>(n_threads));
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(body);
    for (auto& th : pool) th.join();
  }
  if (status.load() != KMC_B200_OK) LastError() = message;
  return status.load();
}

// pread / pwrite of a byte range, restarted on EINTR and short transfers.
bool ReadRange(int fd, void* dst, size_t bytes, off_t offset) {
  size_t done = 0;
  while (done < bytes) {
    ssize_t const got = ::pread(fd, static_cast<char*>(dst) + done, bytes - done, offset + static_cast<off_t>(done));
    if (got < 0 && errno == EINTR) continue;
    if (got <= 0) break;
    done += static_cast<size_t>(got);
  }
  return done == bytes;
}

bool WriteRange(int fd, const void* src, size_t bytes, off_t offset) {
  size_t done = 0;
  while (done < bytes) {
    ssize_t const put = ::pwrite(fd, static_cast<const char*>(src) + done, bytes - done, offset + static_cast<off_t>(done));
    if (put < 0 && errno == EINTR) continue;
    if (put <= 0) break;
    done += static_cast<size_t>(put);
  }
  return done == bytes;
}

struct Fd {  // closes on scope exit
  int fd = -1;
  ~Fd() {
    if (fd >= 0) ::close(fd);
  }
  int Close() {
    int const rc = fd >= 0 ? ::close(fd) : 0;
    fd = -1;
    return rc;
  }
};

int ReadWholeFile(const char* path, void* dst, size_t bytes) {
  Fd f{::open(path, O_RDONLY | O_CLOEXEC)};
  if (f.fd < 0) return Fail(KMC_B200_ERR_IO, std::string("unable to open KITTI pointcloud binary file: ") + path);
  if (!ReadRange(f.fd, dst, bytes, 0)) return Fail(KMC_B200_ERR_IO, std::string("short read: ") + path);
  return KMC_B200_OK;
}

int WriteWholeFile(const char* path, const void* src, size_t bytes) {
  Fd f{::open(path, O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644)};
  if (f.fd < 0) return Fail(KMC_B200_ERR_IO, std::string("unable to create output file: ") + path);
  bool const written = WriteRange(f.fd, src, bytes, 0);
  if (f.Close() != 0 || !written) return Fail(KMC_B200_ERR_IO, std::string("short write: ") + path);
  return KMC_B200_OK;
}

// Size of a KITTI .bin in whole points.  The reference's loader accepts any size that is a multiple of 4 bytes and
// truncates a trailing partial point (data_io.cpp:107-112); so does this.
int BinFilePoints(const char* path, int64_t* n_points) {
  struct stat st;
  if (::stat(path, &st) != 0) return Fail(KMC_B200_ERR_IO, std::string("unable to open KITTI pointcloud binary file: ") + path);
  if (st.st_size % 4 != 0)
    return Fail(KMC_B200_ERR_IO, std::string("KITTI pointcloud binary file is incorrectly formatted (size is not a multiple of 4 bytes): ") + path);
  *n_points = static_cast<int64_t>(st.st_size / 16);
  return KMC_B200_OK;
}

// One .bin file of any size through the slots: pread straight into a slot's pinned input buffer, pwrite straight from its
// pinned output buffer — no whole-file buffers.  `launch` as for StreamChunks.
template <class Launch>
int StreamBinFile(kmc_b200_handle* h, const char* path_in, const char* path_out, int64_t n, Launch&& launch) {
  Fd in{::open(path_in, O_RDONLY | O_CLOEXEC)};
  if (in.fd < 0) return Fail(KMC_B200_ERR_IO, std::string("unable to open KITTI pointcloud binary file: ") + path_in);
  Fd out{::open(path_out, O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644)};
  if (out.fd < 0) return Fail(KMC_B200_ERR_IO, std::string("unable to create output file: ") + path_out);
  int rc = StreamChunks(
      h, n, false,
      [&](int slot, int64_t first, int64_t count, const float** src) -> int {
        if (!ReadRange(in.fd, h->h_in[slot], static_cast<size_t>(count) * 16, static_cast<off_t>(first) * 16))
          return Fail(KMC_B200_ERR_IO, std::string("short read: ") + path_in);
        *src = h->h_in[slot];
        return KMC_B200_OK;
      },
      launch, [&](int slot, int64_t) -> float* { return h->h_out[slot]; },
      [&](int slot, int64_t first, int64_t count) -> int {
        if (!WriteRange(out.fd, h->h_out[slot], static_cast<size_t>(count) * 16, static_cast<off_t>(first) * 16))
          return Fail(KMC_B200_ERR_IO, std::string("short write: ") + path_out);
        return KMC_B200_OK;
      });
  if (out.Close() != 0 && rc == KMC_B200_OK) rc = Fail(KMC_B200_ERR_IO, std::string("short write: ") + path_out);
  return rc;
}

}  // namespace
}  // extern "C++"

int kmc_b200_deskew_bin_file(kmc_b200_handle* h, const char* path_in, const char* path_out, const kmc_b200_frame_params* params,
                             int64_t* n_points_out) try {
  TraceRange const trace("kmc_b200_deskew_bin_file");
  if (!h || !path_in || !path_out || !params) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_bin_file: null argument");
  int64_t n = 0;
  if (int rc = BinFilePoints(path_in, &n)) return rc;
  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard const guard(h->device);
  KMC_CUDA_TRY(guard.status());
  kmc_b200_frame_params const P = *params;
  int const rc = StreamBinFile(h, path_in, path_out, n, [&](int slot, int64_t, int64_t count, const float* src, float* dst) -> int {
    auto const cfg = kmc_b200::dev::PickConfig(count, true, false, h->sm_count, false);
    KMC_CUDA_TRY(kmc_b200::dev::LaunchDeskewFrame(src, dst, count, P, KMC_B200_TIME_FROM_AZIMUTH, cfg, h->sm_count, h->stream[slot]));
    return KMC_B200_OK;
  });
  if (rc != KMC_B200_OK) return rc;
  if (n_points_out) *n_points_out = n;
  return KMC_B200_OK;
}
KMC_CATCH_AT_BOUNDARY("deskew_bin_file")

// Files are packed, in order, into groups that fit one staging slot of the handle.  Three slots rotate through
//   read (io_threads x pread straight into the slot's pinned input buffer)  ->  H2D  ->  batched deskew kernel  ->  D2H
//   ->  write (io_threads x pwrite straight from the slot's pinned output buffer),
// the calling thread doing the reads and the launches, one helper thread retiring slots (event wait + writes), so the
// disk / page cache, both PCIe directions and the SMs are busy at the same time.  No intermediate host copies: the .bin
// format is the kernel's input layout.
int kmc_b200_deskew_bin_files(kmc_b200_handle* h, int32_t n_files, const char* const* paths_in, const char* const* paths_out,
                              const kmc_b200_frame_params* params, int mode, int32_t io_threads, int64_t* points_out) try {
  TraceRange const trace("kmc_b200_deskew_bin_files");
  if (!h) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_bin_files: null handle");
  if (n_files < 0) return Fail(KMC_B200_ERR_BAD_SIZE, "deskew_bin_files: negative n_files");
  if (!ValidMode(mode)) return Fail(KMC_B200_ERR_BAD_MODE, "deskew_bin_files: unknown time mode");
  if (n_files == 0) return KMC_B200_OK;
  if (!paths_in || !paths_out || !params) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_bin_files: null argument");
  if (io_threads <= 0) io_threads = static_cast<int32_t>(std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
  constexpr int kSlots = kmc_b200_handle::kSlots;

  // sizes, global offsets, groups
  std::vector<int64_t> offsets(static_cast<size_t>(n_files) + 1, 0);
  for (int32_t f = 0; f < n_files; ++f) {
    if (!paths_in[f] || !paths_out[f]) return Fail(KMC_B200_ERR_NULL_POINTER, "deskew_bin_files: null path");
    int64_t n = 0;
    if (int rc = BinFilePoints(paths_in[f], &n)) return rc;
    offsets[static_cast<size_t>(f) + 1] = offsets[static_cast<size_t>(f)] + n;
    if (points_out) points_out[f] = n;
  }
  struct Group {
    int32_t first_file, end_file;
  };
  std::vector<Group> groups;
  for (int32_t f = 0; f < n_files;) {
    int32_t e = f + 1;
    while (e < n_files && offsets[static_cast<size_t>(e) + 1] - offsets[static_cast<size_t>(f)] <= h->capacity) ++e;
    groups.push_back({f, e});
    f = e;
  }
  int64_t const n_total = offsets[static_cast<size_t>(n_files)];

  std::lock_guard<std::mutex> lock(h->mu);
  DeviceGuard const guard(h->device);
  KMC_CUDA_TRY(guard.status());
  if (int rc = EnsureTables(h, n_files)) return rc;
  KMC_CUDA_TRY(cudaMemcpyAsync(h->d_offsets, offsets.data(), offsets.size() * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream[0]));
  KMC_CUDA_TRY(cudaMemcpyAsync(h->d_params, params, static_cast<size_t>(n_files) * sizeof(kmc_b200_frame_params), cudaMemcpyHostToDevice, h->stream[0]));
  KMC_CUDA_TRY(cudaStreamSynchronize(h->stream[0]));

  // slot hand-over between the submitting (this) thread and the retiring thread
  std::mutex mu;
  std::condition_variable cv;
  bool slot_free[kSlots];
  for (bool& b : slot_free) b = true;
  std::deque<size_t> submitted;  // group indices in submission order
  bool no_more = false;
  int retire_status = KMC_B200_OK;
  std::string retire_message;

  auto retire_group = [&](size_t g) -> int {
    int const slot = static_cast<int>(g % kSlots);
    cudaError_t const e = cudaEventSynchronize(h->done[slot]);
    if (e != cudaSuccess) return FailCuda(e, "cudaEventSynchronize(done)");
    Group const grp = groups[g];
    int64_t const base = offsets[static_cast<size_t>(grp.first_file)];
    int const rc = ParallelFor(grp.end_file - grp.first_file, io_threads, [&](int64_t k) -> int {
      size_t const f = static_cast<size_t>(grp.first_file + k);
      return WriteWholeFile(paths_out[f], h->h_out[slot] + 4 * (offsets[f] - base), static_cast<size_t>(offsets[f + 1] - offsets[f]) * 16);
    });
    if (rc == KMC_B200_OK && h->file_done)  // in file order, from this (the retiring) thread
      for (int32_t f = grp.first_file; f < grp.end_file; ++f)
        h->file_done(f, offsets[static_cast<size_t>(f) + 1] - offsets[static_cast<size_t>(f)], h->file_done_user);
    return rc;
  };
  std::thread retirer([&] {
    cudaSetDevice(h->device);
    for (;;) {
      size_t g;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !submitted.empty() || no_more; });
        if (submitted.empty()) return;
        g = submitted.front();
        submitted.pop_front();
      }
      int rc;
      try {
        rc = retire_group(g);
      } catch (...) {  // nothing may escape a thread function
        rc = kmc_b200::internal::FailException("deskew_bin_files (retiring thread)");
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        if (rc != KMC_B200_OK && retire_status == KMC_B200_OK) {
          retire_status = rc;
          retire_message = LastError();
        }
        slot_free[g % kSlots] = true;
      }
      cv.notify_all();
    }
  });
  struct Joiner {  // whatever path leaves this function — an exception included — the retiring thread is told to stop and joined
    std::thread& t;
    std::mutex& mu;
    std::condition_variable& cv;
    bool& no_more;
    void Join() {
      if (!t.joinable()) return;
      {
        std::lock_guard<std::mutex> lk(mu);
        no_more = true;
      }
      cv.notify_all();
      t.join();
    }
    ~Joiner() { Join(); }
  } joiner{retirer, mu, cv, no_more};

  int status = KMC_B200_OK;
  for (size_t g = 0; g < groups.size() && status == KMC_B200_OK; ++g) {
    int const slot = static_cast<int>(g % kSlots);
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return slot_free[slot]; });
      if (retire_status != KMC_B200_OK) break;
      slot_free[slot] = false;
    }
    Group const grp = groups[g];
    int64_t const base = offsets[static_cast<size_t>(grp.first_file)];
    int64_t const count = offsets[static_cast<size_t>(grp.end_file)] - base;
    if (count > h->capacity) {
      // One file larger than a staging slot (a group of its own): let the pipeline drain, then stream the file through all
      // three slots in capacity-sized chunks (pread into / pwrite from the pinned slots, no whole-file buffer).
      {
        std::unique_lock<std::mutex> lk(mu);
        slot_free[slot] = true;
        cv.wait(lk, [&] { return submitted.empty() && std::all_of(slot_free, slot_free + kSlots, [](bool b) { return b; }); });
        if (retire_status != KMC_B200_OK) break;
      }
      size_t const f = static_cast<size_t>(grp.first_file);
      status = StreamBinFile(h, paths_in[f], paths_out[f], count, [&](int s, int64_t first, int64_t chunk, const float* src, float* dst) -> int {
        auto const cfg = kmc_b200::dev::PickConfig(chunk, true, false, h->sm_count, true);
        KMC_CUDA_TRY(kmc_b200::dev::LaunchDeskewBatch(src, dst, h->d_offsets, h->d_params, n_files, chunk, base + first, n_total, mode, cfg,
                                                      h->sm_count, h->stream[s]));
        return KMC_B200_OK;
      });
      if (status == KMC_B200_OK && h->file_done) h->file_done(grp.first_file, count, h->file_done_user);
      continue;
    }
    status = ParallelFor(grp.end_file - grp.first_file, io_threads, [&](int64_t k) -> int {
      size_t const f = static_cast<size_t>(grp.first_file + k);
      return ReadWholeFile(paths_in[f], h->h_in[slot] + 4 * (offsets[f] - base), static_cast<size_t>(offsets[f + 1] - offsets[f]) * 16);
    });
    auto submit = [&]() -> int {
      if (count > 0) {
        size_t const bytes = static_cast<size_t>(count) * 16;
        KMC_CUDA_TRY(cudaMemcpyAsync(h->d_in[slot], h->h_in[slot], bytes, cudaMemcpyHostToDevice, h->stream[slot]));
        auto const cfg = kmc_b200::dev::PickConfig(count, true, false, h->sm_count, true);
        KMC_CUDA_TRY(kmc_b200::dev::LaunchDeskewBatch(h->d_in[slot], h->d_out[slot], h->d_offsets, h->d_params, n_files, count, base,
                                                      n_total, mode, cfg, h->sm_count, h->stream[slot]));
        KMC_CUDA_TRY(cudaMemcpyAsync(h->h_out[slot], h->d_out[slot], bytes, cudaMemcpyDeviceToHost, h->stream[slot]));
      }
      KMC_CUDA_TRY(cudaEventRecord(h->done[slot], h->stream[slot]));
      return KMC_B200_OK;
    };
    if (status == KMC_B200_OK) status = submit();
    {
      std::lock_guard<std::mutex> lk(mu);
      if (status == KMC_B200_OK) submitted.push_back(g);
      else slot_free[slot] = true;
    }
    cv.notify_all();
  }
  std::string const keep = LastError();
  joiner.Join();
  if (status != KMC_B200_OK) {
    for (int s = 0; s < kSlots; ++s) cudaStreamSynchronize(h->stream[s]);
    cudaGetLastError();
    LastError() = keep;
    return status;
  }
  if (retire_status != KMC_B200_OK) {
    LastError() = retire_message;
    return retire_status;
  }
  return KMC_B200_OK;
}
KMC_CATCH_AT_BOUNDARY("deskew_bin_files")

}  // extern "C"
