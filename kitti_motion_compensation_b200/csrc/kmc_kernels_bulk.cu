// kmc_kernels_bulk.cu — the single-frame deskew kernel with TMA-engine staging through shared memory
// (cp.async.bulk global->shared with mbarrier completion, compute in place, cp.async.bulk shared->global; SASS UBLKCP).
//
// This is the measured ALTERNATIVE to the register-path kernels in kmc_kernels.cu, not the default.  On B200 both reach
// the same HBM ceiling: interleaved A/B runs (profiles/r01_sweep_ab_bulk_vs_register.log, 10 000 x 130 000 points) give
// 6 573 GB/s for the register path vs 6 331 - 6 595 GB/s for the best staged shapes on the batch, and 6 509 vs
// 6 554 - 6 569 GB/s on a single 520 M-point frame, with torch's copy_ at 6 576 - 6 585 GB/s on the same box.  Like the
// register path, the staged kernels peak when about 32 - 48 KB of loads are in flight per SM and lose 5 - 10 % beyond
// that (profiles/r01_sweep_bulk_*.log).  The register path needs no shared memory, no barriers and no per-tile
// bookkeeping, so it stays the default; the staged kernels are selectable (KMC_B200_TUNE="bulk=1,block=256,unroll=4,
// stages=2,ctas=2") so that the comparison can be re-run on other parts, and they are held bit-identical to the
// default by tests/test_deskew_gpu.py::test_*tma_bulk*.
//
// Pipeline per CTA, STAGES slots of TILE points: the elected thread keeps STAGES-1 bulk loads in flight; all threads wait
// on the slot's mbarrier, deskew their points in place in shared memory, fence the generic->async proxy, barrier; the
// elected thread issues the bulk store of the slot and, before a slot is refilled, waits until the store that last
// used it has finished READING shared memory (cp.async.bulk.wait_group.read).
#include <cstdint>

#include "kmc_kernels.cuh"
#include "kmc_point_math.cuh"

namespace kmc_b200::dev {
namespace {

__device__ __forceinline__ uint32_t SmemAddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void MbarInit(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void MbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void MbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "KMC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra KMC_DONE;\n"
      "bra KMC_WAIT;\n"
      "KMC_DONE:\n"
      "}\n" ::"r"(SmemAddr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void BulkLoad(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(SmemAddr(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(SmemAddr(bar))
               : "memory");
}
__device__ __forceinline__ void BulkStore(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(SmemAddr(smem_src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void BulkWaitRead() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void BulkWaitAll() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void FenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int MODE, int BLOCK, int PTS_PER_THREAD, int STAGES>
__global__ void __launch_bounds__(BLOCK)
    DeskewFrameBulkKernel(const float4* __restrict__ in, float4* __restrict__ out, int64_t n,
                          const __grid_constant__ kmc_b200_frame_params P) {
  constexpr int kTile = BLOCK * PTS_PER_THREAD;            // points per stage
  constexpr uint32_t kTileBytes = kTile * sizeof(float4);  // multiple of 16
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* slots = reinterpret_cast<float4*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(STAGES) * kTileBytes);

  int const tid = threadIdx.x;
  int64_t const n_tiles = n / kTile;  // whole tiles go through the bulk path, the remainder through plain loads
  int64_t const my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) MbarInit(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto tile_ptr = [&](int64_t k) { return (blockIdx.x + k * static_cast<int64_t>(gridDim.x)) * kTile; };

  if (tid == 0) {  // prologue: STAGES-1 loads in flight
    for (int k = 0; k < STAGES - 1 && k < my_tiles; ++k) {
      MbarExpectTx(&full[k], kTileBytes);
      BulkLoad(slots + static_cast<size_t>(k) * kTile, in + tile_ptr(k), kTileBytes, &full[k]);
    }
  }
  for (int64_t k = 0; k < my_tiles; ++k) {
    int const s = static_cast<int>(k % STAGES);
    uint32_t const parity = static_cast<uint32_t>((k / STAGES) & 1);
    if (tid == 0) {
      int64_t const next = k + STAGES - 1;
      if (next < my_tiles) {
        int const ns = static_cast<int>(next % STAGES);  // the slot tile k-1 was stored from
        BulkWaitRead<0>();                               // its store has finished reading shared memory
        MbarExpectTx(&full[ns], kTileBytes);
        BulkLoad(slots + static_cast<size_t>(ns) * kTile, in + tile_ptr(next), kTileBytes, &full[ns]);
      }
    }
    MbarWait(&full[s], parity);
    float4* tile = slots + static_cast<size_t>(s) * kTile;
#pragma unroll
    for (int j = 0; j < PTS_PER_THREAD; ++j) {
      int const i = j * BLOCK + tid;  // consecutive lanes -> consecutive 16-byte words: conflict-free
      tile[i] = DeskewPoint<MODE>(tile[i], P);
    }
    FenceProxyAsync();  // generic-proxy writes -> visible to the async proxy (TMA engine)
    __syncthreads();
    if (tid == 0) BulkStore(out + tile_ptr(k), tile, kTileBytes);
  }
  if (tid == 0) BulkWaitAll();

  // remainder (< one tile): plain 128-bit loads, spread over the grid
  int64_t const rem_begin = n_tiles * kTile;
  for (int64_t i = rem_begin + static_cast<int64_t>(blockIdx.x) * BLOCK + tid; i < n; i += static_cast<int64_t>(gridDim.x) * BLOCK)
    out[i] = DeskewPoint<MODE>(in[i], P);
}

// ---------------------------------------------------------------------------------------------------------------
// Batch of frames stored back to back, staged the same way.  The batch is cut into fixed tiles of BLOCK*PPT points,
// grid-strided over the CTAs.  For every tile the elected thread (a) starts the bulk load of the points, (b) finds the
// frame of the tile's first point in the offset table (guess + verify, binary search as fallback), (c) starts the bulk
// load of that frame's 64-byte record — and of the next non-empty frame's record when the tile crosses a frame
// boundary — into the slot, all completing on the slot's mbarrier.  Consumers read the records from shared memory.
// Tiles that touch more than two frames (frames much smaller than a tile) take a per-point lookup path.
// ---------------------------------------------------------------------------------------------------------------
struct alignas(16) TileMeta {
  int64_t boundary;  // chunk-local index of the first point that is not in `frame` any more
  int32_t kind;      // 0: whole tile in one frame; 1: two frames (records 0 and 1); 2: more than two (per-point lookup)
  int32_t frame;
};

template <int MODE, int BLOCK, int PPT, int STAGES>
__global__ void __launch_bounds__(BLOCK)
    DeskewBatchBulkKernel(const float4* __restrict__ in, float4* __restrict__ out, const int64_t* __restrict__ offsets,
                          const kmc_b200_frame_params* __restrict__ table, int n_frames, int64_t n, int64_t point_base,
                          double frames_per_point) {
  constexpr int kTile = BLOCK * PPT;
  constexpr uint32_t kTileBytes = kTile * sizeof(float4);
  constexpr uint32_t kRecBytes = sizeof(kmc_b200_frame_params);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* slots = reinterpret_cast<float4*>(smem_raw);
  auto* recs = reinterpret_cast<kmc_b200_frame_params*>(smem_raw + static_cast<size_t>(STAGES) * kTileBytes);  // [STAGES][2]
  auto* metas = reinterpret_cast<TileMeta*>(reinterpret_cast<unsigned char*>(recs) + static_cast<size_t>(STAGES) * 2 * kRecBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(metas) + static_cast<size_t>(STAGES) * sizeof(TileMeta));

  int const tid = threadIdx.x;
  int64_t const n_tiles = (n + kTile - 1) / kTile;
  int64_t const my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto tile_first = [&](int64_t k) { return (blockIdx.x + k * static_cast<int64_t>(gridDim.x)) * kTile; };

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) MbarInit(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // elected thread, step 1: which frame(s) does tile k touch?  Pure global loads — done ONE TILE AHEAD of the issue so
  // that their latency (two dependent L2 reads) is hidden behind the tile being processed.
  auto look = [&](int64_t k) {
    TileMeta m;
    m.boundary = 0;
    m.kind = -1;  // no such tile
    m.frame = 0;
    int f1 = 0;
    if (k < my_tiles) {
      int64_t const first = tile_first(k);
      int64_t const count = (n - first < kTile) ? (n - first) : kTile;
      int const f0 = LocateFrame(offsets, n_frames, first + point_base, frames_per_point);
      int64_t const end0 = __ldg(offsets + f0 + 1) - point_base;
      m.boundary = end0;
      m.frame = f0;
      m.kind = 0;
      f1 = f0;
      if (end0 < first + count) {  // the tile runs into the next non-empty frame
        f1 = f0 + 1;
        while (f1 + 1 < n_frames && __ldg(offsets + f1 + 1) - point_base <= end0) ++f1;  // frames ending at end0 are empty
        m.kind = (__ldg(offsets + f1 + 1) - point_base >= first + count) ? 1 : 2;
      }
    }
    struct R {
      TileMeta m;
      int f1;
    };
    return R{m, f1};
  };
  // step 2: fill slot `s` with tile k: points + record(s), all completing on the slot's mbarrier
  auto issue = [&](int64_t k, int s, TileMeta const& m, int f1) {
    int64_t const first = tile_first(k);
    int64_t const count = (n - first < kTile) ? (n - first) : kTile;
    metas[s] = m;  // ordered before the arrive below (release), read after the consumers' wait (acquire)
    MbarExpectTx(&full[s], static_cast<uint32_t>(count) * 16u + (m.kind == 1 ? 2 * kRecBytes : kRecBytes));
    BulkLoad(slots + static_cast<size_t>(s) * kTile, in + first, static_cast<uint32_t>(count) * 16u, &full[s]);
    BulkLoad(&recs[2 * s], table + m.frame, kRecBytes, &full[s]);
    if (m.kind == 1) BulkLoad(&recs[2 * s + 1], table + f1, kRecBytes, &full[s]);
  };

  decltype(look(0)) ahead{};
  if (tid == 0) {
    ahead = look(0);
    for (int k = 0; k < STAGES - 1 && k < my_tiles; ++k) {
      issue(k, k, ahead.m, ahead.f1);
      ahead = look(k + 1);
    }
    if (my_tiles < STAGES - 1) ahead = look(my_tiles);  // nothing left
  }

  for (int64_t k = 0; k < my_tiles; ++k) {
    int const s = static_cast<int>(k % STAGES);
    uint32_t const parity = static_cast<uint32_t>((k / STAGES) & 1);
    if (tid == 0) {
      int64_t const next = k + STAGES - 1;
      if (next < my_tiles) {
        BulkWaitRead<0>();  // the store that last used that slot has finished reading shared memory
        issue(next, static_cast<int>(next % STAGES), ahead.m, ahead.f1);
        ahead = look(next + 1);  // consumed one iteration from now
      }
    }
    MbarWait(&full[s], parity);
    float4* tile = slots + static_cast<size_t>(s) * kTile;
    int64_t const first = tile_first(k);
    int const count = static_cast<int>((n - first < kTile) ? (n - first) : kTile);
    TileMeta const m = metas[s];
    if (m.kind == 0) {
      kmc_b200_frame_params const P = recs[2 * s];
      if (count == kTile) {
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          int const i = j * BLOCK + tid;
          tile[i] = DeskewPoint<MODE>(tile[i], P);
        }
      } else {
        for (int i = tid; i < count; i += BLOCK) tile[i] = DeskewPoint<MODE>(tile[i], P);
      }
    } else if (m.kind == 1) {
      for (int i = tid; i < count; i += BLOCK)
        tile[i] = DeskewPoint<MODE>(tile[i], (first + i < m.boundary) ? recs[2 * s] : recs[2 * s + 1]);
    } else {
      for (int i = tid; i < count; i += BLOCK) {
        int const f = LocateFrame(offsets, n_frames, first + i + point_base, frames_per_point);
        kmc_b200_frame_params const P = table[f];
        tile[i] = DeskewPoint<MODE>(tile[i], P);
      }
    }
    FenceProxyAsync();
    __syncthreads();
    if (tid == 0) BulkStore(out + first, tile, static_cast<uint32_t>(count) * 16u);
  }
  if (tid == 0) BulkWaitAll();
}

template <int MODE, int BLOCK, int PPT, int STAGES>
cudaError_t LaunchBatchBulkT(const float* in, float* out, const int64_t* offsets, const kmc_b200_frame_params* table, int32_t n_frames,
                             int64_t n, int64_t point_base, int64_t n_batch_points, int ctas_per_sm, int sm_count, cudaStream_t stream) {
  constexpr size_t kSmem = static_cast<size_t>(STAGES) * (BLOCK * PPT * sizeof(float4) + 2 * sizeof(kmc_b200_frame_params) +
                                                          sizeof(TileMeta) + sizeof(uint64_t));
  auto kernel = DeskewBatchBulkKernel<MODE, BLOCK, PPT, STAGES>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmem));
  if (e != cudaSuccess) return e;
  int64_t const n_tiles = (n + BLOCK * PPT - 1) / (BLOCK * PPT);
  int64_t grid = static_cast<int64_t>(sm_count) * ctas_per_sm;
  if (grid > n_tiles) grid = n_tiles;
  double const frames_per_point = static_cast<double>(n_frames) / static_cast<double>(n_batch_points);
  kernel<<<static_cast<unsigned>(grid), BLOCK, kSmem, stream>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out),
                                                                offsets, table, n_frames, n, point_base, frames_per_point);
  return cudaGetLastError();
}

template <int MODE, int BLOCK, int PPT, int STAGES>
cudaError_t LaunchBulkT(const float* in, float* out, int64_t n, const kmc_b200_frame_params& P, int ctas_per_sm, int sm_count,
                        cudaStream_t stream) {
  constexpr size_t kSmem = static_cast<size_t>(STAGES) * BLOCK * PPT * sizeof(float4) + STAGES * sizeof(uint64_t);
  auto kernel = DeskewFrameBulkKernel<MODE, BLOCK, PPT, STAGES>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmem));
  if (e != cudaSuccess) return e;
  int64_t const n_tiles = n / (BLOCK * PPT);
  int64_t grid = static_cast<int64_t>(sm_count) * ctas_per_sm;
  if (grid > n_tiles) grid = n_tiles > 0 ? n_tiles : 1;
  kernel<<<static_cast<unsigned>(grid), BLOCK, kSmem, stream>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), n, P);
  return cudaGetLastError();
}

}  // namespace

// shape: block 128/256, points per thread 2/4, stages 3/4 — chosen through KMC_B200_TUNE (bulk=1,block=..,unroll=..,stages=..)
cudaError_t LaunchDeskewFrameBulk(const float* in, float* out, int64_t n, const kmc_b200_frame_params& P, int mode, int block,
                                  int pts_per_thread, int stages, int ctas_per_sm, int sm_count, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
#define KMC_BULK_CASE(B, PPT, S)                                                                                              \
  if (block == B && pts_per_thread == PPT && stages == S)                                                                     \
    return mode == KMC_B200_TIME_FROM_AZIMUTH                                                                                 \
               ? LaunchBulkT<KMC_B200_TIME_FROM_AZIMUTH, B, PPT, S>(in, out, n, P, ctas_per_sm, sm_count, stream)             \
               : LaunchBulkT<KMC_B200_TIME_FROM_W, B, PPT, S>(in, out, n, P, ctas_per_sm, sm_count, stream);
  KMC_BULK_CASE(128, 2, 3)
  KMC_BULK_CASE(128, 2, 4)
  KMC_BULK_CASE(128, 2, 6)
  KMC_BULK_CASE(128, 2, 8)
  KMC_BULK_CASE(128, 4, 3)
  KMC_BULK_CASE(128, 4, 4)
  KMC_BULK_CASE(256, 1, 4)
  KMC_BULK_CASE(256, 1, 8)
  KMC_BULK_CASE(256, 2, 2)
  KMC_BULK_CASE(256, 2, 3)
  KMC_BULK_CASE(256, 2, 4)
  KMC_BULK_CASE(256, 2, 5)
  KMC_BULK_CASE(256, 2, 6)
  KMC_BULK_CASE(256, 2, 8)
  KMC_BULK_CASE(256, 4, 2)
  KMC_BULK_CASE(256, 4, 3)
  KMC_BULK_CASE(256, 4, 4)
  KMC_BULK_CASE(256, 8, 3)
  KMC_BULK_CASE(256, 8, 4)
#undef KMC_BULK_CASE
  return cudaErrorInvalidConfiguration;  // not an instantiated (block, points per thread, stages) shape
}

cudaError_t LaunchDeskewBatchBulk(const float* in, float* out, const int64_t* offsets, const kmc_b200_frame_params* table,
                                  int32_t n_frames, int64_t n, int64_t point_base, int64_t n_batch_points, int mode, int block,
                                  int pts_per_thread, int stages, int ctas_per_sm, int sm_count, cudaStream_t stream) {
  if (n <= 0 || n_frames <= 0) return cudaSuccess;
#define KMC_BBULK_CASE(B, PPT, S)                                                                                                  \
  if (block == B && pts_per_thread == PPT && stages == S)                                                                          \
    return mode == KMC_B200_TIME_FROM_AZIMUTH                                                                                      \
               ? LaunchBatchBulkT<KMC_B200_TIME_FROM_AZIMUTH, B, PPT, S>(in, out, offsets, table, n_frames, n, point_base,         \
                                                                         n_batch_points, ctas_per_sm, sm_count, stream)            \
               : LaunchBatchBulkT<KMC_B200_TIME_FROM_W, B, PPT, S>(in, out, offsets, table, n_frames, n, point_base,               \
                                                                   n_batch_points, ctas_per_sm, sm_count, stream);
  KMC_BBULK_CASE(128, 2, 3)
  KMC_BBULK_CASE(128, 2, 4)
  KMC_BBULK_CASE(128, 4, 2)
  KMC_BBULK_CASE(128, 4, 3)
  KMC_BBULK_CASE(256, 2, 2)
  KMC_BBULK_CASE(256, 2, 3)
  KMC_BBULK_CASE(256, 2, 4)
  KMC_BBULK_CASE(256, 4, 2)
  KMC_BBULK_CASE(256, 4, 3)
  KMC_BBULK_CASE(256, 8, 2)
#undef KMC_BBULK_CASE
  return cudaErrorInvalidConfiguration;
}

}  // namespace kmc_b200::dev
