// phase_probe.cu — does HBM3e on B200 move more bytes when the whole GPU alternates between a READ phase and a WRITE
// phase instead of mixing loads and stores?  (write-only fill_ measures 7.2-7.5 TB/s, a 1:1 copy 6.55 TB/s.)
//
// A persistent copy kernel; every thread holds K 256-bit values (K x 32 B) in registers per tile.  mode 0: load tile, store
// tile, no coordination (what the deskew kernel does).  mode 1: loads are only issued while (globaltimer mod period) is in
// the read window, stores only in the write window, so all SMs are in the same phase at the same time without a barrier.
// Build + run on the GPU box:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/phase_probe tools/probes/phase_probe.cu && /tmp/phase_probe
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

struct V8 {
  float f[8];
};
__device__ __forceinline__ V8 Load(const V8* p) {
  V8 v;
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v.f[0]), "=f"(v.f[1]), "=f"(v.f[2]), "=f"(v.f[3]), "=f"(v.f[4]), "=f"(v.f[5]), "=f"(v.f[6]), "=f"(v.f[7])
               : "l"(p));
  return v;
}
__device__ __forceinline__ void Store(V8* p, const V8& v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v.f[0]), "f"(v.f[1]), "f"(v.f[2]), "f"(v.f[3]),
               "f"(v.f[4]), "f"(v.f[5]), "f"(v.f[6]), "f"(v.f[7])
               : "memory");
}
__device__ __forceinline__ unsigned long long GlobalTimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int K, int BLOCK>
__global__ void __launch_bounds__(BLOCK) PhasedCopy(const V8* __restrict__ in, V8* __restrict__ out, long long n_vec, int mode,
                                                    unsigned period_ns, unsigned read_ns) {
  long long const tile = static_cast<long long>(BLOCK) * K;
  long long const n_tiles = n_vec / tile;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const V8* src = in + t * tile + threadIdx.x;
    V8* dst = out + t * tile + threadIdx.x;
    if (mode == 1) {
      while (static_cast<unsigned>(GlobalTimer() % period_ns) >= read_ns) {
      }
    }
    V8 v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = Load(src + k * BLOCK);
    if (mode == 1) {
      while (static_cast<unsigned>(GlobalTimer() % period_ns) < read_ns) {
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      v[k].f[0] += 1.0f;
      Store(dst + k * BLOCK, v[k]);
    }
  }
}

__global__ void TimerSpread(unsigned long long* gt, long long* clk) {
  if (threadIdx.x == 0) {
    gt[blockIdx.x] = GlobalTimer();
    clk[blockIdx.x] = clock64();
  }
}
__global__ void TimerResolution(unsigned long long* out) {
  unsigned long long prev = GlobalTimer(), min_step = ~0ull;
  for (int i = 0; i < 20000; ++i) {
    unsigned long long const t = GlobalTimer();
    if (t != prev) {
      if (t - prev < min_step) min_step = t - prev;
      prev = t;
    }
  }
  out[0] = min_step;
}

template <int K, int BLOCK>
double Run(const V8* in, V8* out, long long n_vec, int ctas_per_sm, int mode, unsigned period_ns, unsigned read_ns, int reps) {
  int sm = 0;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  PhasedCopy<K, BLOCK><<<sm * ctas_per_sm, BLOCK>>>(in, out, n_vec, mode, period_ns, read_ns);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int r = 0; r < reps; ++r) PhasedCopy<K, BLOCK><<<sm * ctas_per_sm, BLOCK>>>(in, out, n_vec, mode, period_ns, read_ns);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  if (cudaGetLastError() != cudaSuccess) return -1;
  long long const tile = static_cast<long long>(BLOCK) * K;
  double const bytes = 2.0 * 32.0 * static_cast<double>((n_vec / tile) * tile);
  return bytes * reps / (ms * 1e-3) / 1e9;
}

int main() {
  long long const n_vec = (8ll << 30) / 32;  // 8 GiB in + 8 GiB out
  V8 *in, *out;
  if (cudaMalloc(&in, n_vec * 32) != cudaSuccess || cudaMalloc(&out, n_vec * 32) != cudaSuccess) {
    std::printf("alloc failed\n");
    return 1;
  }
  cudaMemset(in, 0, n_vec * 32);
  cudaMemset(out, 0, n_vec * 32);
  unsigned long long* gt;
  long long* clk;
  cudaMallocManaged(&gt, 1024 * 8);
  cudaMallocManaged(&clk, 1024 * 8);
  TimerSpread<<<148, 32>>>(gt, clk);
  cudaDeviceSynchronize();
  unsigned long long g0 = ~0ull, g1 = 0;
  long long c0 = (1ll << 62), c1 = -(1ll << 62);
  for (int i = 0; i < 148; ++i) {
    if (gt[i] < g0) g0 = gt[i];
    if (gt[i] > g1) g1 = gt[i];
    if (clk[i] < c0) c0 = clk[i];
    if (clk[i] > c1) c1 = clk[i];
  }
  TimerResolution<<<1, 1>>>(gt);
  cudaDeviceSynchronize();
  std::printf("globaltimer spread over 148 CTAs at launch: %llu ns, resolution %llu ns; clock64 spread: %lld cycles\n", g1 - g0, gt[0], c1 - c0);
  int const reps = 10;
  std::printf("unphased copy (mode 0):\n");
  std::printf("  K=1 128x9 : %7.0f GB/s\n", Run<1, 128>(in, out, n_vec, 9, 0, 0, 0, reps));
  std::printf("  K=4 128x4 : %7.0f GB/s\n", Run<4, 128>(in, out, n_vec, 4, 0, 0, 0, reps));
  std::printf("  K=8 128x4 : %7.0f GB/s\n", Run<8, 128>(in, out, n_vec, 4, 0, 0, 0, reps));
  std::printf("  K=8 256x2 : %7.0f GB/s\n", Run<8, 256>(in, out, n_vec, 2, 0, 0, 0, reps));
  std::printf("  K=8 256x3 : %7.0f GB/s\n", Run<8, 256>(in, out, n_vec, 3, 0, 0, 0, reps));
  std::printf("phased copy (mode 1), read window = half the period:\n");
  unsigned const periods[] = {4000, 6000, 8000, 12000, 16000, 24000, 32000, 64000};
  for (unsigned p : periods) {
    std::printf("  period %6u ns: K=8 128x4 %7.0f   K=8 256x2 %7.0f   K=8 256x3 %7.0f   K=4 128x4 %7.0f GB/s\n", p,
                Run<8, 128>(in, out, n_vec, 4, 1, p, p / 2, reps), Run<8, 256>(in, out, n_vec, 2, 1, p, p / 2, reps),
                Run<8, 256>(in, out, n_vec, 3, 1, p, p / 2, reps), Run<4, 128>(in, out, n_vec, 4, 1, p, p / 2, reps));
  }
  return 0;
}
