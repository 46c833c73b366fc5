// latency_probe.cu — what is the fastest way to push ONE KITTI-size scan (130 000 points, 2.08 MB each way) from pinned
// host memory through a streaming kernel and back?  Measures the building blocks and the candidate schedules, wall clock
// around submit..synchronize on the calling thread, median of 300 repetitions each.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/latency_probe tools/probes/latency_probe.cu && /tmp/latency_probe
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <functional>
#include <vector>

#define CK(x)                                                                               \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess) {                                                                \
      std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      return 1;                                                                             \
    }                                                                                       \
  } while (0)

__global__ void Work(const float4* __restrict__ in, float4* __restrict__ out, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = in[i];
    float const s = 0.5f - 0.1591549f * p.y / (fabsf(p.x) + fabsf(p.y) + 1e-6f);
    p.x = fmaf(s, 1.3f, p.x);
    p.y = fmaf(s, 0.02f, p.y);
    p.z = fmaf(s, -0.01f, p.z);
    out[i] = p;
  }
}

static double Median(std::function<void()> const& fn, int reps = 300) {
  for (int i = 0; i < 20; ++i) fn();
  std::vector<double> us;
  for (int i = 0; i < reps; ++i) {
    auto const a = std::chrono::steady_clock::now();
    fn();
    us.push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count());
  }
  std::sort(us.begin(), us.end());
  return us[us.size() / 2];
}

int main(int argc, char** argv) {
  int const n = argc > 1 ? std::atoi(argv[1]) : 130000;
  size_t const bytes = size_t(n) * 16;
  float4 *h_in, *h_out, *d_in, *d_out;
  CK(cudaMallocHost(&h_in, bytes));
  CK(cudaMallocHost(&h_out, bytes));
  CK(cudaMalloc(&d_in, bytes));
  CK(cudaMalloc(&d_out, bytes));
  for (int i = 0; i < n; ++i) h_in[i] = make_float4(1.0f + i % 97, 2.0f - i % 31, 0.5f, 0.25f);
  cudaStream_t s[4];
  cudaEvent_t ev[4];
  for (auto& x : s) CK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
  for (auto& x : ev) CK(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
  int const grid = 148 * 4, block = 256;
  std::printf("n = %d points, %.2f MB each way\n", n, bytes / 1e6);

  auto report = [&](const char* name, double us) { std::printf("%-64s %8.1f us\n", name, us); };
  report("empty: cudaStreamSynchronize on an idle stream", Median([&] { cudaStreamSynchronize(s[0]); }));
  report("H2D only", Median([&] { cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, s[0]); cudaStreamSynchronize(s[0]); }));
  report("D2H only", Median([&] { cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, s[0]); cudaStreamSynchronize(s[0]); }));
  report("kernel only (device -> device)", Median([&] { Work<<<grid, block, 0, s[0]>>>(d_in, d_out, n); cudaStreamSynchronize(s[0]); }));
  report("H2D + kernel + D2H, one stream (1 part)", Median([&] {
           cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, s[0]);
           Work<<<grid, block, 0, s[0]>>>(d_in, d_out, n);
           cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, s[0]);
           cudaStreamSynchronize(s[0]);
         }));
  for (int parts : {2, 3, 4}) {
    char name[96];
    std::snprintf(name, sizeof name, "H2D + kernel + D2H in %d parts over %d streams", parts, std::min(parts, 3));
    report(name, Median([&] {
             int const each = ((n + parts - 1) / parts + 1023) & ~1023;
             for (int k = 0; k < parts; ++k) {
               int const first = k * each, cnt = std::min(each, n - first);
               if (cnt <= 0) break;
               cudaStream_t st = s[k % 3];
               cudaMemcpyAsync(d_in + first, h_in + first, size_t(cnt) * 16, cudaMemcpyHostToDevice, st);
               Work<<<grid, block, 0, st>>>(d_in + first, d_out + first, cnt);
               cudaMemcpyAsync(h_out + first, d_out + first, size_t(cnt) * 16, cudaMemcpyDeviceToHost, st);
             }
             for (int k = 0; k < 3; ++k) cudaStreamSynchronize(s[k]);
           }));
  }
  {  // the 1-part schedule as a CUDA graph
    cudaGraph_t g;
    cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(s[0], cudaStreamCaptureModeThreadLocal));
    cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, s[0]);
    Work<<<grid, block, 0, s[0]>>>(d_in, d_out, n);
    cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, s[0]);
    CK(cudaStreamEndCapture(s[0], &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    report("H2D + kernel + D2H as one CUDA graph launch", Median([&] { cudaGraphLaunch(ge, s[0]); cudaStreamSynchronize(s[0]); }));
  }
  report("zero-copy: kernel reads pinned host, writes pinned host", Median([&] { Work<<<grid, block, 0, s[0]>>>(h_in, h_out, n); cudaStreamSynchronize(s[0]); }));
  for (int g2 : {148, 148 * 2, 148 * 8, 148 * 16}) {
    char name[96];
    std::snprintf(name, sizeof name, "zero-copy, grid %d x 256", g2);
    report(name, Median([&] { Work<<<g2, block, 0, s[0]>>>(h_in, h_out, n); cudaStreamSynchronize(s[0]); }));
  }
  report("hybrid: H2D copy, kernel writes pinned host (1 part)", Median([&] {
           cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, s[0]);
           Work<<<grid, block, 0, s[0]>>>(d_in, h_out, n);
           cudaStreamSynchronize(s[0]);
         }));
  for (int parts : {2, 4}) {
    char name[96];
    std::snprintf(name, sizeof name, "hybrid: H2D copy, kernel writes pinned host (%d parts, 2 streams)", parts);
    report(name, Median([&] {
             int const each = ((n + parts - 1) / parts + 1023) & ~1023;
             for (int k = 0; k < parts; ++k) {
               int const first = k * each, cnt = std::min(each, n - first);
               if (cnt <= 0) break;
               cudaStream_t st = s[k % 2];
               cudaMemcpyAsync(d_in + first, h_in + first, size_t(cnt) * 16, cudaMemcpyHostToDevice, st);
               Work<<<grid, block, 0, st>>>(d_in + first, h_out + first, cnt);
             }
             cudaStreamSynchronize(s[0]);
             cudaStreamSynchronize(s[1]);
           }));
  }
  report("hybrid: kernel reads pinned host, D2H copy (1 part)", Median([&] {
           Work<<<grid, block, 0, s[0]>>>(h_in, d_out, n);
           cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, s[0]);
           cudaStreamSynchronize(s[0]);
         }));
  report("split: half zero-copy kernel, half copy-engine pipeline, concurrently", Median([&] {
           int const half = (n / 2) & ~1023;
           Work<<<grid, block, 0, s[0]>>>(h_in, h_out, half);
           cudaMemcpyAsync(d_in + half, h_in + half, size_t(n - half) * 16, cudaMemcpyHostToDevice, s[1]);
           Work<<<grid, block, 0, s[1]>>>(d_in + half, d_out + half, n - half);
           cudaMemcpyAsync(h_out + half, d_out + half, size_t(n - half) * 16, cudaMemcpyDeviceToHost, s[1]);
           cudaStreamSynchronize(s[0]);
           cudaStreamSynchronize(s[1]);
         }));
  return 0;
}
