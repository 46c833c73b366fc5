// kmc_host_pool.hpp — a small persistent pool of host threads for the memory passes of the host entry points
// (layout conversion of the reference's double clouds, staging of pageable buffers into pinned memory).  One memcpy
// thread moves 8-12 GB/s, a quarter of the PCIe link; creating threads per call costs 30-60 us each, which is most of
// a KITTI-size frame's budget — so the workers live as long as their handle and spin for a short while between jobs
// before they go to sleep.  Internal to libkmc_b200.so.
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

namespace kmc_b200::internal {

class HostPool {
 public:
  // n_workers threads besides the caller (0 = everything runs on the calling thread).
  explicit HostPool(int n_workers) {
    for (int i = 0; i < n_workers; ++i) threads_.emplace_back([this] { WorkerLoop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lock(mu_);
      stop_ = true;
      generation_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  HostPool(HostPool const&) = delete;
  HostPool& operator=(HostPool const&) = delete;

  int workers() const { return static_cast<int>(threads_.size()); }

  // Runs fn(i) for every i in [0, n_blocks) on the workers and the calling thread; returns when all have finished.
  // One Run at a time per pool (the owning handle's mutex serialises callers).  fn must not throw.
  template <class F>
  void Run(int64_t n_blocks, F&& fn) {
    if (n_blocks <= 0) return;
    if (threads_.empty() || n_blocks == 1) {
      for (int64_t i = 0; i < n_blocks; ++i) fn(i);
      return;
    }
    auto job = std::make_shared<Job>();
    job->n = n_blocks;
    job->ctx = &fn;
    job->call = [](void* ctx, int64_t i) { (*static_cast<std::remove_reference_t<F>*>(ctx))(i); };
    {
      std::lock_guard<std::mutex> lock(mu_);
      current_ = job;
      generation_.fetch_add(1, std::memory_order_release);
    }
    if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
    Drain(*job);
    // late workers may still be inside a block: wait for the count, not for the workers
    for (int spins = 0; job->done.load(std::memory_order_acquire) < n_blocks; ++spins) {
      if (spins > 2000) std::this_thread::yield();
    }
    // fn dies with this frame; a worker that still holds `job` can only claim indices >= n and never calls it again
  }

 private:
  struct Job {
    int64_t n = 0;
    void* ctx = nullptr;
    void (*call)(void*, int64_t) = nullptr;
    std::atomic<int64_t> next{0};
    std::atomic<int64_t> done{0};
  };

  static void Drain(Job& job) {
    for (;;) {
      int64_t const i = job.next.fetch_add(1, std::memory_order_acq_rel);
      if (i >= job.n) return;
      job.call(job.ctx, i);
      job.done.fetch_add(1, std::memory_order_release);
    }
  }

  void WorkerLoop() {
    uint64_t seen = 0;
    for (;;) {
      // wait for a new generation: spin for ~100 us (the gaps between the passes of one call), then sleep
      uint64_t gen = generation_.load(std::memory_order_acquire);
      if (gen == seen) {
        auto const give_up = std::chrono::steady_clock::now() + std::chrono::microseconds(100);
        for (int spins = 1; gen == seen; ++spins) {
#if defined(__x86_64__)
          __builtin_ia32_pause();
#endif
          gen = generation_.load(std::memory_order_acquire);
          if ((spins & 63) == 0 && std::chrono::steady_clock::now() > give_up) break;
        }
      }
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> lock(mu_);
        if (generation_.load(std::memory_order_acquire) == seen && !stop_) {
          sleepers_.fetch_add(1, std::memory_order_release);
          cv_.wait(lock, [&] { return generation_.load(std::memory_order_acquire) != seen || stop_; });
          sleepers_.fetch_sub(1, std::memory_order_release);
        }
        if (stop_) return;
        seen = generation_.load(std::memory_order_acquire);
        job = current_;
      }
      if (job) Drain(*job);
    }
  }

  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::shared_ptr<Job> current_;
  std::atomic<uint64_t> generation_{0};
  std::atomic<int> sleepers_{0};
  bool stop_ = false;
};

}  // namespace kmc_b200::internal
