// test_host_pool.cpp — the library's host thread pool (csrc/kmc_host_pool.hpp) on its own: every block of every job runs
// exactly once, jobs of any size (0, 1, fewer / more blocks than workers), thousands of back-to-back jobs (workers spinning and
// sleeping in between), several caller threads taking turns (what SharedPool does with a mutex), and destruction with sleeping
// workers.  Built with -fsanitize=thread when the toolchain has it (tests/test_host_pool.py).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <mutex>
#include <numeric>
#include <thread>
#include <vector>

#include "kmc_host_pool.hpp"

using kmc_b200::internal::HostPool;

static int failures = 0;
#define CHECK(cond)                                                         \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);         \
      ++failures;                                                           \
    }                                                                       \
  } while (0)

int main() {
  for (int workers : {0, 1, 3, 7}) {
    HostPool pool(workers);
    CHECK(pool.workers() == workers);
    // every index exactly once, for job sizes around the worker count
    for (int64_t n : {0, 1, 2, 3, 7, 8, 9, 64, 1000}) {
      std::vector<std::atomic<int>> hits(static_cast<size_t>(n));
      for (auto& h : hits) h.store(0);
      pool.Run(n, [&](int64_t i) { hits[static_cast<size_t>(i)].fetch_add(1); });
      for (auto& h : hits) CHECK(h.load() == 1);
    }
    // thousands of jobs back to back: results of job k are complete when Run returns
    std::vector<double> data(4096);
    for (int job = 0; job < 3000; ++job) {
      pool.Run(64, [&](int64_t b) {
        for (int64_t i = b * 64; i < (b + 1) * 64; ++i) data[static_cast<size_t>(i)] = static_cast<double>(job) + static_cast<double>(i);
      });
      double sum = std::accumulate(data.begin(), data.end(), 0.0);
      CHECK(sum == 4096.0 * job + 4095.0 * 4096.0 / 2.0);
      if (job % 500 == 0) std::this_thread::sleep_for(std::chrono::microseconds(400));  // let the workers fall asleep
    }
    // several callers taking turns on one pool
    std::mutex turn;
    std::atomic<int64_t> total{0};
    std::vector<std::thread> callers;
    for (int c = 0; c < 4; ++c) {
      callers.emplace_back([&] {
        for (int job = 0; job < 500; ++job) {
          std::vector<int> local(33, 0);
          {
            std::lock_guard<std::mutex> lock(turn);
            pool.Run(33, [&](int64_t i) { local[static_cast<size_t>(i)] = static_cast<int>(i) + 1; });
          }
          total.fetch_add(std::accumulate(local.begin(), local.end(), int64_t{0}));
        }
      });
    }
    for (auto& t : callers) t.join();
    CHECK(total.load() == int64_t{4} * 500 * (33 * 34 / 2));
    std::this_thread::sleep_for(std::chrono::milliseconds(2));  // destruction with sleeping workers
  }
  std::printf("%s\n", failures == 0 ? "host pool ok" : "host pool FAILED");
  return failures == 0 ? 0 : 1;
}
