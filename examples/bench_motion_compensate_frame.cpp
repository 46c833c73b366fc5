// bench_motion_compensate_frame — times the one call a user of the reference makes on this path,
//   kmc::Pointcloud kmc::MotionCompensateFrame(kmc::Frame const&, kmc::Time)      (reference motion_compensation.hpp:13)
// through libkitti_motion_compensation_lib.so, on a KITTI .bin scan: load (KittiPclLoader), pseudo stamps
// (GetPseudoTimeStamps), a frame with a Mercator-magnitude start pose and an aggressive motion (BASELINE config 1), then
// `reps` calls — result allocation, host passes, both copies and the kernel all inside the timed region.
//   bench_motion_compensate_frame <scan.bin> [reps=200] [threads=1]
// With threads > 1 that many host threads call concurrently (the reference function is re-entrant).  Prints one JSON line.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "kitti_motion_compensation/data_io.hpp"
#include "kitti_motion_compensation/lie_algebra.hpp"
#include "kitti_motion_compensation/motion_compensation.hpp"
#include "kitti_motion_compensation/timestamp_mocking.hpp"

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <scan.bin> [reps=200] [threads=1]\n", argv[0]);
    return 2;
  }
  int const reps = argc > 2 ? std::max(1, std::atoi(argv[2])) : 200;
  int const threads = argc > 3 ? std::max(1, std::atoi(argv[3])) : 1;
  kmc::KittiPclLoader loader;
  auto [cloud, intensities] = loader.LoadPointcloud(kmc::Path{argv[1]});
  kmc::Time const start{47072.283701593}, middle{47072.335337762}, end{47072.386973931};  // drive_0005 frame 0
  kmc::VectorXd const stamps{kmc::GetPseudoTimeStamps(cloud, start, end)};
  kmc::Oxts const packet{middle, 49.011212804408, 8.4228850417969, 112.83492279053, 0.022447, 1e-05, -1.2219096732051, 0.0, 0.0, 0.0};
  kmc::Affine3d const T_start{kmc::OxtsToPose(packet)};
  kmc::Twist xi;
  xi << 1.34, 0.03, -0.01, -0.003, 0.004, 0.05;  // 13 m/s, 0.5 rad/s
  kmc::Affine3d const T_end{T_start * kmc::lie::Exp(xi)};
  kmc::Frame const frame{T_start, T_end, kmc::LidarScan{start, middle, end, cloud, intensities, stamps}};
  double const n{static_cast<double>(cloud.rows())};

  for (int i = 0; i < 5; ++i) (void)kmc::MotionCompensateFrame(frame, middle);  // context, handle, staging, pool, allocator
  std::vector<double> stamp_us;
  for (int r = 0; r < reps + 5; ++r) {  // kmc::GetPseudoTimeStamps(Pointcloud const&, Time, Time), timestamp_mocking.hpp:11
    auto const a = std::chrono::steady_clock::now();
    kmc::VectorXd const again{kmc::GetPseudoTimeStamps(cloud, start, end)};
    auto const b = std::chrono::steady_clock::now();
    if (r >= 5) stamp_us.push_back(std::chrono::duration<double, std::micro>(b - a).count());
    if (again(0) != stamps(0)) return 3;
  }
  std::sort(stamp_us.begin(), stamp_us.end());

  // eight frames: a loop of MotionCompensateFrame calls against one kmc::MotionCompensateFrames call (one pipeline)
  std::vector<const kmc::Frame*> const eight(8, &frame);
  std::vector<kmc::Time> const eight_times(8, middle);
  std::vector<double> loop_us, batch_us;
  for (int r = 0; r < 30; ++r) {
    // both variants KEEP their eight results until the end of the repetition (a caller that writes or projects them does);
    // dropping each result at once would let the allocator hand the same warm 4 MB block to every call of the loop
    auto const a = std::chrono::steady_clock::now();
    {
      std::vector<kmc::Pointcloud> kept;
      kept.reserve(8);
      for (int k = 0; k < 8; ++k) kept.push_back(kmc::MotionCompensateFrame(frame, middle));
    }
    auto const b = std::chrono::steady_clock::now();
    {
      std::vector<kmc::Pointcloud> const kept{kmc::MotionCompensateFrames(eight, eight_times)};
    }
    auto const c = std::chrono::steady_clock::now();
    loop_us.push_back(std::chrono::duration<double, std::micro>(b - a).count() / 8);
    batch_us.push_back(std::chrono::duration<double, std::micro>(c - b).count() / 8);
  }
  std::sort(loop_us.begin(), loop_us.end());
  std::sort(batch_us.begin(), batch_us.end());

  std::vector<std::vector<double>> us(static_cast<size_t>(threads));
  double checksum{0.0};
  auto worker = [&](int id) {
    for (int r = 0; r < reps; ++r) {
      auto const a = std::chrono::steady_clock::now();
      kmc::Pointcloud const out{kmc::MotionCompensateFrame(frame, middle)};
      auto const b = std::chrono::steady_clock::now();
      us[static_cast<size_t>(id)].push_back(std::chrono::duration<double, std::micro>(b - a).count());
      if (id == 0 && r == 0) checksum = out(0, 0) + out(out.rows() - 1, 2);
    }
  };
  auto const t0 = std::chrono::steady_clock::now();
  if (threads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(worker, t);
    for (auto& t : pool) t.join();
  }
  double const wall_s{std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()};
  std::vector<double> all;
  for (auto const& v : us) all.insert(all.end(), v.begin(), v.end());
  std::sort(all.begin(), all.end());
  double const median{all[all.size() / 2]};
  std::printf("{\"api\": \"kmc::MotionCompensateFrame(Frame const&, Time)\", \"points\": %.0f, \"reps\": %d, \"threads\": %d, "
              "\"us_median\": %.2f, \"us_min\": %.2f, \"us_p90\": %.2f, \"mpoints_per_s_median_call\": %.2f, "
              "\"mpoints_per_s_aggregate\": %.2f, \"checksum\": %.9f, \"get_pseudo_time_stamps_us_median\": %.2f, "
              "\"eight_frames_loop_us_per_frame\": %.2f, \"eight_frames_batch_call_us_per_frame\": %.2f}\n",
              n, reps, threads, median, all.front(), all[all.size() * 9 / 10], n / median,
              n * static_cast<double>(reps) * threads / wall_s / 1e6, checksum, stamp_us[stamp_us.size() / 2], loop_us[loop_us.size() / 2],
              batch_us[batch_us.size() / 2]);
  return 0;
}
