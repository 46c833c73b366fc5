// motion_compensate_runs — command-line driver with the reference's usage
// (reference examples/motion_compensate_runs.cpp:9-47):  motion_compensate_runs <DATA_DIR> [RUN ...]
// Every run folder <DATA_DIR>/<RUN> gets a velodyne_points/data_motion_compensated/ directory with one deskewed .bin per
// frame (first and last frame copied through).  With no RUN arguments every sub-directory of DATA_DIR that ends in
// "_sync" is processed.  Links against libkitti_motion_compensation_lib.so (this repository's drop-in) + libkmc_b200.so.
#include <algorithm>
#include <filesystem>
#include <iostream>
#include <string>
#include <vector>

#include "kitti_motion_compensation/handlers.hpp"

int main(int argc, char** argv) {
  namespace fs = std::filesystem;
  if (argc < 2) {
    std::cerr << "usage: " << argv[0] << " <DATA_DIR> [RUN ...]\n";
    return 2;
  }
  fs::path const data_dir{argv[1]};
  std::vector<std::string> runs(argv + 2, argv + argc);
  if (runs.empty()) {
    for (auto const& entry : fs::directory_iterator(data_dir)) {
      std::string const name{entry.path().filename().string()};
      if (entry.is_directory() && name.size() > 5 && name.compare(name.size() - 5, 5, "_sync") == 0) runs.push_back(name);
    }
    std::sort(runs.begin(), runs.end());
  }
  for (std::string const& run : runs) {
    std::cout << "Motion compensating run: " << run << std::endl;
    kmc::MotionCompensateRun(data_dir / run);
  }
  return 0;
}
