// motion_compensate_runs — command-line driver with the reference's usage
// (reference examples/motion_compensate_runs.cpp:9-47):  motion_compensate_runs <DATA_DIR> [RUN ...]
// Every run folder <DATA_DIR>/<RUN> gets a velodyne_points/data_motion_compensated/ directory with one deskewed .bin per
// frame (first and last frame copied through).  With no RUN arguments every sub-directory of DATA_DIR is processed, as in
// the reference.  Links against libkitti_motion_compensation_lib.so (this repository's drop-in) + libkmc_b200.so.
#include <algorithm>
#include <filesystem>
#include <iostream>
#include <string>
#include <vector>

#include "kitti_motion_compensation/handlers.hpp"

int main(int argc, char** argv) {
  namespace fs = std::filesystem;
  if (argc < 2) {
    std::cout << "Missing command line arguments.\n\n\tAll runs of a data directory:\n\t\t" << argv[0]
              << " <DATA_DIR>\n\n\tSelected runs:\n\t\t" << argv[0] << " <DATA_DIR> <RUN_1> <RUN_2> ... <RUN_N>\n"
              << std::endl;
    return -1;  // as the reference's main does
  }
  fs::path const data_dir{argv[1]};
  std::vector<std::string> runs(argv + 2, argv + argc);
  if (runs.empty()) {
    // the reference takes EVERY directory entry that is a directory (examples/motion_compensate_runs.cpp:14-19)
    std::cout << "Motion compensating all runs found in data directory: " << data_dir.string() << std::endl;
    for (auto const& entry : fs::directory_iterator(data_dir))
      if (entry.is_directory()) runs.push_back(entry.path().filename().string());
    std::sort(runs.begin(), runs.end());
  } else {
    std::cout << "Motion compensating the following runs in data directory: " << data_dir.string() << std::endl;
  }
  for (std::string const& run : runs) std::cout << "\t" << run << std::endl;
  std::cout << "\n" << std::endl;
  for (std::string const& run : runs) {
    std::cout << "Motion compensating: " << run << std::endl;
    kmc::MotionCompensateRun(data_dir / run);  // the reference concatenates strings (DATA_DIR must end in '/'); a path join accepts both
  }
  return 0;
}
