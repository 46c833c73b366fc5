// handlers.hpp — run-level drivers (reference: include/.../handlers.hpp:7-15, src/.../handlers.cpp:15-65).
// GenerateProjectionVisualizationOfRun (OpenCV drawing) is out of scope.
#pragma once

#include <cstddef>

#include "kitti_motion_compensation/data_types.hpp"

namespace kmc {

std::size_t NumberOfFilesInDirectory(std::filesystem::path path);

// Frames 0 and n-1 have no oxts packet on both sides and are copied through uncompensated.  The reference writes the
// FIRST cloud under the last id (handlers.cpp:36-38 loads and writes `first_*` twice); this implementation writes
// the last frame's own cloud.
void CopyOverUncompensatedFirstAndLastFrame(Path const run_folder);

// Motion-compensates frames 1 .. n-2 of <run_folder>/velodyne_points/data into .../data_motion_compensated, each to
// its camera trigger time (scan.stamp_middle).  Scans stay in their on-disk float32 xyzi form end to end: file ->
// pinned staging -> fused CUDA kernel (azimuth -> time fused) -> file, no double-precision cloud is materialised.
void MotionCompensateRun(Path const run_folder);

}  // namespace kmc
