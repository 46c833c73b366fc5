// kmc_run.cpp — kmc_b200_run_prepare / kmc_b200_motion_compensate_run: a whole KITTI raw run folder through the deskew
// pipeline, split into its host half (text files -> per-frame records, no GPU) and the file pipeline.
//
// Reference: handlers.cpp:41-65 (MotionCompensateRun), :19-39 (first / last frame copy), data_io.cpp:18-66 (LoadTimeStamp,
// LoadOxts), :140-166 (LoadLidarScan), :253-285 (MakeFrame, LoadSingleFrame), utils.cpp:9-41 (id padding, tokenizer, clock
// parsing).  The reference re-opens each time-stamp file for every frame and scans to the wanted line, loads three OxTS
// packets per frame and expands every scan to doubles; here each text file is read once, every OxTS packet is parsed
// once, and the scans go disk -> pinned memory -> GPU -> pinned memory -> disk in their on-disk float32 layout
// (kmc_b200_deskew_bin_files).  Host code only; all arithmetic on the path is behind the C ABI.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "kmc_b200.h"
#include "kmc_internal.hpp"

namespace {

namespace fs = std::filesystem;
using kmc_b200::internal::SetError;

std::string PaddedId(size_t id) {  // utils.cpp:9-14: ten digits, zero padded
  char buf[32];
  std::snprintf(buf, sizeof buf, "%010zu", id);
  return buf;
}

// utils.cpp:16-29 splits on single spaces; "2011-09-26 13:04:32.345" -> token 1 is the clock; utils.cpp:31-41 reads
// hours = chars 0-1, minutes = chars 3-4, seconds = the rest.
bool ClockToSeconds(std::string const& line, double* out) {
  size_t const space = line.find(' ');
  if (space == std::string::npos || line.size() < space + 8) return false;
  std::string const clock = line.substr(space + 1);
  if (clock.size() < 7 || clock[2] != ':' || clock[5] != ':') return false;
  // the reference's std::stoi / std::stod throw on text that does not start with a number (utils.cpp:34-41): malformed here
  auto two_digits = [&](size_t at, long* value) {
    if (clock[at] < '0' || clock[at] > '9' || clock[at + 1] < '0' || clock[at + 1] > '9') return false;
    *value = (clock[at] - '0') * 10 + (clock[at + 1] - '0');
    return true;
  };
  long hours = 0, minutes = 0;
  if (!two_digits(0, &hours) || !two_digits(3, &minutes)) return false;
  std::string const sec = clock.substr(6, 18);
  if (sec.empty() || !((sec[0] >= '0' && sec[0] <= '9') || sec[0] == '.')) return false;
  char* end = nullptr;
  double const seconds = std::strtod(sec.c_str(), &end);
  if (end == sec.c_str()) return false;
  *out = static_cast<double>((60 * hours * 60) + (minutes * 60)) + seconds;
  return true;
}

int LoadTimeStamps(fs::path const& file, size_t n, std::vector<double>* out) {
  std::ifstream in(file);
  if (!in.is_open()) return SetError(KMC_B200_ERR_IO, "failed to open timestamp file: " + file.string());
  out->clear();
  std::string line;
  while (out->size() < n && std::getline(in, line)) {
    double t = 0.0;
    if (!ClockToSeconds(line, &t)) return SetError(KMC_B200_ERR_IO, "malformed time stamp line in " + file.string() + ": " + line);
    out->push_back(t);
  }
  if (out->size() < n) return SetError(KMC_B200_ERR_IO, "timestamp file has fewer lines than there are scans: " + file.string());
  return KMC_B200_OK;
}

// data_io.cpp:37-66: the first six numbers of the packet's single line are lat lon alt roll pitch yaw.
int LoadOxtsPose(fs::path const& file, double T[16]) {
  std::ifstream in(file);
  if (!in.is_open()) return SetError(KMC_B200_ERR_IO, "the Oxts file you tried to load did not open: " + file.string());
  std::string line;
  std::getline(in, line);
  std::istringstream tokens(line);
  double v[6];
  for (double& x : v)
    if (!(tokens >> x)) return SetError(KMC_B200_ERR_IO, "malformed Oxts packet: " + file.string());
  return kmc_b200_oxts_to_pose(v[0], v[1], v[2], v[3], v[4], v[5], 1.0, T);
}

int CopyFile(fs::path const& from, fs::path const& to) {
  std::error_code ec;
  fs::copy_file(from, to, fs::copy_options::overwrite_existing, ec);
  if (ec) return SetError(KMC_B200_ERR_IO, "unable to copy " + from.string() + " to " + to.string() + ": " + ec.message());
  return KMC_B200_OK;
}

double Seconds(std::chrono::steady_clock::time_point a) {
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
}

// Everything MotionCompensateRun needs before the first scan is touched: the number of frames and, for the frames
// 1 .. n-2 that have an OxTS packet on both sides, the per-frame kernel record.
int PrepareRun(fs::path const& run, size_t* n_out, std::vector<kmc_b200_frame_params>* params) {
  fs::path const velodyne{run / "velodyne_points"};
  fs::path const data{velodyne / "data"};
  std::error_code ec;
  if (!fs::is_directory(data, ec)) return SetError(KMC_B200_ERR_IO, "not a KITTI run folder (no velodyne_points/data): " + run.string());
  size_t n = 0;  // handlers.cpp:15-17: every directory entry counts as a frame
  for (auto it = fs::directory_iterator(data, ec); !ec && it != fs::directory_iterator(); it.increment(ec)) ++n;
  if (ec) return SetError(KMC_B200_ERR_IO, "unable to list " + data.string() + ": " + ec.message());
  *n_out = n;
  params->clear();
  if (n < 3) return KMC_B200_OK;
  std::vector<double> start, middle, end, oxts_time;
  if (int rc = LoadTimeStamps(velodyne / "timestamps_start.txt", n, &start)) return rc;
  if (int rc = LoadTimeStamps(velodyne / "timestamps.txt", n, &middle)) return rc;
  if (int rc = LoadTimeStamps(velodyne / "timestamps_end.txt", n, &end)) return rc;
  if (int rc = LoadTimeStamps(run / "oxts" / "timestamps.txt", n, &oxts_time)) return rc;
  std::vector<double> pose(16 * n);
  for (size_t i = 0; i < n; ++i)
    if (int rc = LoadOxtsPose(run / "oxts" / "data" / (PaddedId(i) + ".txt"), &pose[16 * i])) return rc;
  params->resize(n - 2);
  for (size_t i = 1; i + 1 < n; ++i) {
    double T_start[16], T_end[16];
    // MakeFrame: the scan's start lies between packets i-1 and i, its end between packets i and i+1
    int rc = kmc_b200_pose_at_time(oxts_time[i - 1], &pose[16 * (i - 1)], oxts_time[i], &pose[16 * i], start[i], T_start);
    if (rc == KMC_B200_OK) rc = kmc_b200_pose_at_time(oxts_time[i], &pose[16 * i], oxts_time[i + 1], &pose[16 * (i + 1)], end[i], T_end);
    if (rc == KMC_B200_OK) rc = kmc_b200_frame_params_from_poses(T_start, T_end, start[i], end[i], middle[i], &(*params)[i - 1]);
    if (rc < 0) return SetError(rc, "frame " + std::to_string(i) + ": " + kmc_b200_last_error());
  }
  return KMC_B200_OK;
}

}  // namespace

extern "C" int kmc_b200_run_prepare(const char* run_folder, int64_t capacity_frames, kmc_b200_frame_params* params_out,
                                    int64_t* n_frames_out) try {
  if (!run_folder || !n_frames_out) return SetError(KMC_B200_ERR_NULL_POINTER, "run_prepare: null argument");
  size_t n = 0;
  std::vector<kmc_b200_frame_params> params;
  if (int rc = PrepareRun(fs::path{run_folder}, &n, &params)) return rc;
  *n_frames_out = static_cast<int64_t>(n);
  if (params_out) {
    if (capacity_frames < static_cast<int64_t>(params.size()))
      return SetError(KMC_B200_ERR_CAPACITY, "run_prepare: params_out holds fewer than n_frames - 2 records");
    for (size_t k = 0; k < params.size(); ++k) params_out[k] = params[k];
  }
  return KMC_B200_OK;
}
KMC_CATCH_AT_BOUNDARY("run_prepare")

extern "C" int kmc_b200_motion_compensate_run(kmc_b200_handle* h, const char* run_folder, int32_t io_threads,
                                              kmc_b200_run_stats* stats) try {
  if (!h || !run_folder) return SetError(KMC_B200_ERR_NULL_POINTER, "motion_compensate_run: null argument");
  auto const t_begin = std::chrono::steady_clock::now();
  fs::path const run{run_folder};
  fs::path const data{run / "velodyne_points" / "data"};
  fs::path const out_dir{run / "velodyne_points" / "data_motion_compensated"};
  size_t n = 0;
  std::vector<kmc_b200_frame_params> params;
  if (int rc = PrepareRun(run, &n, &params)) return rc;
  std::error_code ec;
  fs::create_directories(out_dir, ec);
  if (ec) return SetError(KMC_B200_ERR_IO, "unable to create " + out_dir.string() + ": " + ec.message());
  kmc_b200_run_stats local{};
  local.frames = static_cast<int64_t>(n);
  if (n == 0) {
    if (stats) *stats = local;
    return KMC_B200_OK;
  }
  // first and last frame: no OxTS packet on one side, copied through byte for byte
  if (int rc = CopyFile(data / (PaddedId(0) + ".bin"), out_dir / (PaddedId(0) + ".bin"))) return rc;
  if (int rc = CopyFile(data / (PaddedId(n - 1) + ".bin"), out_dir / (PaddedId(n - 1) + ".bin"))) return rc;
  local.seconds_prepare = Seconds(t_begin);
  if (!params.empty()) {
    size_t const m = params.size();
    std::vector<std::string> in_paths(m), out_paths(m);
    std::vector<const char*> in_c(m), out_c(m);
    for (size_t k = 0; k < m; ++k) {
      in_paths[k] = (data / (PaddedId(k + 1) + ".bin")).string();
      out_paths[k] = (out_dir / (PaddedId(k + 1) + ".bin")).string();
      in_c[k] = in_paths[k].c_str();
      out_c[k] = out_paths[k].c_str();
    }
    auto const t_pipe = std::chrono::steady_clock::now();
    std::vector<int64_t> points(m);
    if (int rc = kmc_b200_deskew_bin_files(h, static_cast<int32_t>(m), in_c.data(), out_c.data(), params.data(),
                                           KMC_B200_TIME_FROM_AZIMUTH, io_threads, points.data()))
      return rc;
    local.seconds_pipeline = Seconds(t_pipe);
    local.frames_deskewed = static_cast<int64_t>(m);
    for (int64_t p : points) local.points_deskewed += p;
  }
  local.seconds_total = Seconds(t_begin);
  if (stats) *stats = local;
  return KMC_B200_OK;
}
KMC_CATCH_AT_BOUNDARY("motion_compensate_run")
