// utils.hpp — the string/time helpers the frame loaders need (reference: include/.../utils.hpp:8-14,
// src/.../utils.cpp:10-38).  The reference's viz::utils image annotation helper is out of scope (OpenCV).
#pragma once

#include <cstddef>
#include <string>
#include <vector>

namespace kmc {

// 15 -> "0000000015" (KITTI file stems are 10 digits wide)
std::string IdToZeroPaddedString(size_t const id, size_t const pad = 10);

// split on single spaces
std::vector<std::string> TokenizeString(std::string raw_string);

// "13:04:34.309763177" -> seconds since midnight
double MmHhSsToSeconds(std::string const mm_hh_ss);

}  // namespace kmc
