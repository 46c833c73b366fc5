// kmc_internal.hpp — the few symbols the translation units of libkmc_b200.so share that are not part of the C ABI.
#pragma once

#include <string>

namespace kmc_b200::internal {

// Records `what` as the calling thread's kmc_b200_last_error() text and returns `status`.
int SetError(int status, const std::string& what);

}  // namespace kmc_b200::internal
