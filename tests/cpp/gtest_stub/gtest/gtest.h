// <gtest/gtest.h> for the reference's UNMODIFIED test files (/root/reference/test/*.cpp) in an image without GoogleTest:
// forwards to mini_gtest.hpp, which implements the macros those files use with GoogleTest's meaning.
#pragma once
#include "../../mini_gtest.hpp"
