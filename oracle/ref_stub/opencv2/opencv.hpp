// Placeholder for <opencv2/opencv.hpp>, used ONLY by `make -C oracle ref`.
// The reference's data_types.hpp:6 includes OpenCV for its Image struct (a cv::Mat member); the four hot-path
// translation units never touch it.  This stub lets those untouched sources compile on a box that has Eigen3 but no
// OpenCV.  It is test infrastructure, not product code.
#pragma once
namespace cv {
class Mat {};
}  // namespace cv
