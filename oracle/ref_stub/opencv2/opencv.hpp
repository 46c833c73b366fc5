// Stand-in for <opencv2/opencv.hpp>, used ONLY by `make -C oracle ref` (this image has no OpenCV C++ headers).
//
// The reference includes OpenCV for its visualisation tools: data_types.hpp:6,64 (cv::Mat inside Image),
// data_io.cpp:229-236 (cv::imread), camera_model.cpp:14,30-31 (Mat::clone, cv::Point, cv::circle, cv::Scalar) and
// utils.cpp:44-53 (vconcat / putText / imwrite).  None of that is pixel arithmetic the hot path depends on, so this stub
// keeps the *calls* and drops the *pictures*:
//   * cv::Mat carries only a size and a DRAW LIST; cv::circle appends (centre, radius, colour) to it instead of rasterising.
//     That is what lets tests read back exactly which integer pixels and colours the reference's own
//     ProjectPointcloudOnImage (camera_model.cpp:5-36) would have drawn.
//   * cv::imread returns an empty Mat (no decoder), cv::imwrite / putText do nothing, vconcat concatenates draw lists.
// TEST INFRASTRUCTURE, not product code, and not a copy of OpenCV: only the names the reference touches exist.
#pragma once

#include <iostream>
#include <memory>
#include <string>
#include <vector>

namespace cv {

struct Point {
  int x = 0, y = 0;
  Point() = default;
  Point(int x_, int y_) : x(x_), y(y_) {}  // the reference passes doubles: implicit double -> int truncation, as cv::Point_<int>
};

struct Scalar {
  double val[4];
  Scalar(double v0 = 0, double v1 = 0, double v2 = 0, double v3 = 0) : val{v0, v1, v2, v3} {}
};

struct DrawnCircle {
  Point center;
  int radius;
  Scalar color;
  int thickness, line_type, shift;
};

class Mat {
 public:
  int rows = 0, cols = 0;
  std::vector<DrawnCircle> drawn;  // what cv::circle was asked to draw, in call order
  Mat() = default;
  Mat(int rows_, int cols_) : rows(rows_), cols(cols_) {}
  Mat clone() const { return *this; }
  bool empty() const { return rows == 0 || cols == 0; }
};

enum ImreadModes { IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
enum HersheyFonts { FONT_HERSHEY_DUPLEX = 2 };
#define CV_RGB(r, g, b) cv::Scalar((b), (g), (r), 0)

inline Mat imread(const std::string&, int = IMREAD_COLOR) { return Mat(); }
inline bool imwrite(const std::string&, const Mat&) { return false; }
inline void circle(Mat& img, Point center, int radius, const Scalar& color, int thickness = 1, int line_type = 8, int shift = 0) {
  img.drawn.push_back(DrawnCircle{center, radius, color, thickness, line_type, shift});
}
inline void putText(Mat&, const std::string&, Point, int, double, Scalar, int = 1) {}
inline void vconcat(const Mat& top, const Mat& bottom, Mat& out) {
  Mat m(top.rows + bottom.rows, top.cols);
  m.drawn = top.drawn;
  m.drawn.insert(m.drawn.end(), bottom.drawn.begin(), bottom.drawn.end());
  out = m;
}

}  // namespace cv
