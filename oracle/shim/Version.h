// oracle/shim/Version.h -- TEST INFRASTRUCTURE.  Stand-in for the reference's CMake-generated Version.h (Version.h.in): only
// the typedefs Calibration.h needs, so that Calibration.cpp compiles into oracle/_ref.
#pragma once
#include <opencv2/core.hpp>
namespace ark {
typedef cv::Point2f Point2f;
typedef cv::Vec3f Vec3f;
typedef cv::Vec3i Vec3i;
}  // namespace ark
