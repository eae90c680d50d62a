// mini_cv.h -- the few OpenCV types that appear in the reference's public signatures and in its callers' fit sections
// (cv::Mat with ptr<T> / at<T> / size() / setTo, cv::Size, cv::Point, cv::Vec), for build environments WITHOUT OpenCV (this
// container).  With <opencv2/core.hpp> present the real headers are used and this file is not included.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#ifndef CV_8U
#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_8UC3 16
#define CV_32FC3 21
#endif

namespace cv {
struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Point {
    int x = 0, y = 0;
    Point() {}
    Point(int x_, int y_) : x(x_), y(y_) {}
};
template <class T, int N>
struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    Vec(T a, T b, T c) { static_assert(N == 3, "three-element constructor"); val[0] = a; val[1] = b; val[2] = c; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 3> Vec3f;
typedef Vec<uint8_t, 3> Vec3b;
typedef Vec<int, 3> Vec3i;

class Mat {
   public:
    int rows = 0, cols = 0;
    uint8_t* data = nullptr;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(const Size& s, int type) { create(s.height, s.width, type); }
    static Mat zeros(const Size& s, int type) { return Mat(s, type); }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        buf_ = std::make_shared<std::vector<uint8_t>>((size_t)r * step, (uint8_t)0);
        data = buf_->data();
    }
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_8U ? 1 : type_ == CV_8UC3 ? 3 : type_ == CV_32FC3 ? 12 : 4; }
    bool empty() const { return rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    template <class T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <class T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
    Mat clone() const { Mat m(rows, cols, type_); if (data) std::memcpy(m.data, data, (size_t)rows * step); return m; }
    void setTo(double v) {   // single-channel fill
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                if (type_ == CV_8U) at<uint8_t>(r, c) = (uint8_t)v;
                else if (type_ == CV_32S) at<int32_t>(r, c) = (int32_t)v;
                else if (type_ == CV_32F) at<float>(r, c) = (float)v;
            }
    }
    template <class T, int N> void setTo(const Vec<T, N>& v) {
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) at<Vec<T, N>>(r, c) = v;
    }
   private:
    int type_ = CV_8U;
    std::shared_ptr<std::vector<uint8_t>> buf_;
};
}  // namespace cv
