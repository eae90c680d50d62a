// oracle/shim/opencv2/core.hpp -- TEST INFRASTRUCTURE.  Container-only stand-ins for the few OpenCV types the
// reference's triangle painters (AvatarHelpers.cpp) touch, so that the reference source itself can be compiled into
// oracle/_ref without OpenCV: cv::Mat as a non-owning 2-D view (ptr<T>(row), at<T>(row, col)), cv::Size, cv::Point2f,
// cv::Vec3i.  No arithmetic lives here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>   // the reference headers rely on OpenCV pulling these in
#include <vector>

// type codes are macros in OpenCV
#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_32FC3 21

namespace cv {

struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Point2f {
    float x = 0.f, y = 0.f;
    Point2f() {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct Vec3i {
    int val[3] = {0, 0, 0};
    Vec3i() {}
    Vec3i(int a, int b, int c) { val[0] = a; val[1] = b; val[2] = c; }
    int operator()(int i) const { return val[i]; }
    int& operator()(int i) { return val[i]; }
    int operator[](int i) const { return val[i]; }
    int& operator[](int i) { return val[i]; }
};

struct Vec3f {
    float val[3] = {0.f, 0.f, 0.f};
    Vec3f() {}
    Vec3f(float a, float b, float c) { val[0] = a; val[1] = b; val[2] = c; }
    float operator[](int i) const { return val[i]; }
    float& operator[](int i) { return val[i]; }
};
struct Vec3b {
    unsigned char val[3] = {0, 0, 0};
    unsigned char operator[](int i) const { return val[i]; }
    unsigned char& operator[](int i) { return val[i]; }
};
struct Vec4d {
    double val[4] = {0, 0, 0, 0};
    double operator[](int i) const { return val[i]; }
    double& operator[](int i) { return val[i]; }
};


class Mat {
public:
    Mat() {}
    Mat(int rows_, int cols_, size_t elem_bytes, void* data_) : rows(rows_), cols(cols_), elem(elem_bytes), data(static_cast<unsigned char*>(data_)) {}
    Mat(Size s, int type) : rows(s.height), cols(s.width), elem(type == CV_32FC3 ? 12 : (type == CV_8U ? 1 : 4)) {   // owning
        own.assign((size_t)rows * cols * elem, 0);
        data = own.data();
    }
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), elem(o.elem), data(o.data), own(o.own) { if (!own.empty()) data = own.data(); }
    Mat(Mat&& o) noexcept : rows(o.rows), cols(o.cols), elem(o.elem), data(o.data), own(std::move(o.own)) { if (!own.empty()) data = own.data(); }
    Mat& operator=(const Mat& o) { rows = o.rows; cols = o.cols; elem = o.elem; own = o.own; data = own.empty() ? o.data : own.data(); return *this; }
    Mat& operator=(Mat&& o) noexcept { rows = o.rows; cols = o.cols; elem = o.elem; own = std::move(o.own); data = own.empty() ? o.data : own.data(); return *this; }
    bool empty() const { return rows == 0 || cols == 0; }
    static Mat zeros(Size s, int type) { return Mat(s, type); }
    template <class V> void setTo(V v) {   // 8-bit and 32-bit integer images only (what AvatarRenderer.cpp fills)
        if (elem == 1) std::memset(data, (int)v, (size_t)rows * cols);
        else for (size_t i = 0; i < (size_t)rows * cols; ++i) reinterpret_cast<int32_t*>(data)[i] = (int32_t)v;
    }
    Size size() const { return Size(cols, rows); }
    template <class T> T* ptr(int r) { return reinterpret_cast<T*>(data + (size_t)r * cols * elem); }
    template <class T> const T* ptr(int r) const { return reinterpret_cast<const T*>(data + (size_t)r * cols * elem); }
    template <class T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    int rows = 0, cols = 0;
    size_t elem = 1;
    unsigned char* data = nullptr;
    std::vector<unsigned char> own;
};

}  // namespace cv
