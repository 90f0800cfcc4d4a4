// Minimal stand-in for the parts of OpenCV's C++ API that examples/tm_yolov3_tiny_uint8.cpp names, so that the UNMODIFIED example
// source compiles here (the image has no C++ OpenCV) and its post-processing functions can be called by the oracle's pin test
// (oracle/yolo_example_shim.cpp).  TEST INFRASTRUCTURE.  Only cv::Rect_<T> carries semantics the post-processing depends on; it
// restates OpenCV's published definition (modules/core/include/opencv2/core/types.hpp, 4.x: Rect_::area() = width * height;
// operator& = operator&= on a copy: x1 = max(a.x, b.x), y1 = max(a.y, b.y), width = min(a.x + a.width, b.x + b.width) - x1,
// height = min(a.y + a.height, b.y + b.height) - y1, empty rectangle when width <= 0 or height <= 0).  Everything else (Mat,
// imread, ...) exists to satisfy the compiler for code the test never executes.
#pragma once
#include <algorithm>
#include <string>

namespace cv {
template <typename T>
struct Point_
{
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point;
template <typename T>
struct Size_
{
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;
template <typename T>
struct Rect_
{
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    Rect_(const Point_<T>& p, const Size_<T>& s) : x(p.x), y(p.y), width(s.width), height(s.height) {}
    template <typename U>
    Rect_(const Rect_<U>& r) : x((T)r.x), y((T)r.y), width((T)r.width), height((T)r.height) {}
    T area() const { return width * height; }
};
typedef Rect_<int> Rect;
template <typename T>
static inline Rect_<T>& operator&=(Rect_<T>& a, const Rect_<T>& b)
{
    T x1 = std::max(a.x, b.x);
    T y1 = std::max(a.y, b.y);
    a.width = std::min(a.x + a.width, b.x + b.width) - x1;
    a.height = std::min(a.y + a.height, b.y + b.height) - y1;
    a.x = x1;
    a.y = y1;
    if (a.width <= 0 || a.height <= 0) a = Rect_<T>();
    return a;
}
template <typename T>
static inline Rect_<T> operator&(const Rect_<T>& a, const Rect_<T>& b)
{
    Rect_<T> c = a;
    return c &= b;
}
struct Scalar
{
    double v[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a, v[1] = b, v[2] = c, v[3] = d; }
};
struct Mat
{
    int rows, cols;
    unsigned char* data;
    Mat() : rows(0), cols(0), data(nullptr) {}
    bool empty() const { return data == nullptr; }
    int channels() const { return 3; }
    Mat clone() const { return *this; }
    void convertTo(Mat&, int) const {}
    template <typename T>
    T* ptr(int = 0) { return (T*)data; }
    template <typename T>
    const T* ptr(int = 0) const { return (const T*)data; }
};
enum { COLOR_GRAY2RGB = 8, COLOR_BGR2RGB = 4, FONT_HERSHEY_SIMPLEX = 0 };
#define CV_32FC3 21
static inline Mat imread(const std::string&, int = 1) { return Mat(); }
static inline bool imwrite(const std::string&, const Mat&) { return false; }
static inline void cvtColor(const Mat&, Mat&, int) {}
static inline void resize(const Mat&, Mat&, Size) {}
template <typename R>
static inline void rectangle(Mat&, const R&, const Scalar&, int = 1) {}
static inline Size getTextSize(const std::string&, int, double, int, int* base) { if (base) *base = 0; return Size(); }
static inline void putText(Mat&, const std::string&, Point, int, double, Scalar) {}
} // namespace cv
