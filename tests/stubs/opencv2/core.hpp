// Minimal DECLARATION-ONLY stand-in for <opencv2/core.hpp> -- just enough surface for
// `g++ -fsyntax-only` of practical-multi-view_b200/host/pmv_adapters.h against the reference's own
// headers in a container that has no C++ OpenCV.  Not an implementation; never linked.
#pragma once
#include <cstddef>
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <unordered_map>
#include <iostream>
#include <string>
#include <vector>
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC1 5
#define CV_64FC1 6
#define CV_64FC3 22
namespace cv {
typedef std::string String;
typedef unsigned char uchar;
typedef signed char schar;
template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T x, T y) : x(x), y(y) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {} };
typedef Point_<int> Point; typedef Point_<float> Point2f;
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T x, T y, T z) : x(x), y(y), z(z) {} };
typedef Point3_<float> Point3f;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int x, int y, int w, int h) : x(x), y(y), width(w), height(h) {} };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {} };
struct MatStep { size_t s; operator size_t() const { return s; } };
class Mat {
public:
    uchar* data; int rows, cols; MatStep step;
    Mat(); Mat(Size, int); Mat(int, int, int);
    static Mat zeros(int, int, int); static Mat zeros(Size, int);
    Size size() const; bool empty() const; Mat clone() const; int type() const;
    Mat operator()(const Rect&) const; Mat mul(const Mat&) const;
    void locateROI(Size&, Point&) const;
    template <typename T> T& at(int, int = 0); template <typename T> const T& at(int, int = 0) const;
    template <typename T> T* ptr(int = 0); template <typename T> const T* ptr(int = 0) const;
};
Mat operator-(const Mat&);
template <typename T> class Mat_ : public Mat { public: Mat_(int, int); Mat_(int, int, T*); };
long long getTickCount(); double getTickFrequency();
void transpose(const Mat&, Mat&);
}  // namespace cv
