// oracle/ref_shim/opencv2/core.hpp -- TEST INFRASTRUCTURE ONLY.
//
// A small FUNCTIONAL stand-in for the slice of <opencv2/core.hpp> that the reference's own sources use, so that
// /root/reference/{Frame,Feature,Feature3D,ShiTomasiFeatureExtractor,ProjectionResidual,CeresBundleAdjustment,
// OpenCV*FeatureExtractor,OpenCVLucasKanadeFM}.cpp compile UNCHANGED, from where they lie, into oracle/_ref/
// (recipe: oracle/ref_build.py) without C++ OpenCV in the image.  Written from OpenCV's documented behaviour;
// nothing here is reference code.  The heavy third-party kernels (calcOpticalFlowPyrLK, goodFeaturesToTrack,
// FAST, blur) are not restated here: the shim forwards them through function-pointer hooks to the real cv2
// wheel (tests install the hooks) or, if no hook is set, to the plain-C oracle.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_64FC3 CV_MAKETYPE(CV_64F, 3)
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> 3) & 511) + 1)

typedef unsigned char uchar;   // OpenCV's cvdef.h puts these in the global namespace (Frame.cpp:65 relies on it)
typedef signed char schar;
typedef unsigned short ushort;

namespace cv {
typedef std::string String;
using ::uchar; using ::schar;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x, T y) : x(x), y(y) {}
    // cv::Point_<int>(Point_<float>) rounds (saturate_cast); same-type and widening copies are exact
    template <typename U> Point_(const Point_<U>& o) : x(conv(o.x)), y(conv(o.y)) {}
private:
    template <typename U> static T conv(U v) { return std::is_integral<T>::value && !std::is_integral<U>::value ? (T)std::lrint((double)v) : (T)v; }
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x, T y, T z) : x(x), y(y), z(z) {}
};
typedef Point3_<float> Point3f;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size& o) const { return width == o.width && height == o.height; } };
struct Size2f { float width, height; Size2f() : width(0), height(0) {} Size2f(float w, float h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int x, int y, int w, int h) : x(x), y(y), width(w), height(h) {} };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {} };
struct TermCriteria { enum { COUNT = 1, EPS = 2 }; int type, maxCount; double epsilon;
    TermCriteria(int t = COUNT + EPS, int c = 30, double e = 0.01) : type(t), maxCount(c), epsilon(e) {} };
struct MatStep { size_t s; MatStep(size_t s = 0) : s(s) {} operator size_t() const { return s; } };

class Mat {
public:
    uchar* data; int rows, cols; MatStep step;
    Mat() : data(nullptr), rows(0), cols(0), step(0), type_(0), whole_rows_(0), whole_cols_(0), datastart_(nullptr) {}
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t stp = 0) : data((uchar*)ext), rows(r), cols(c), type_(type), whole_rows_(r), whole_cols_(c), datastart_((uchar*)ext)
    { step = stp ? stp : (size_t)c * elemSize(); }                                            // user data: not owned, not copied
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }                      // create() zero-fills
    static Mat zeros(Size sz, int type) { return Mat(sz, type); }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step.s + 16, (uchar)0);
        data = datastart_ = buf_->data(); whole_rows_ = r; whole_cols_ = c;
    }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return CV_MAT_CN(type_); }
    size_t elemSize() const { static const int sz[] = {1, 1, 2, 2, 4, 4, 8, 2}; return (size_t)sz[CV_MAT_DEPTH(type_)] * CV_MAT_CN(type_); }
    bool isContinuous() const { return step.s == (size_t)cols * elemSize() || rows == 1; }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step.s, data + (size_t)r * step.s, (size_t)cols * elemSize());
        return m;
    }
    Mat operator()(const Rect& r) const                                                       // view into the parent (shares storage)
    {
        Mat m = *this;
        m.data = data + (size_t)r.y * step.s + (size_t)r.x * elemSize(); m.rows = r.height; m.cols = r.width;
        return m;
    }
    void locateROI(Size& whole, Point& ofs) const
    {
        size_t d = (size_t)(data - datastart_);
        ofs.y = (int)(d / step.s); ofs.x = (int)((d - (size_t)ofs.y * step.s) / elemSize());
        whole = Size(whole_cols_, whole_rows_);
    }
    Mat mul(const Mat& o) const
    {
        if (type_ != CV_64FC1 || o.type_ != CV_64FC1 || rows != o.rows || cols != o.cols) throw std::runtime_error("shim Mat::mul: CV_64FC1 only");
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) { const double *a = ptr<double>(r), *b = o.ptr<double>(r); double* d = m.ptr<double>(r);
            for (int c = 0; c < cols; c++) d[c] = a[c] * b[c]; }
        return m;
    }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step.s); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step.s); }
    template <typename T> T& at(int i, int j) { return ((T*)(data + (size_t)i * step.s))[j]; }
    template <typename T> const T& at(int i, int j) const { return ((const T*)(data + (size_t)i * step.s))[j]; }
    template <typename T> T& at(int i) { return const_cast<T&>(static_cast<const Mat*>(this)->at<T>(i)); }
    template <typename T> const T& at(int i) const
    {
        if (rows == 1) return ((const T*)data)[i];
        if (cols == 1) return *(const T*)(data + (size_t)i * step.s);
        return ((const T*)(data + (size_t)(i / cols) * step.s))[i % cols];
    }
protected:
    int type_, whole_rows_, whole_cols_;
    uchar* datastart_;
    std::shared_ptr<std::vector<uchar>> buf_;
};

template <typename T> struct shim_depth;
template <> struct shim_depth<double> { enum { value = CV_64F }; };
template <> struct shim_depth<float> { enum { value = CV_32F }; };
template <> struct shim_depth<uchar> { enum { value = CV_8U }; };
template <> struct shim_depth<int> { enum { value = CV_32S }; };
template <typename T> class Mat_ : public Mat {
public:
    Mat_(int r, int c) : Mat(r, c, CV_MAKETYPE(shim_depth<T>::value, 1)) {}
    Mat_(int r, int c, T* ext) : Mat(r, c, CV_MAKETYPE(shim_depth<T>::value, 1), ext) {}
};

// cv::OutputArray binds const Mat& (OpenCV writes through it): the reference relies on that (Feature3D.cpp:8, 94)
struct _OutputArray { Mat* m; _OutputArray(const Mat& x) : m(const_cast<Mat*>(&x)) {} };
typedef const _OutputArray& OutputArray;

inline Mat operator-(const Mat& a)
{
    if (a.depth() != CV_64F) throw std::runtime_error("shim operator-: CV_64F only");
    Mat m = a.clone(); const int n = a.cols * a.channels();
    for (int r = 0; r < m.rows; r++) { double* d = m.ptr<double>(r); for (int c = 0; c < n; c++) d[c] = -d[c]; }
    return m;
}
inline Mat operator-(const Mat& a, const Mat& b)
{
    if (a.type() != CV_64FC1 || b.type() != CV_64FC1 || a.rows != b.rows || a.cols != b.cols) throw std::runtime_error("shim operator-: equal-size CV_64FC1 only");
    Mat m = a.clone();
    for (int r = 0; r < m.rows; r++) { double* d = m.ptr<double>(r); const double* s = b.ptr<double>(r); for (int c = 0; c < m.cols; c++) d[c] -= s[c]; }
    return m;
}
inline Mat operator*(double k, const Mat& a)
{
    if (a.type() != CV_64FC1) throw std::runtime_error("shim scalar * Mat: CV_64FC1 only");
    Mat m = a.clone();
    for (int r = 0; r < m.rows; r++) { double* d = m.ptr<double>(r); for (int c = 0; c < m.cols; c++) d[c] *= k; }
    return m;
}
inline void transpose(const Mat& src, OutputArray dst)
{
    if (src.type() != CV_64FC1) throw std::runtime_error("shim transpose: CV_64FC1 only");
    Mat s = src.clone();                                                                       // dst may alias src (CeresBundleAdjustment.cpp:79)
    Mat d(s.cols, s.rows, s.type());
    for (int r = 0; r < s.rows; r++) for (int c = 0; c < s.cols; c++) d.at<double>(c, r) = s.at<double>(r, c);
    if (dst.m->data && dst.m->rows == d.rows && dst.m->cols == d.cols && dst.m->type() == d.type())
        for (int r = 0; r < d.rows; r++) std::memcpy(dst.m->ptr<double>(r), d.ptr<double>(r), sizeof(double) * d.cols);
    else *dst.m = d;
}
inline void split(const Mat& src, std::vector<Mat>& mv)
{
    const int cn = src.channels(); mv.resize(cn);
    if (src.depth() != CV_64F) throw std::runtime_error("shim split: CV_64F only");
    for (int k = 0; k < cn; k++) { mv[k] = Mat(src.rows, src.cols, CV_64FC1);
        for (int r = 0; r < src.rows; r++) { const double* s = src.ptr<double>(r); double* d = mv[k].ptr<double>(r);
            for (int c = 0; c < src.cols; c++) d[c] = s[c * cn + k]; } }
}
inline void merge(const std::vector<Mat>& mv, OutputArray dst)
{
    const int cn = (int)mv.size(); const int rows = mv[0].rows, cols = mv[0].cols;
    Mat d(rows, cols, CV_MAKETYPE(CV_64F, cn));
    for (int k = 0; k < cn; k++) { if (mv[k].type() != CV_64FC1 || mv[k].rows != rows || mv[k].cols != cols) throw std::runtime_error("shim merge: equal-size CV_64FC1 planes only");
        for (int r = 0; r < rows; r++) { const double* s = mv[k].ptr<double>(r); double* o = d.ptr<double>(r);
            for (int c = 0; c < cols; c++) o[c * cn + k] = s[c]; } }
    *dst.m = d;
}
inline void minMaxLoc(const Mat& src, double* mn, double* mx)
{
    if (src.type() != CV_64FC1) throw std::runtime_error("shim minMaxLoc: CV_64FC1 only");
    double lo = HUGE_VAL, hi = -HUGE_VAL;
    for (int r = 0; r < src.rows; r++) { const double* s = src.ptr<double>(r);
        for (int c = 0; c < src.cols; c++) { if (s[c] < lo) lo = s[c]; if (s[c] > hi) hi = s[c]; } }
    if (mn) *mn = lo; if (mx) *mx = hi;
}
inline long long getTickCount() { return 0; }
inline double getTickFrequency() { return 1e9; }
}  // namespace cv
