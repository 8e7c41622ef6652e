// cv_shim.h -- TEST INFRASTRUCTURE: the few pieces of the OpenCV C++ API that the reference's matcher statements touch (OpenCV's
// C++ headers are not in this image), with OpenCV's semantics: reference-counted Mat views (row / rowRange / colRange share
// storage), convertTo u8 -> f32, float Mat arithmetic element by element, norm(a, b, NORM_L1) accumulated in double.  Only used by
// oracle/ref_match.cpp to compile functions taken from /root/reference/src unmodified.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_32F 5

namespace cv {
enum { NORM_L1 = 2 };
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Point { int x = 0, y = 0; Point() {} Point(int x_, int y_) : x(x_), y(y_) {} };
typedef Point Point2i;
typedef unsigned char uchar;
// cvRound / cvFloor / cvCeil as OpenCV's SSE2 builds do them: nearest-even (cvtsd2si), floor, ceil
inline int cvRound(double v) { return (int)std::lrint(v); }
inline int cvRound(float v) { return (int)std::lrint(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }
// cv::fastAtan2 (modules/core/src/mathfuncs_core.simd.hpp, atan_f32): degrees, 7th-order polynomial, all in float
inline float fastAtan2(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s, p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float ax = std::abs(x), ay = std::abs(y);
    float a, c, c2;
    if (ax >= ay) { c = ay / (ax + (float)2.2204460492503131e-16); c2 = c * c; a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    else { c = ax / (ay + (float)2.2204460492503131e-16); c2 = c * c; a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} bool operator==(const Size& o) const { return width == o.width && height == o.height; } };
struct Rect { int x = 0, y = 0, width = 0, height = 0; Rect() {} Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };
inline Point2f& operator*=(Point2f& p, float s) { p.x *= s; p.y *= s; return p; }
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;                                      // bytes per row (public in OpenCV: MatStep)
    size_t step1() const { return step / (type_ == CV_32F ? 4 : 1); }
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    // Mat::zeros(...) is a MatExpr in OpenCV: assigning it to a Mat of the same size and type fills that Mat's OWN storage (a row
    // range of another matrix stays a view of it: src/ORBextractor.cc:1045, 1100-1101)
    struct ZerosExpr { int rows, cols, type; };
    static ZerosExpr zeros(int r, int c, int type) { return ZerosExpr{r, c, type}; }
    Mat(const ZerosExpr& z) { create(z.rows, z.cols, z.type); }
    Mat& operator=(const ZerosExpr& z) {
        if (rows != z.rows || cols != z.cols || type_ != z.type || !buf_) create(z.rows, z.cols, z.type);
        for (int i = 0; i < rows; ++i) std::memset(buf_->data() + off_ + (size_t)i * step_, 0, (size_t)cols * esz());
        return *this;
    }
    Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
    Size size() const { return Size(cols, rows); }
    Mat getMat() const { return *this; }                 // InputArray / OutputArray are plain Mat references here
    void release() { *this = Mat(); }
    unsigned char* ptr(int r) { return buf_->data() + off_ + (size_t)r * step_; }
    unsigned char* ptr_mut(int r) const { return buf_->data() + off_ + (size_t)r * step_; }
    Mat(int r, int c, int type, const void* data) { create(r, c, type); std::memcpy(buf_->data(), data, (size_t)r * step_); }
    static Mat ones(int r, int c, int type) {
        Mat m(r, c, type);
        for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) { if (type == CV_32F) m.at<float>(i, j) = 1.f; else m.at<unsigned char>(i, j) = 1; }
        return m;
    }
    int type() const { return type_; }
    void copyTo(Mat& dst) const { dst = clone(); }
    bool empty() const { return rows == 0 || cols == 0; }
    Mat row(int i) const { return rowRange(i, i + 1); }
    Mat rowRange(int a, int b) const { Mat m = *this; m.off_ += (size_t)a * step_; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m = *this; m.off_ += (size_t)a * esz(); m.cols = b - a; return m; }
    Mat col(int j) const { return colRange(j, j + 1); }
    Mat t() const {                                       // materialised transpose (f32)
        Mat o(cols, rows, CV_32F);
        for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) o.at<float>(j, i) = at<float>(i, j);
        return o;
    }
    Mat clone() const {
        Mat o(rows, cols, type_);
        for (int i = 0; i < rows; ++i) std::memcpy(o.buf_->data() + (size_t)i * o.step_, buf_->data() + off_ + (size_t)i * step_, (size_t)cols * esz());
        return o;
    }
    double dot(const Mat& o) const {                      // cv::Mat::dot on float data: dotProd_<float> accumulates in double
        double s = 0;
        for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) s += (double)at<float>(i, j) * (double)o.at<float>(i, j);
        return s;
    }
    template <class T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }                 // vectors: element i
    template <class T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(buf_->data() + off_ + (size_t)r * step_ + (size_t)c * sizeof(T)); }
    template <class T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(buf_->data() + off_ + (size_t)r * step_ + (size_t)c * sizeof(T)); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(buf_->data() + off_ + (size_t)r * step_); }
    void convertTo(Mat& dst, int type) const {          // u8 / f32 -> f32 (dst may be *this)
        Mat out(rows, cols, type);
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < cols; ++j) out.at<float>(i, j) = type_ == CV_32F ? at<float>(i, j) : (float)at<unsigned char>(i, j);
        dst = out;
    }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type; step_ = (size_t)c * esz(); off_ = 0; step = step_;
        buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * step_);
    }
private:
    size_t esz() const { return type_ == CV_32F ? 4 : 1; }
    std::shared_ptr<std::vector<unsigned char>> buf_;
    size_t off_ = 0, step_ = 0;
    int type_ = CV_8U;
};
inline Mat operator*(float s, const Mat& m) {
    Mat o(m.rows, m.cols, CV_32F);
    for (int i = 0; i < m.rows; ++i) for (int j = 0; j < m.cols; ++j) o.at<float>(i, j) = s * m.at<float>(i, j);
    return o;
}
inline Mat operator-(const Mat& a, const Mat& b) {
    Mat o(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) o.at<float>(i, j) = a.at<float>(i, j) - b.at<float>(i, j);
    return o;
}
// Matrix products of float matrices as cv::gemm evaluates the MatExpr `alpha * A * B (+ C)` for small matrices
// (GEMMSingleMul<float, double>): the dot products and the addition of C in double, one rounding to float at the end.
struct MatScaled { Mat m; double alpha; };
struct MatMul {
    Mat a, b; double alpha;
    Mat eval(const Mat* c) const {
        Mat o(a.rows, b.cols, CV_32F);
        for (int i = 0; i < a.rows; ++i)
            for (int j = 0; j < b.cols; ++j) {
                double s = 0;
                for (int k = 0; k < a.cols; ++k) s += (double)a.at<float>(i, k) * (double)b.at<float>(k, j);
                o.at<float>(i, j) = (float)(alpha * s + (c ? (double)c->at<float>(i, j) : 0.0));
            }
        return o;
    }
    operator Mat() const { return eval(nullptr); }
};
inline MatScaled operator-(const Mat& m) { return MatScaled{m, -1.0}; }
inline MatMul operator*(const Mat& a, const Mat& b) { return MatMul{a, b, 1.0}; }
inline MatMul operator*(const MatScaled& a, const Mat& b) { return MatMul{a.m, b, a.alpha}; }
inline Mat operator+(const MatMul& ab, const Mat& c) { return ab.eval(&c); }

inline double norm(const Mat& a) {                      // NORM_L2 of a float matrix: squares accumulated in double (normL2_<float, double>)
    double s = 0;
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) s += (double)a.at<float>(i, j) * (double)a.at<float>(i, j);
    return std::sqrt(s);
}
inline double norm(const Mat& a, const Mat& b, int /*NORM_L1*/) {   // OpenCV: normDiffL1_<float, double>
    double s = 0;
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) s += std::fabs(a.at<float>(i, j) - b.at<float>(i, j));
    return s;
}
}  // namespace cv
