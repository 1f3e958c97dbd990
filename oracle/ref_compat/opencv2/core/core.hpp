// Stand-in for the OpenCV C++ headers, just wide enough to compile the REFERENCE's own src/ORBextractor.cc unmodified
// and the functions of src/Frame.cc / src/ORBmatcher.cc that oracle/ref_compat/extract_fn.py lifts out of the reference
// tree at build time (test infrastructure: see oracle/ref_compat/README.md).  This image has OpenCV only as the cv2
// Python module, so the container types are re-declared here with OpenCV's semantics (reference-counted Mat with ROI
// views, create() that keeps a buffer of matching size and type, MatExpr assignment in place, convertTo into a fresh
// buffer when the type changes) and the five image PRIMITIVES the extractor calls - resize, copyMakeBorder, FAST,
// GaussianBlur, fastAtan2 - are declared only; ref_glue.cc serves them.  The small float-matrix algebra that
// Frame::ComputeStereoMatches and ORBmatcher::SearchByProjection use (Mat::ones, scalar * Mat, Mat - Mat, Mat * Mat,
// t(), norm L1) is implemented with OpenCV's accumulation types (float data, double accumulators).
#pragma once
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <fstream>   // the real opencv2/core/core.hpp pulls the stream headers in; DBoW2 relies on that
#include <iostream>
#include <sstream>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5
#define CV_PI 3.1415926535897932384626433832795

typedef unsigned char uchar;

namespace cv {

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };

inline int cvRoundImpl(double v) { return (int)std::lrint(v); } // round half to even, as OpenCV's SSE2 path
} // namespace cv
inline int cvRound(double v) { return cv::cvRoundImpl(v); }
inline int cvRound(float v) { return (int)std::lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {
using ::cvRound; using ::cvFloor; using ::cvCeil;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
template <typename T> inline Point_<T> &operator*=(Point_<T> &a, float b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

struct MatExpr { int rows, cols, type; }; // only Mat::zeros is ever built lazily

class Mat {
public:
    int rows, cols;
    size_t step; // bytes per row
    uchar *data;

    Mat() : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) {}
    Mat(Size sz, int type) : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) { create(r, c, type); }
    Mat(int r, int c, int type, void *p, size_t step_ = 0)
        : rows(r), cols(c), step(step_ ? step_ : (size_t)c * esz(type)), data((uchar *)p), type_(type) {}
    Mat(const MatExpr &e) : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) { *this = e; }

    static size_t esz(int type) { return type == CV_32F ? 4 : 1; }
    size_t elemSize() const { return esz(type_); }

    // OpenCV's create(): a matrix that already has this size and type keeps its buffer (that is what lets the
    // reference write descriptors and pyramid levels through views); anything else gets a fresh one
    void create(int r, int c, int type)
    {
        if (data && r == rows && c == cols && type == type_) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * esz(type);
        buf_ = std::shared_ptr<uchar>(new uchar[(size_t)r * step + 16], std::default_delete<uchar[]>());
        data = buf_.get();
    }
    Mat &operator=(const MatExpr &e)
    {
        create(e.rows, e.cols, e.type);
        for (int r = 0; r < rows; ++r) std::memset(data + (size_t)r * step, 0, (size_t)cols * elemSize());
        return *this;
    }
    static MatExpr zeros(int r, int c, int type) { MatExpr e = {r, c, type}; return e; }
    static Mat ones(int r, int c, int type)
    {
        Mat m(r, c, type);
        for (int y = 0; y < r; ++y)
            for (int x = 0; x < c; ++x)
                if (type == CV_32F) m.at<float>(y, x) = 1.f; else m.at<uchar>(y, x) = 1;
        return m;
    }
    static Mat eye(int r, int c, int type)
    {
        Mat m; m = zeros(r, c, type);
        for (int i = 0; i < std::min(r, c); ++i)
            if (type == CV_32F) m.at<float>(i, i) = 1.f; else m.at<uchar>(i, i) = 1;
        return m;
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t step1() const { return step / elemSize(); }
    Mat getMat() const { return *this; }
    Mat clone() const
    {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat &dst) const { dst = clone(); }
    // cv::Mat::push_back(const Mat&): appends the rows of m (Frame::AssignFeatures grows descriptor matrices row by row)
    void push_back(const Mat &m)
    {
        if (m.empty()) return;
        Mat out;
        out.create(rows + m.rows, empty() ? m.cols : cols, empty() ? m.type_ : type_);
        for (int r = 0; r < rows; ++r) std::memcpy(out.data + (size_t)r * out.step, data + (size_t)r * step, (size_t)cols * elemSize());
        for (int r = 0; r < m.rows; ++r)
            std::memcpy(out.data + (size_t)(rows + r) * out.step, m.data + (size_t)r * m.step, (size_t)m.cols * m.elemSize());
        *this = out;
    }
    // cv::Mat::convertTo without scaling; a type change always lands in a fresh buffer (in-place calls on a view,
    // as in Frame.cc:2240, therefore never touch the viewed image)
    void convertTo(Mat &dst, int rtype) const
    {
        Mat out;
        out.create(rows, cols, rtype);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                const float v = type_ == CV_32F ? at<float>(r, c) : (float)at<uchar>(r, c);
                if (rtype == CV_32F) out.at<float>(r, c) = v;
                else out.at<uchar>(r, c) = (uchar)std::min(255, std::max(0, (int)std::lrintf(v)));
            }
        dst = out;
    }
    Mat operator()(const Rect &r) const { Mat m = *this; m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize(); m.rows = r.height; m.cols = r.width; return m; }
    Mat rowRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * elemSize(); m.cols = b - a; return m; }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat col(int c) const { return colRange(c, c + 1); }
    Mat t() const
    {
        Mat m(cols, rows, type_);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c)
                if (type_ == CV_32F) m.at<float>(c, r) = at<float>(r, c); else m.at<uchar>(c, r) = at<uchar>(r, c);
        return m;
    }
    double dot(const Mat &b) const // cv::Mat::dot on CV_32F: double accumulator
    {
        double s = 0;
        for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) s += (double)at<float>(r, c) * (double)b.at<float>(r, c);
        return s;
    }
    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
    template <typename T> T &at(int r, int c) { return ((T *)(data + (size_t)r * step))[c]; }
    template <typename T> const T &at(int r, int c) const { return ((const T *)(data + (size_t)r * step))[c]; }
    // single-index access of a vector-shaped matrix
    template <typename T> T &at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    template <typename T> const T &at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }

private:
    std::shared_ptr<uchar> buf_;
    int type_;
};

// ---- float matrix algebra (CV_32F only), evaluated eagerly -------------------------------------------------------------
inline Mat operator*(float s, const Mat &a)
{
    Mat m(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = a.at<float>(r, c) * s;
    return m;
}
inline Mat operator*(const Mat &a, float s) { return s * a; }
inline Mat operator/(const Mat &a, float s)
{
    Mat m(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = a.at<float>(r, c) / s;
    return m;
}
inline Mat operator-(const Mat &a, const Mat &b)
{
    Mat m(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = a.at<float>(r, c) - b.at<float>(r, c);
    return m;
}
inline Mat operator+(const Mat &a, const Mat &b)
{
    Mat m(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = a.at<float>(r, c) + b.at<float>(r, c);
    return m;
}
inline Mat operator-(const Mat &a) { return -1.f * a; }
inline Mat operator*(const Mat &a, const Mat &b) // cv::gemm on CV_32F accumulates in double
{
    Mat m(a.rows, b.cols, CV_32F);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < b.cols; ++c) {
            double s = 0;
            for (int k = 0; k < a.cols; ++k) s += (double)a.at<float>(r, k) * (double)b.at<float>(k, c);
            m.at<float>(r, c) = (float)s;
        }
    return m;
}
inline double norm(const Mat &a, const Mat &b, int normType) // NORM_L1 of a difference: double accumulator
{
    assert(normType == NORM_L1 && a.type() == CV_32F && b.type() == CV_32F);
    (void)normType;
    double s = 0;
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) s += std::fabs((double)a.at<float>(r, c) - (double)b.at<float>(r, c));
    return s;
}

inline double norm(const Mat &a) // NORM_L2 of a CV_32F matrix: double accumulator, then the square root
{
    double s = 0;
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) s += (double)a.at<float>(r, c) * (double)a.at<float>(r, c);
    return std::sqrt(s);
}

// InputArray / OutputArray: the extractor calls empty()/getMat() on inputs and create()/release()/getMat() on the output
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat &m) : m_(const_cast<Mat *>(&m)) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
protected:
    Mat *m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat &m) { m_ = &m; }
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void release() const { m_->release(); }
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

// ---- cv::FileStorage / cv::FileNode: DBoW2's TemplatedVocabulary has virtual save / load members that mention them (the
// YAML route, Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1500-1716); PointSLOT loads its vocabulary with
// loadFromTextFile, so these stand-ins only have to compile - a storage never opens
class FileNode {
public:
    enum { NONE = 0, SEQ = 5, MAP = 6 };
    FileNode operator[](const char *) const { return FileNode(); }
    FileNode operator[](const std::string &) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    int type() const { return NONE; }
    bool empty() const { return true; }
    operator int() const { return 0; }
    operator double() const { return 0.0; }
    operator float() const { return 0.f; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const char *, int) {}
    FileStorage(const std::string &, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char *) const { return FileNode(); }
    FileNode operator[](const std::string &) const { return FileNode(); }
};
template <typename T> inline FileStorage &operator<<(FileStorage &fs, const T &) { return fs; }

// ---- the primitives: served by ref_glue.cc (cv2 callbacks, or the C oracle's pinned restatements) ---------------------
void resize(const Mat &src, Mat &dst, Size dsize, double fx, double fy, int interpolation);
void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int borderType);
void FAST(const Mat &image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression);
void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY, int borderType);
float fastAtan2(float y, float x);

struct KeyPointsFilter { // used only by the reference's dead ComputeKeyPointsOld
    static void retainBest(std::vector<KeyPoint> &keypoints, int npoints);
};

} // namespace cv
