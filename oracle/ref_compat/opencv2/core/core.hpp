// Stand-in for the OpenCV C++ headers, just wide enough to compile the REFERENCE's own src/ORBextractor.cc unmodified
// (test infrastructure: see oracle/ref_compat/README.md).  This image has OpenCV only as the cv2 Python module, so the
// container types are re-declared here with OpenCV's semantics (reference-counted Mat with ROI views, create() that
// keeps a buffer of matching size, MatExpr assignment in place) and the five image PRIMITIVES the extractor calls -
// resize, copyMakeBorder, FAST, GaussianBlur, fastAtan2 - are forwarded through C callbacks to the real OpenCV 4.13
// inside cv2 (oracle/ref_path.py).  Nothing here implements an OpenCV algorithm.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_PI 3.1415926535897932384626433832795

typedef unsigned char uchar;

namespace cv {

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };

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

struct MatExpr { int rows, cols, type; }; // only Mat::zeros is ever built

class Mat {
public:
    int rows, cols;
    size_t step;
    uchar *data;

    Mat() : rows(0), cols(0), step(0), data(nullptr) {}
    Mat(Size sz, int type) : rows(0), cols(0), step(0), data(nullptr) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr) { create(r, c, type); }
    Mat(int r, int c, int, void *p, size_t step_) : rows(r), cols(c), step(step_ ? step_ : (size_t)c), data((uchar *)p) {}
    Mat(const MatExpr &e) : rows(0), cols(0), step(0), data(nullptr) { *this = e; }

    // OpenCV's create(): a matrix that already has this size and type keeps its buffer (that is what lets the
    // reference write descriptors and pyramid levels through views)
    void create(int r, int c, int)
    {
        if (data && r == rows && c == cols) return;
        rows = r; cols = c; step = (size_t)c;
        buf_ = std::shared_ptr<uchar>(new uchar[(size_t)r * step + 16], std::default_delete<uchar[]>());
        data = buf_.get();
    }
    Mat &operator=(const MatExpr &e)
    {
        create(e.rows, e.cols, e.type);
        for (int r = 0; r < rows; ++r) std::memset(data + (size_t)r * step, 0, (size_t)cols);
        return *this;
    }
    static MatExpr zeros(int r, int c, int type) { MatExpr e = {r, c, type}; return e; }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    size_t step1() const { return step; }
    Mat getMat() const { return *this; }
    Mat clone() const
    {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, CV_8UC1);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols);
        return m;
    }
    Mat operator()(const Rect &r) const { Mat m = *this; m.data = data + (size_t)r.y * step + r.x; m.rows = r.height; m.cols = r.width; return m; }
    Mat rowRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m = *this; m.data = data + a; m.cols = b - a; return m; }
    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T &at(int r, int c) { return ((T *)(data + (size_t)r * step))[c]; }
    template <typename T> const T &at(int r, int c) const { return ((const T *)(data + (size_t)r * step))[c]; }

private:
    std::shared_ptr<uchar> buf_;
};

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

// ---- the primitives: forwarded to OpenCV 4.13 (cv2) through callbacks, see ref_glue.cc -------------------------------
void resize(const Mat &src, Mat &dst, Size dsize, double fx, double fy, int interpolation);
void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int borderType);
void FAST(const Mat &image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression);
void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY, int borderType);
float fastAtan2(float y, float x);

struct KeyPointsFilter { // used only by the reference's dead ComputeKeyPointsOld
    static void retainBest(std::vector<KeyPoint> &keypoints, int npoints);
};

} // namespace cv
