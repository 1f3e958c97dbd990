// opencv_min.h — the handful of OpenCV types the drop-in classes mention in their signatures.
//
// PointSLOT builds against a system OpenCV 3.x (reference CMakeLists.txt:55); this container has no
// OpenCV C++ headers, so the shim is compiled here against these stand-ins.  In a PointSLOT tree build
// with -DPSLOT_USE_OPENCV and the real headers are used instead; the stand-ins keep the member names and
// layouts the shim relies on (cv::KeyPoint is 28 bytes, cv::Mat exposes data/rows/cols/step).
#pragma once
#ifdef PSLOT_USE_OPENCV
#include <opencv2/core/core.hpp>
#else
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {

typedef unsigned char uchar;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

// Reference-counted dense 2-D matrix, just enough of cv::Mat for 8-bit images, descriptor matrices and
// the 3x3 / 1xN float calibration matrices.
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uchar *data = nullptr;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    // wrap foreign memory (not owned)
    Mat(int r, int c, int type, void *ptr, size_t step_bytes = 0)
        : rows(r), cols(c), step(step_bytes ? step_bytes : (size_t)c * esz(type)), data((uchar *)ptr), type_(type) {}

    void create(int r, int c, int type)
    {
        if (r == rows && c == cols && type == type_ && data && buf_ && buf_.use_count() == 1) return;
        type_ = type; rows = r; cols = c; step = (size_t)c * esz(type);
        buf_ = std::shared_ptr<uchar>(new uchar[(size_t)r * step + 1], std::default_delete<uchar[]>());
        data = buf_.get();
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t elemSize() const { return esz(type_); }
    Mat clone() const
    {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * esz(type_));
        return m;
    }
    Mat row(int r) const { Mat m = *this; m.rows = 1; m.data = data + (size_t)r * step; return m; }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T &at(int r, int c) { return ((T *)(data + (size_t)r * step))[c]; }
    template <typename T> const T &at(int r, int c) const { return ((const T *)(data + (size_t)r * step))[c]; }
    template <typename T> T &at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T &at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }

private:
    static size_t esz(int type) { return type == CV_32F ? 4 : 1; }
    int type_ = CV_8U;
    std::shared_ptr<uchar> buf_;
};

// cv::InputArray / cv::OutputArray are proxy classes in OpenCV; the extractor only ever calls
// empty()/getMat() on the input and create()/release()/getMat() on the output.
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
    _OutputArray() {}
    _OutputArray(Mat &m) { m_ = &m; }
    void create(int r, int c, int type) const { if (m_) m_->create(r, c, type); }
    void release() const { if (m_) m_->release(); }
    Mat &getMatRef() const { return *m_; }
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

} // namespace cv
#endif
