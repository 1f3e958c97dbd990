// ref_glue.cc — C entry points around the REFERENCE's ORB_SLAM2::ORBextractor and the plumbing that serves its five
// OpenCV primitive calls.  Two servers: (1) C callbacks into the real OpenCV 4.13 inside cv2 (oracle/ref_path.py) — the
// parity pin, OpenCV itself in the loop; (2) when no callbacks are installed, the C oracle's restatements of the same
// primitives (oracle/orb_oracle.c, each pinned bit for bit against cv2 in tests/test_oracle_cv2.py) — native speed, used
// for the multi-threaded CPU baseline timing and for large parity sweeps.  Test infrastructure only; built by
// `make -C oracle ref` into oracle/_ref/ (oracle/ref_compat/README.md).
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "ORBextractor.h" // the reference's own header (-I <reference>/include)
#include "../orb_oracle.h"

extern "C" {
typedef void (*ref_resize_cb)(const uchar *src, int sw, int sh, size_t sstep, uchar *dst, int dw, int dh, size_t dstep, int interp);
typedef void (*ref_border_cb)(const uchar *src, int w, int h, size_t sstep, uchar *dst, size_t dstep, int top, int bottom,
                              int left, int right, int border_type);
typedef int (*ref_fast_cb)(const uchar *img, int w, int h, size_t step, int threshold, int nonmax, float *xyr, int cap);
typedef void (*ref_blur_cb)(uchar *img, int w, int h, size_t step, int kw, int kh, double sx, double sy, int border_type);
typedef float (*ref_atan2_cb)(float y, float x);
}

static ref_resize_cb g_resize;
static ref_border_cb g_border;
static ref_fast_cb g_fast;
static ref_blur_cb g_blur;
static ref_atan2_cb g_atan2;

namespace cv {

void resize(const Mat &src, Mat &dst, Size dsize, double, double, int interpolation)
{
    dst.create(dsize.height, dsize.width, CV_8UC1); // keeps the caller's view when the size already matches
    if (!g_resize) {
        if (interpolation != INTER_LINEAR) std::abort();
        orc_resize_linear_u8(src.data, src.cols, src.rows, (ptrdiff_t)src.step, dst.data, dst.cols, dst.rows, (ptrdiff_t)dst.step);
        return;
    }
    g_resize(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step, interpolation);
}

void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int borderType)
{
    dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    if (!g_border) { // plain data movement: BORDER_REFLECT_101 frame around the source, in place when src views dst
        if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101) std::abort();
        const int w = src.cols, h = src.rows;
        for (int y = 0; y < h; ++y) {
            uchar *d = dst.data + (size_t)(y + top) * dst.step + left;
            if (d != src.data + (size_t)y * src.step) std::memmove(d, src.data + (size_t)y * src.step, (size_t)w);
            for (int x = 1; x <= left; ++x) d[-x] = d[x < w ? x : w - 1];
            for (int x = 0; x < right; ++x) d[w + x] = d[w - 2 - x >= 0 ? w - 2 - x : 0];
        }
        const size_t rowb = (size_t)(w + left + right);
        for (int y = 1; y <= top; ++y)
            std::memcpy(dst.data + (size_t)(top - y) * dst.step, dst.data + (size_t)(top + (y < h ? y : h - 1)) * dst.step, rowb);
        for (int y = 0; y < bottom; ++y)
            std::memcpy(dst.data + (size_t)(top + h + y) * dst.step, dst.data + (size_t)(top + (h - 2 - y >= 0 ? h - 2 - y : 0)) * dst.step, rowb);
        return;
    }
    g_border(src.data, src.cols, src.rows, src.step, dst.data, dst.step, top, bottom, left, right, borderType);
}

void FAST(const Mat &image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression)
{
    keypoints.clear();
    if (!g_fast) {
        if (!nonmaxSuppression) std::abort();
        int cap = 1024;
        std::vector<int> x, y, r;
        for (;;) {
            x.resize(cap); y.resize(cap); r.resize(cap);
            const int n = orc_fast9_16(image.data, image.cols, image.rows, (ptrdiff_t)image.step, threshold, cap, x.data(), y.data(), r.data());
            if (n <= cap) {
                keypoints.resize(n);
                for (int i = 0; i < n; ++i) {
                    KeyPoint &k = keypoints[i];
                    k.pt.x = (float)x[i]; k.pt.y = (float)y[i]; k.size = 7.f; k.angle = -1.f; k.response = (float)r[i];
                    k.octave = 0; k.class_id = -1;
                }
                return;
            }
            cap = n;
        }
    }
    int cap = 4096;
    std::vector<float> buf;
    for (;;) {
        buf.resize((size_t)cap * 3);
        const int n = g_fast(image.data, image.cols, image.rows, image.step, threshold, nonmaxSuppression ? 1 : 0, buf.data(), cap);
        if (n <= cap) {
            keypoints.resize(n);
            for (int i = 0; i < n; ++i) { // cv::FAST emits KeyPoint(x, y, 7.f, -1, response)
                KeyPoint &k = keypoints[i];
                k.pt.x = buf[3 * i]; k.pt.y = buf[3 * i + 1]; k.size = 7.f; k.angle = -1.f; k.response = buf[3 * i + 2];
                k.octave = 0; k.class_id = -1;
            }
            return;
        }
        cap = n;
    }
}

void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double sigmaX, double sigmaY, int borderType)
{
    if (dst.data != src.data) dst = src.clone();
    if (!g_blur) {
        if (ksize.width != 7 || ksize.height != 7 || sigmaX != 2 || sigmaY != 2 || borderType != BORDER_REFLECT_101) std::abort();
        Mat tmp = dst.clone();
        orc_gaussian7x7_s2(tmp.data, tmp.cols, tmp.rows, (ptrdiff_t)tmp.step, dst.data, (ptrdiff_t)dst.step);
        return;
    }
    g_blur(dst.data, dst.cols, dst.rows, dst.step, ksize.width, ksize.height, sigmaX, sigmaY, borderType);
}

float fastAtan2(float y, float x) { return g_atan2 ? g_atan2(y, x) : orc_fast_atan2(y, x); }

void KeyPointsFilter::retainBest(std::vector<KeyPoint> &keypoints, int npoints)
{
    // never reached: the reference calls it only from ComputeKeyPointsOld, which operator() does not use
    (void)keypoints; (void)npoints;
    std::fprintf(stderr, "ref_compat: KeyPointsFilter::retainBest is not provided\n");
    std::abort();
}

} // namespace cv

extern "C" {

void ref_set_callbacks(ref_resize_cb r, ref_border_cb b, ref_fast_cb f, ref_blur_cb g, ref_atan2_cb a)
{
    g_resize = r; g_border = b; g_fast = f; g_blur = g; g_atan2 = a;
}

void *ref_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th)
{
    return new ORB_SLAM2::ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
}

void ref_destroy(void *h) { delete static_cast<ORB_SLAM2::ORBextractor *>(h); }

// ORBextractor::operator()(image, cv::Mat(), keypoints, descriptors); kps = cv::KeyPoint records (28 bytes)
int ref_extract(void *h, const uint8_t *img, int w, int h_, size_t pitch, void *kps, uint8_t *desc32, int cap)
{
    ORB_SLAM2::ORBextractor &E = *static_cast<ORB_SLAM2::ORBextractor *>(h);
    cv::Mat image(h_, w, CV_8UC1, (void *)img, pitch), mask, descriptors;
    std::vector<cv::KeyPoint> keypoints;
    E(image, mask, keypoints, descriptors);
    const int n = (int)keypoints.size();
    if (n > cap) return -n;
    static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint layout");
    std::memcpy(kps, keypoints.data(), sizeof(cv::KeyPoint) * (size_t)n);
    for (int i = 0; i < n; ++i) std::memcpy(desc32 + (size_t)i * 32, descriptors.ptr(i), 32);
    return n;
}

// mvImagePyramid[level] of the last operator()
int ref_level(void *h, int level, uint8_t *dst, size_t dst_pitch, int *w, int *h_)
{
    ORB_SLAM2::ORBextractor &E = *static_cast<ORB_SLAM2::ORBextractor *>(h);
    if (level < 0 || level >= (int)E.mvImagePyramid.size()) return -1;
    const cv::Mat &m = E.mvImagePyramid[level];
    *w = m.cols; *h_ = m.rows;
    if (dst)
        for (int r = 0; r < m.rows; ++r) std::memcpy(dst + (size_t)r * dst_pitch, m.ptr(r), (size_t)m.cols);
    return 0;
}

} // extern "C"
