// ref_glue.cc — C entry points around the REFERENCE's ORB_SLAM2::ORBextractor and the callback plumbing that serves
// its OpenCV primitive calls from cv2.  Test infrastructure only; built ONLY by the opt-in recipe `make -C oracle ref`
// (see oracle/ref_compat/README.md: compiling and running the third-party source needs a maintainer's go-ahead and
// has NOT been done in round 1).
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "ORBextractor.h" // the reference's own header (-I <reference>/include)

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
    g_resize(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step, interpolation);
}

void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int borderType)
{
    dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    g_border(src.data, src.cols, src.rows, src.step, dst.data, dst.step, top, bottom, left, right, borderType);
}

void FAST(const Mat &image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression)
{
    keypoints.clear();
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
    g_blur(dst.data, dst.cols, dst.rows, dst.step, ksize.width, ksize.height, sigmaX, sigmaY, borderType);
}

float fastAtan2(float y, float x) { return g_atan2(y, x); }

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
