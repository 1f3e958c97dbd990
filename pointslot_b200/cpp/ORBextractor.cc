// ORBextractor.cc — drop-in replacement of PointSLOT's src/ORBextractor.cc: the class keeps the reference's
// surface (reference include/ORBextractor.h:45-111) and forwards to the C ABI of libpslot_orb.so.
#include "ORBextractor.h"

#include <cassert>
#include <cstdio>
#include <cstdlib>

#include "../../include/pslot_orb.h"

namespace ORB_SLAM2
{

static void die(const char *what)
{
    // the reference has no error channel on this path (no exceptions, no return codes: SURVEY.md §8b);
    // a failing GPU call is a hard configuration error, never silently replaced by a CPU path
    std::fprintf(stderr, "pslot ORBextractor: %s failed: %s\n", what, pslot_last_error());
    std::abort();
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : ORBextractor(_nfeatures, _scaleFactor, _nlevels, _iniThFAST, _minThFAST, 0, false)
{
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST,
                           int device, bool keepHostPyramid)
    : mvImagePyramid(this), nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST),
      minThFAST(_minThFAST), mpCtx(nullptr), mbKeepHostPyramid(keepHostPyramid)
{
    if (pslot_orb_create(&mpCtx, nfeatures, _scaleFactor, nlevels, iniThFAST, minThFAST, device) != PSLOT_OK)
        die("pslot_orb_create");
    mvScaleFactor.resize(nlevels);
    mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    pslot_orb_tables(mpCtx, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                     mvInvLevelSigma2.data(), mnFeaturesPerLevel.data());
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() { pslot_orb_destroy(mpCtx); }

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint> &_keypoints,
                              cv::OutputArray _descriptors)
{
    if (_image.empty()) return; // reference ORBextractor.cc:1046-1047

    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);
    Run(image.data, image.cols, image.rows, (long)image.step, 1, false, _keypoints, _descriptors);
}

void ORBextractor::ExtractFromColor(const cv::Mat &imColor, int channels, bool bRGB, std::vector<cv::KeyPoint> &_keypoints,
                                    cv::OutputArray _descriptors)
{
    if (imColor.empty()) return;
    // with the stand-in cv::Mat the caller describes an interleaved image as rows x (cols * channels) bytes
#ifdef PSLOT_USE_OPENCV
    const int w = imColor.cols;
#else
    const int w = imColor.cols / channels;
#endif
    Run(imColor.data, w, imColor.rows, (long)imColor.step, channels, bRGB, _keypoints, _descriptors);
}

void ORBextractor::Run(const unsigned char *data, int w, int h, long step, int channels, bool rgb,
                       std::vector<cv::KeyPoint> &_keypoints, cv::OutputArray _descriptors)
{
    if (channels == 1) {
        if (pslot_orb_extract_begin(mpCtx, data, w, h, (ptrdiff_t)step) != PSLOT_OK) die("pslot_orb_extract_begin");
        End(_keypoints, _descriptors);
        return;
    }
    const int cap = pslot_orb_max_keypoints(mpCtx);
    _keypoints.resize(cap);
    mvDescBuf.resize((size_t)cap * 32);
    int n = 0;
    auto call = [&](int capacity) {
        pslot_keypoint *k = reinterpret_cast<pslot_keypoint *>(_keypoints.data());
        return pslot_orb_extract_color(mpCtx, data, w, h, (ptrdiff_t)step, channels, rgb ? 1 : 0, k, mvDescBuf.data(), capacity, &n);
    };
    int rc = call(cap);
    if (rc == PSLOT_ERR_CAPACITY) { // the capacity depends on the image aspect ratio; retry with the exact count
        _keypoints.resize(n);
        mvDescBuf.resize((size_t)n * 32);
        rc = call(n);
    }
    if (rc != PSLOT_OK) die("pslot_orb_extract");
    Deliver(n, _keypoints, _descriptors);
}

void ORBextractor::Begin(cv::InputArray _image)
{
    cv::Mat image = _image.getMat();
    if (image.empty()) { // :1046-1047; End() then returns nothing
        if (pslot_orb_extract_begin(mpCtx, nullptr, 0, 0, 0) != PSLOT_OK) die("pslot_orb_extract_begin");
        return;
    }
    if (pslot_orb_extract_begin(mpCtx, image.data, image.cols, image.rows, (ptrdiff_t)image.step) != PSLOT_OK) die("pslot_orb_extract_begin");
}

void ORBextractor::End(std::vector<cv::KeyPoint> &_keypoints, cv::OutputArray _descriptors)
{
    static_assert(sizeof(cv::KeyPoint) == sizeof(pslot_keypoint), "cv::KeyPoint must be 28 bytes");
    int cap = pslot_orb_max_keypoints(mpCtx), n = 0, rc = PSLOT_OK;
    for (int attempt = 0; attempt < 2; ++attempt) {
        _keypoints.resize(cap);
        mvDescBuf.resize((size_t)cap * 32);
        rc = pslot_orb_extract_end(mpCtx, reinterpret_cast<pslot_keypoint *>(_keypoints.data()), mvDescBuf.data(), cap, &n);
        if (rc != PSLOT_ERR_CAPACITY) break; // the capacity depends on the image aspect ratio; retry with the exact count
        cap = n;
    }
    if (rc != PSLOT_OK) die("pslot_orb_extract_end");
    Deliver(n, _keypoints, _descriptors);
}

void ORBextractor::Deliver(int n, std::vector<cv::KeyPoint> &_keypoints, cv::OutputArray _descriptors)
{
    _keypoints.resize(n);
    if (n == 0) {
        _descriptors.release(); // :1065
    } else {
        _descriptors.create(n, 32, CV_8U); // :1068
        cv::Mat d = _descriptors.getMat();
        for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), mvDescBuf.data() + (size_t)i * 32, 32);
    }
    mvImagePyramid.MarkStale(); // the host mirror now lags the device pyramid; refilled on first access
    if (mbKeepHostPyramid) SyncImagePyramid();
}

void LazyPyramid::Fill()
{
    if (mbStale) mpOwner->SyncImagePyramid();
}

void ORBextractor::SyncImagePyramid()
{
    std::vector<cv::Mat> &pyr = mvImagePyramid.mv;
    mvImagePyramid.mbStale = false;
    for (int l = 0; l < nlevels; ++l) {
        int w = 0, h = 0;
        if (pslot_orb_level(mpCtx, l, nullptr, 0, &w, &h) != PSLOT_OK) die("pslot_orb_level");
        pyr[l].create(h, w, CV_8UC1);
        if (pslot_orb_level(mpCtx, l, pyr[l].data, (ptrdiff_t)pyr[l].step, &w, &h) != PSLOT_OK) die("pslot_orb_level");
    }
}

} // namespace ORB_SLAM2
