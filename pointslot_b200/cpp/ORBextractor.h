// ORBextractor.h — drop-in replacement of PointSLOT's include/ORBextractor.h (reference lines 45-111).
//
// Same namespace, class name, constructor, operator() and getters as the reference, so Tracking.cc:385-388
// (`new ORBextractor(nFeatures, fScaleFactor, nLevels, fIniThFAST, fMinThFAST)`) and Frame::ExtractORB
// compile unchanged.  Every result comes from libpslot_orb.so (include/pslot_orb.h); nothing is computed
// on the CPU here.
//
// The public member `mvImagePyramid` (reference :85) keeps its name and its vector-like use
// (`mvImagePyramid[level].rowRange(...)`, `.size()`, iteration) but fills itself LAZILY: the GPU stereo matcher reads
// the pyramid on the device, so the 1.4 MB per image are only downloaded when host code first touches the member after
// an operator() (the reference's Frame::ComputeObjStereoMatches, Frame.cc:2323, 2422-2442, or a viewer).  The levels are
// borderless (the reference's 19-px replicated frame is never read by any consumer).  Constructing with
// keepHostPyramid = true downloads eagerly inside operator() instead.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <list>
#include <vector>

#include "compat/opencv_min.h"

struct pslot_orb;

namespace ORB_SLAM2
{

class ORBextractor;

// std::vector<cv::Mat> look-alike that downloads the device pyramid on first host access after each extraction
class LazyPyramid
{
public:
    typedef std::vector<cv::Mat>::iterator iterator;
    explicit LazyPyramid(ORBextractor *owner) : mpOwner(owner), mbStale(false) {}
    cv::Mat &operator[](size_t i) { Fill(); return mv[i]; }
    cv::Mat &at(size_t i) { Fill(); return mv.at(i); }
    size_t size() const { return mv.size(); }
    bool empty() const { return mv.empty(); }
    void resize(size_t n) { mv.resize(n); }
    iterator begin() { Fill(); return mv.begin(); }
    iterator end() { Fill(); return mv.end(); }
    operator std::vector<cv::Mat> &() { Fill(); return mv; }
    void MarkStale() { mbStale = true; }
    bool IsStale() const { return mbStale; }

private:
    void Fill();
    friend class ORBextractor;
    std::vector<cv::Mat> mv;
    ORBextractor *mpOwner;
    bool mbStale;
};

class ORBextractor
{
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    // extension: CUDA device ordinal and whether operator() mirrors the pyramid into mvImagePyramid
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device,
                 bool keepHostPyramid);
    ~ORBextractor();
    ORBextractor(const ORBextractor &) = delete;
    ORBextractor &operator=(const ORBextractor &) = delete;

    // Compute the ORB features and descriptors on an image.  ORB are dispersed on the image using an
    // octree.  Mask is ignored, exactly as in the reference (ORBextractor.h:58).
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints,
                    cv::OutputArray descriptors);

    // Extension: same as operator() for an interleaved 8-bit colour frame (3 or 4 channels); the
    // cvtColor(RGB/BGR(A) -> GRAY) that Tracking::GrabImageStereo applies first (Tracking.cc:1014-1039) runs on the
    // device, fused with the upload.  bRGB = Tracking's mbRGB (Camera.RGB).
    void ExtractFromColor(const cv::Mat &imColor, int channels, bool bRGB, std::vector<cv::KeyPoint> &keypoints,
                          cv::OutputArray descriptors);

    // Extension: operator() in two halves, so that one host thread keeps both views in flight (the reference spawns two
    // threads per frame for that, Frame.cc:161-164): Begin uploads the image and enqueues the whole extraction on this
    // extractor's stream, End waits and fills keypoints / descriptors exactly as operator() does.
    void Begin(cv::InputArray image);
    void End(std::vector<cv::KeyPoint> &keypoints, cv::OutputArray descriptors);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    LazyPyramid mvImagePyramid;

    // Download the pyramid of the last operator() into mvImagePyramid now (borderless levels); host reads of the member
    // do this by themselves.
    void SyncImagePyramid();
    // The C-ABI context behind this object (used by Frame::ComputeStereoMatches).
    pslot_orb *Handle() const { return mpCtx; }

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<int> mnFeaturesPerLevel;

    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    void Run(const unsigned char *data, int w, int h, long step, int channels, bool rgb,
             std::vector<cv::KeyPoint> &keypoints, cv::OutputArray descriptors);
    void Deliver(int n, std::vector<cv::KeyPoint> &keypoints, cv::OutputArray descriptors);

    pslot_orb *mpCtx;
    bool mbKeepHostPyramid;
    std::vector<unsigned char> mvDescBuf;
};

} // namespace ORB_SLAM2

#endif
