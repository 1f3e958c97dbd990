// Frame.cc — stereo-construction path of PointSLOT's Frame over the C ABI of libpslot_orb.so
// (reference src/Frame.cc:144-200, 1658-1664, 2049-2079, 2142-2316).
#include "Frame.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "../../include/pslot_orb.h"

namespace ORB_SLAM2
{

bool Frame::mbInitialComputations = true;
float Frame::cx, Frame::cy, Frame::fx, Frame::fy, Frame::invfx, Frame::invfy;

Frame::Frame()
    : mpORBvocabulary(nullptr), mpORBextractorLeft(nullptr), mpORBextractorRight(nullptr), mTimeStamp(0), mbf(0),
      mb(0), mThDepth(0), N(0), mnScaleLevels(0), mfScaleFactor(0), mfLogScaleFactor(0)
{
}

Frame::Frame(const cv::Mat &imLeft, const cv::Mat &imRight, const double &timeStamp, ORBextractor *extractorLeft,
             ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef, const float &bf,
             const float &thDepth)
    : mpORBvocabulary(voc), mpORBextractorLeft(extractorLeft), mpORBextractorRight(extractorRight),
      mTimeStamp(timeStamp), mK(K.clone()), mDistCoef(distCoef.clone()), mbf(bf), mb(0), mThDepth(thDepth), N(0)
{
    // Scale Level Info (:152-158)
    mnScaleLevels = mpORBextractorLeft->GetLevels();
    mfScaleFactor = mpORBextractorLeft->GetScaleFactor();
    mfLogScaleFactor = std::log(mfScaleFactor);
    mvScaleFactors = mpORBextractorLeft->GetScaleFactors();
    mvInvScaleFactors = mpORBextractorLeft->GetInverseScaleFactors();
    mvLevelSigma2 = mpORBextractorLeft->GetScaleSigmaSquares();
    mvInvLevelSigma2 = mpORBextractorLeft->GetInverseScaleSigmaSquares();

    // ORB extraction: two host threads drive the two extractor contexts concurrently (:161-164);
    // each context owns its CUDA stream, so the two images overlap on the device
    std::thread threadLeft(&Frame::ExtractORB, this, 0, imLeft);
    std::thread threadRight(&Frame::ExtractORB, this, 1, imRight);
    threadLeft.join();
    threadRight.join();

    N = (int)mvKeys.size();
    if (mvKeys.empty()) return;

    UndistortKeyPoints();

    // The reference calls ComputeStereoMatches() here (:173) while `mb` is still unassigned (it is set at
    // :197), i.e. maxD = mbf/mb reads an indeterminate value.  The intended value mb = mbf/fx is set first
    // (DESIGN.md parity rule 2); fx comes from K exactly as :187 does.
    if (mbInitialComputations) {
        fx = K.at<float>(0, 0);
        fy = K.at<float>(1, 1);
        cx = K.at<float>(0, 2);
        cy = K.at<float>(1, 2);
        invfx = 1.0f / fx;
        invfy = 1.0f / fy;
        mbInitialComputations = false;
    }
    mb = mbf / fx;

    ComputeStereoMatches();
}

void Frame::ExtractORB(int flag, const cv::Mat &im)
{
    if (flag == 0)
        (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors);
    else
        (*mpORBextractorRight)(im, cv::Mat(), mvKeysRight, mDescriptorsRight);
}

void Frame::UndistortKeyPoints()
{
    if (mDistCoef.empty() || mDistCoef.at<float>(0) == 0.0) {
        mvKeysUn = mvKeys;
        return;
    }
#ifdef PSLOT_USE_OPENCV
    // distorted cameras: same cv::undistortPoints call as the reference (:2057-2078); not on the GPU path
    cv::Mat mat(N, 2, CV_32F);
    for (int i = 0; i < N; i++) { mat.at<float>(i, 0) = mvKeys[i].pt.x; mat.at<float>(i, 1) = mvKeys[i].pt.y; }
    mat = mat.reshape(2);
    cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK);
    mat = mat.reshape(1);
    mvKeysUn.resize(N);
    for (int i = 0; i < N; i++) { cv::KeyPoint kp = mvKeys[i]; kp.pt.x = mat.at<float>(i, 0); kp.pt.y = mat.at<float>(i, 1); mvKeysUn[i] = kp; }
#else
    std::fprintf(stderr, "pslot Frame: distorted keypoints need OpenCV (build with -DPSLOT_USE_OPENCV)\n");
    std::abort();
#endif
}

void Frame::ComputeStereoMatches()
{
    mvuRight = std::vector<float>(N, -1.0f);
    mvDepth = std::vector<float>(N, -1.0f);
    if (N == 0) return;
    const float minZ = mb;
    const float maxD = mbf / minZ; // (:2172-2174)

    // cv::Mat descriptor rows are contiguous for matrices made by create(); copy if a caller handed a view
    const cv::Mat &dl = mDescriptors, &dr = mDescriptorsRight;
    std::vector<unsigned char> bl, br;
    const unsigned char *pl = dl.data, *pr = dr.data;
    if (dl.step != 32) { bl.resize((size_t)N * 32); for (int i = 0; i < N; ++i) std::memcpy(&bl[(size_t)i * 32], dl.ptr(i), 32); pl = bl.data(); }
    const int Nr = (int)mvKeysRight.size();
    if (Nr && dr.step != 32) { br.resize((size_t)Nr * 32); for (int i = 0; i < Nr; ++i) std::memcpy(&br[(size_t)i * 32], dr.ptr(i), 32); pr = br.data(); }

    int rc = pslot_stereo_match(mpORBextractorLeft->Handle(), mpORBextractorRight->Handle(),
                                reinterpret_cast<const pslot_keypoint *>(mvKeys.data()), N, pl,
                                reinterpret_cast<const pslot_keypoint *>(mvKeysRight.data()), Nr, pr, mbf, maxD,
                                mvuRight.data(), mvDepth.data());
    if (rc != PSLOT_OK) {
        std::fprintf(stderr, "pslot Frame: pslot_stereo_match failed: %s\n", pslot_last_error());
        std::abort();
    }
}

} // namespace ORB_SLAM2
