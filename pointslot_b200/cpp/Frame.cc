// Frame.cc — stereo-construction path of PointSLOT's Frame over the C ABI of libpslot_orb.so
// (reference src/Frame.cc:144-200, 1636-1664, 1808-1861, 2027-2038, 2049-2079, 2112-2140, 2142-2316).
#include "Frame.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "../../include/pslot_orb.h"

namespace ORB_SLAM2
{

bool Frame::mbInitialComputations = true;
float Frame::cx, Frame::cy, Frame::fx, Frame::fy, Frame::invfx, Frame::invfy;
float Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;

Frame::Frame()
    : mpORBvocabulary(nullptr), mpORBextractorLeft(nullptr), mpORBextractorRight(nullptr), mTimeStamp(0), mbf(0),
      mb(0), mThDepth(0), N(0), mnScaleLevels(0), mfScaleFactor(0), mfLogScaleFactor(0)
{
}

Frame::ObjectStageFn Frame::ObjectStageHook = nullptr;

// The path every stereo constructor of the reference runs (:152-199, and the same lines of :203-760).
void Frame::StereoFrontEnd(const cv::Mat &imLeft, const cv::Mat &imRight, cv::Mat &K, cv::Mat *imColor, int objectMode)
{
    // Scale Level Info (:152-158)
    mnScaleLevels = mpORBextractorLeft->GetLevels();
    mfScaleFactor = mpORBextractorLeft->GetScaleFactor();
    mfLogScaleFactor = std::log(mfScaleFactor);
    mvScaleFactors = mpORBextractorLeft->GetScaleFactors();
    mvInvScaleFactors = mpORBextractorLeft->GetInverseScaleFactors();
    mvLevelSigma2 = mpORBextractorLeft->GetScaleSigmaSquares();
    mvInvLevelSigma2 = mpORBextractorLeft->GetInverseScaleSigmaSquares();

    // ORB extraction: the reference drives the two extractors from two host threads (:161-164).  Each drop-in extractor
    // owns a CUDA stream, so ONE thread keeps both views in flight - enqueue left, enqueue right, collect left, collect
    // right - and the two thread launches per frame go away.  (Frame::ExtractORB is kept for callers that want the
    // reference's own form.)
    mpORBextractorLeft->Begin(imLeft);
    mpORBextractorRight->Begin(imRight);
    mpORBextractorLeft->End(mvKeys, mDescriptors);
    mpORBextractorRight->End(mvKeysRight, mDescriptorsRight);
    mnExtractTokenLeft = pslot_orb_extract_token(mpORBextractorLeft->Handle());
    mnExtractTokenRight = pslot_orb_extract_token(mpORBextractorRight->Handle());

    N = (int)mvKeys.size();
    if (mvKeys.empty()) return;

    UndistortKeyPoints();

    // The reference calls ComputeStereoMatches() here (:173) while `mb` is still unassigned (it is set at
    // :197), i.e. maxD = mbf/mb reads an indeterminate value.  The intended value mb = mbf/fx is set first
    // (DESIGN.md parity rule 2); fx comes from K exactly as :187 does.
    if (mbInitialComputations) {
        ComputeImageBounds(imLeft);
        mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (mnMaxX - mnMinX);
        mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (mnMaxY - mnMinY);
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

    // object modes: detections / masks / AssignFeatures / ExtractObjORB + ComputeObjStereoMatches are PointSLOT's own
    // code (reference :243-305, 369-404, 469-492, 547-600, 716-740); they run here, where the reference runs them
    if (objectMode && ObjectStageHook && imColor) {
        ObjectStageHook(*this, imLeft, imRight, *imColor, objectMode);
        N = (int)mvKeys.size(); // AssignFeatures may have re-partitioned the keypoints (:308, :407, ...)
    }

    mvpMapPoints = std::vector<MapPoint *>(N, static_cast<MapPoint *>(nullptr)); // (:175)

    AssignFeaturesToGrid(); // (:199)
}

#define PSLOT_FRAME_INIT                                                                                               \
    mpORBvocabulary(voc), mpORBextractorLeft(extractorLeft), mpORBextractorRight(extractorRight), mTimeStamp(timeStamp), \
        mK(K.clone()), mDistCoef(distCoef.clone()), mbf(bf), mb(0), mThDepth(thDepth), N(0)

Frame::Frame(const cv::Mat &imLeft, const cv::Mat &imRight, const double &timeStamp, ORBextractor *extractorLeft,
             ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef, const float &bf,
             const float &thDepth)
    : PSLOT_FRAME_INIT
{
    StereoFrontEnd(imLeft, imRight, K, nullptr, 0);
}

// semantic dynamic SLAM (:203)
Frame::Frame(const cv::Mat &imLeft, const cv::Mat &imRight, cv::Mat &imColor, const double &timeStamp,
             ORBextractor *extractorLeft, ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef,
             const float &bf, const float &thDepth, Detector * /*YOLODetector*/)
    : PSLOT_FRAME_INIT
{
    StereoFrontEnd(imLeft, imRight, K, &imColor, 1);
}

// object tracking + dynamic SLAM (:328)
Frame::Frame(const cv::Mat &imLeft, const cv::Mat &imRight, cv::Mat &imColor, const double &timeStamp,
             ORBextractor *extractorLeft, ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef,
             const float &bf, const float &thDepth, Detector * /*YOLODetector*/, cv::MultiTracker * /*multiTracker*/,
             std::vector<cv::Ptr<cv::Tracker>> /*vTrackers*/)
    : PSLOT_FRAME_INIT
{
    StereoFrontEnd(imLeft, imRight, K, &imColor, 2);
}

// object tracking (:428)
Frame::Frame(const cv::Mat &imLeft, const cv::Mat &imRight, cv::Mat &imColor, const double &timeStamp,
             ORBextractor *extractorLeft, ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef,
             const float &bf, const float &thDepth, cv::MultiTracker * /*multiTracker*/,
             std::vector<cv::Ptr<cv::Tracker>> /*vTrackers*/)
    : PSLOT_FRAME_INIT
{
    StereoFrontEnd(imLeft, imRight, K, &imColor, 2);
}

// autonomous driving mode: YOLO + DeepSORT, dense object features (:516)
Frame::Frame(const cv::Mat &imLeft, const cv::Mat &imRight, cv::Mat &imColor, const double &timeStamp,
             ORBextractor *extractorLeft, ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef,
             const float &bf, const float &thDepth, Detector * /*YOLODetector*/, DS::DeepSort * /*deepSort*/)
    : PSLOT_FRAME_INIT
{
    StereoFrontEnd(imLeft, imRight, K, &imColor, 3);
}

// offline "ultimate test" mode (:660)
Frame::Frame(const cv::Mat &imLeft, const cv::Mat &imRight, cv::Mat &imColor, const double &timeStamp,
             ORBextractor *extractorLeft, ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef,
             const float &bf, const float &thDepth, const int &SLOTMode)
    : PSLOT_FRAME_INIT
{
    StereoFrontEnd(imLeft, imRight, K, &imColor, SLOTMode > 0 ? SLOTMode : 4);
}
#undef PSLOT_FRAME_INIT

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

void Frame::ComputeImageBounds(const cv::Mat &imLeft)
{
    if (!mDistCoef.empty() && mDistCoef.at<float>(0) != 0.0) {
#ifdef PSLOT_USE_OPENCV
        cv::Mat mat(4, 2, CV_32F);
        mat.at<float>(0, 0) = 0.0; mat.at<float>(0, 1) = 0.0;
        mat.at<float>(1, 0) = imLeft.cols; mat.at<float>(1, 1) = 0.0;
        mat.at<float>(2, 0) = 0.0; mat.at<float>(2, 1) = imLeft.rows;
        mat.at<float>(3, 0) = imLeft.cols; mat.at<float>(3, 1) = imLeft.rows;
        mat = mat.reshape(2);
        cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK);
        mat = mat.reshape(1);
        mnMinX = std::min(mat.at<float>(0, 0), mat.at<float>(2, 0));
        mnMaxX = std::max(mat.at<float>(1, 0), mat.at<float>(3, 0));
        mnMinY = std::min(mat.at<float>(0, 1), mat.at<float>(1, 1));
        mnMaxY = std::max(mat.at<float>(2, 1), mat.at<float>(3, 1));
#else
        std::fprintf(stderr, "pslot Frame: distorted image bounds need OpenCV (build with -DPSLOT_USE_OPENCV)\n");
        std::abort();
#endif
    } else {
        mnMinX = 0.0f;
        mnMaxX = imLeft.cols;
        mnMinY = 0.0f;
        mnMaxY = imLeft.rows;
    }
}

void Frame::AssignFeaturesToGrid()
{
    // the device builds mGrid as a CSR table (and keeps it, with mvKeysUn, for GetFeaturesInAreas); the member
    // vectors are then filled from it in the same order the reference's push_back loop produces (:1647-1654)
    const pslot_grid_spec spec = {mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv, FRAME_GRID_COLS,
                                  FRAME_GRID_ROWS};
    std::vector<int32_t> cellStart(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1), indices(std::max(N, 1));
    int rc = pslot_assign_grid(mpORBextractorLeft->Handle(), reinterpret_cast<const pslot_keypoint *>(mvKeysUn.data()),
                               N, &spec, cellStart.data(), indices.data(), nullptr);
    if (rc != PSLOT_OK) {
        std::fprintf(stderr, "pslot Frame: pslot_assign_grid failed: %s\n", pslot_last_error());
        std::abort();
    }
    mnGridToken = pslot_grid_token(mpORBextractorLeft->Handle());
    for (unsigned int i = 0; i < FRAME_GRID_COLS; i++)
        for (unsigned int j = 0; j < FRAME_GRID_ROWS; j++) {
            const int c = i * FRAME_GRID_ROWS + j;
            mGrid[i][j].assign(indices.begin() + cellStart[c], indices.begin() + cellStart[c + 1]);
        }
}

void Frame::EnsureDeviceGrid() const
{
    pslot_orb *h = mpORBextractorLeft->Handle();
    if (mnGridToken != 0 && pslot_grid_token(h) == mnGridToken) return; // the device still holds this frame's set
    const pslot_grid_spec spec = {mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv, FRAME_GRID_COLS,
                                  FRAME_GRID_ROWS};
    if (pslot_assign_grid(h, reinterpret_cast<const pslot_keypoint *>(mvKeysUn.data()), N, &spec, nullptr, nullptr, nullptr) != PSLOT_OK) {
        std::fprintf(stderr, "pslot Frame: pslot_assign_grid failed: %s\n", pslot_last_error());
        std::abort();
    }
    mnGridToken = pslot_grid_token(h);
}

bool Frame::PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY)
{
    posX = round((kp.pt.x - mnMinX) * mfGridElementWidthInv);
    posY = round((kp.pt.y - mnMinY) * mfGridElementHeightInv);
    // Keypoint's coordinates are undistorted, which could cause to go out of the image
    return !(posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS);
}

std::vector<size_t> Frame::GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel,
                                             const int maxLevel) const
{
    // one window: a host walk over mGrid, as in the reference; the batched form below runs on the device
    std::vector<size_t> vIndices;
    vIndices.reserve(N);
    const int nMinCellX = std::max(0, (int)std::floor((x - mnMinX - r) * mfGridElementWidthInv));
    if (nMinCellX >= FRAME_GRID_COLS) return vIndices;
    const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)std::ceil((x - mnMinX + r) * mfGridElementWidthInv));
    if (nMaxCellX < 0) return vIndices;
    const int nMinCellY = std::max(0, (int)std::floor((y - mnMinY - r) * mfGridElementHeightInv));
    if (nMinCellY >= FRAME_GRID_ROWS) return vIndices;
    const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)std::ceil((y - mnMinY + r) * mfGridElementHeightInv));
    if (nMaxCellY < 0) return vIndices;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
            for (size_t idx : mGrid[ix][iy]) {
                const cv::KeyPoint &kpUn = mvKeysUn[idx];
                if (bCheckLevels) {
                    if (kpUn.octave < minLevel) continue;
                    if (maxLevel >= 0 && kpUn.octave > maxLevel) continue;
                }
                if (std::fabs(kpUn.pt.x - x) < r && std::fabs(kpUn.pt.y - y) < r) vIndices.push_back(idx);
            }
    return vIndices;
}

std::vector<std::vector<size_t>> Frame::GetFeaturesInAreas(const std::vector<AreaQuery> &queries) const
{
    static_assert(sizeof(AreaQuery) == sizeof(pslot_area_query), "AreaQuery must mirror pslot_area_query");
    const int nq = (int)queries.size();
    std::vector<std::vector<size_t>> res(nq);
    if (nq == 0 || N == 0) return res;
    EnsureDeviceGrid();
    int cap = 64;
    std::vector<int32_t> counts(nq), idx;
    for (;;) {
        idx.resize((size_t)nq * cap);
        int rc = pslot_features_in_area(mpORBextractorLeft->Handle(),
                                        reinterpret_cast<const pslot_area_query *>(queries.data()), nq, cap,
                                        counts.data(), idx.data());
        if (rc != PSLOT_OK) {
            std::fprintf(stderr, "pslot Frame: pslot_features_in_area failed: %s\n", pslot_last_error());
            std::abort();
        }
        const int need = *std::max_element(counts.begin(), counts.end());
        if (need <= cap) break;
        cap = need; // a window held more features than the buffer: ask again with room for the largest
    }
    for (int q = 0; q < nq; ++q) res[q].assign(idx.begin() + (size_t)q * cap, idx.begin() + (size_t)q * cap + counts[q]);
    return res;
}

void Frame::StereoMatchSets(const std::vector<cv::KeyPoint> &keysL, const cv::Mat &dl, const std::vector<cv::KeyPoint> &keysR,
                            const cv::Mat &dr, std::vector<float> &uRight, std::vector<float> &depth)
{
    const int Nl = (int)keysL.size(), Nr = (int)keysR.size();
    uRight = std::vector<float>(Nl, -1.0f);
    depth = std::vector<float>(Nl, -1.0f);
    if (Nl == 0) return;
    const float minZ = mb;
    const float maxD = mbf / minZ; // (:2172-2174)

    // cv::Mat descriptor rows are contiguous for matrices made by create(); copy if a caller handed a view
    std::vector<unsigned char> bl, br;
    const unsigned char *pl = dl.data, *pr = dr.data;
    if (dl.step != 32) { bl.resize((size_t)Nl * 32); for (int i = 0; i < Nl; ++i) std::memcpy(&bl[(size_t)i * 32], dl.ptr(i), 32); pl = bl.data(); }
    if (Nr && dr.step != 32) { br.resize((size_t)Nr * 32); for (int i = 0; i < Nr; ++i) std::memcpy(&br[(size_t)i * 32], dr.ptr(i), 32); pr = br.data(); }

    int rc = pslot_stereo_match(mpORBextractorLeft->Handle(), mpORBextractorRight->Handle(),
                                reinterpret_cast<const pslot_keypoint *>(keysL.data()), Nl, pl,
                                reinterpret_cast<const pslot_keypoint *>(keysR.data()), Nr, pr, mbf, maxD, uRight.data(),
                                depth.data());
    if (rc != PSLOT_OK) {
        std::fprintf(stderr, "pslot Frame: pslot_stereo_match failed: %s\n", pslot_last_error());
        std::abort();
    }
}

void Frame::ComputeStereoMatches() // (:2142-2316)
{
    // In the constructors this runs right after the two extractions: the feature sets are still on the device, untouched
    // (the tokens taken after End() say so), and need not travel back up.  Any other caller - or a frame whose extractors
    // have moved on - takes the general form, which uploads mvKeys / mDescriptors as they are now.
    const int Nl = (int)mvKeys.size();
    if (Nl > 0 && mnExtractTokenLeft && mnExtractTokenRight) {
        mvuRight = std::vector<float>(Nl, -1.0f);
        mvDepth = std::vector<float>(Nl, -1.0f);
        int n = 0;
        const int rc = pslot_stereo_match_resident(mpORBextractorLeft->Handle(), mpORBextractorRight->Handle(), mnExtractTokenLeft,
                                                   mnExtractTokenRight, mbf, mbf / mb, mvuRight.data(), mvDepth.data(), Nl, &n);
        if (rc == PSLOT_OK && n == Nl) return;
        if (rc != PSLOT_ERR_STATE && rc != PSLOT_OK) {
            std::fprintf(stderr, "pslot Frame: pslot_stereo_match_resident failed: %s\n", pslot_last_error());
            std::abort();
        }
    }
    StereoMatchSets(mvKeys, mDescriptors, mvKeysRight, mDescriptorsRight, mvuRight, mvDepth);
}

void Frame::ComputeObjStereoMatches() // (:2318-2503): the same routine on the object feature set
{
    if (mvTempObjKeys.size() <= 0) return; // (:2320)
    StereoMatchSets(mvTempObjKeys, mTempObjPointsDescriptors, mvTempObjKeysRight, mTempObjPointsDescriptorsRight,
                    mvuTempObjKeysRight, mvTempObjDepth);
}

void Frame::ComputeBoW() // (:2040-2047)
{
    if (mBowVec.empty()) {
        // Converter::toDescriptorVector(mDescriptors) + transform(vCurrentDesc, mBowVec, mFeatVec, 4): the rows of
        // mDescriptors are the features, in order
        std::vector<unsigned char> flat(32 * (size_t)mDescriptors.rows);
        for (int i = 0; i < mDescriptors.rows; ++i) std::memcpy(&flat[32 * (size_t)i], mDescriptors.ptr(i), 32);
        mpORBvocabulary->transform(flat.data(), mDescriptors.rows, nullptr, mBowVec, mFeatVec, 4);
    }
}

void OpencvORBDetector(cv::Mat im, cv::Mat ObjMask, std::vector<cv::KeyPoint> &kp, cv::Mat &descriptor) // (:2623-2627)
{
    struct Ctx {
        pslot_cvorb *h = nullptr;
        ~Ctx() { pslot_cvorb_destroy(h); }
    };
    static thread_local Ctx ctx; // ExtractObjORB calls this from two threads at once: one device context each
    if (!ctx.h && pslot_cvorb_create(&ctx.h, 1000, 1.2f, 8, 19, 20, 0) != PSLOT_OK) {
        std::fprintf(stderr, "pslot OpencvORBDetector: %s\n", pslot_last_error());
        std::abort();
    }
    int cap = 1000 + 64 * 8 + 256, n = 0;
    for (;;) {
        kp.resize((size_t)cap);
        descriptor.create(cap, 32, CV_8U);
        const int rc = pslot_cvorb_detect_and_compute(ctx.h, im.data, im.cols, im.rows, (ptrdiff_t)im.step,
                                                      ObjMask.empty() ? nullptr : ObjMask.data, (ptrdiff_t)ObjMask.step,
                                                      reinterpret_cast<pslot_keypoint *>(kp.data()), descriptor.data, cap, &n);
        if (rc == PSLOT_ERR_CAPACITY && n > cap) { cap = n; continue; }
        if (rc != PSLOT_OK) {
            std::fprintf(stderr, "pslot OpencvORBDetector: %s\n", pslot_last_error());
            std::abort();
        }
        break;
    }
    kp.resize((size_t)n);
    cv::Mat d(n, 32, CV_8U);
    for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), descriptor.ptr(i), 32);
    descriptor = d;
}

void Frame::ExtractObjORB(cv::Mat left, cv::Mat right, cv::Mat LeftObjMask, cv::Mat RightObjMask) // (:2646-2663)
{
    // The reference runs OpencvORBDetector for the two views on two threads (:2648-2651); here both views go through ONE
    // batched call (one set of launches, one result round trip).  Views that do not share a size or pitch fall back to
    // the reference's own two-thread form over the single-image entry point.
    std::vector<cv::KeyPoint> keypoints_1, keypoints_2;
    cv::Mat descriptors_1, descriptors_2;
    const bool masks = !LeftObjMask.empty() && !RightObjMask.empty();
    const bool same = left.cols == right.cols && left.rows == right.rows && left.step == right.step &&
                      LeftObjMask.empty() == RightObjMask.empty() && (!masks || LeftObjMask.step == RightObjMask.step);
    if (same) {
        struct Ctx {
            pslot_cvorb *h = nullptr;
            ~Ctx() { pslot_cvorb_destroy(h); }
        };
        static thread_local Ctx ctx;
        if (!ctx.h && pslot_cvorb_create(&ctx.h, 1000, 1.2f, 8, 19, 20, 0) != PSLOT_OK) {
            std::fprintf(stderr, "pslot ExtractObjORB: %s\n", pslot_last_error());
            std::abort();
        }
        const unsigned char *imgs[2] = {left.data, right.data};
        const unsigned char *msk[2] = {masks ? LeftObjMask.data : nullptr, masks ? RightObjMask.data : nullptr};
        int cap = 1000 + 64 * 8 + 256, n[2] = {0, 0};
        std::vector<pslot_keypoint> k;
        std::vector<unsigned char> d;
        for (;;) {
            k.resize(2 * (size_t)cap);
            d.resize(2 * (size_t)cap * 32);
            const int rc = pslot_cvorb_detect_and_compute_batch(ctx.h, 2, imgs, left.cols, left.rows, (ptrdiff_t)left.step,
                                                                masks ? msk : nullptr, masks ? (ptrdiff_t)LeftObjMask.step : 0,
                                                                k.data(), d.data(), cap, n);
            if (rc == PSLOT_ERR_CAPACITY && std::max(n[0], n[1]) > cap) { cap = std::max(n[0], n[1]); continue; }
            if (rc != PSLOT_OK) {
                std::fprintf(stderr, "pslot ExtractObjORB: %s\n", pslot_last_error());
                std::abort();
            }
            break;
        }
        static_assert(sizeof(cv::KeyPoint) == sizeof(pslot_keypoint), "cv::KeyPoint layout");
        auto take = [&](int s, std::vector<cv::KeyPoint> &kp, cv::Mat &desc) {
            kp.resize((size_t)n[s]);
            if (n[s]) std::memcpy(static_cast<void *>(kp.data()), k.data() + (size_t)s * cap, sizeof(pslot_keypoint) * (size_t)n[s]);
            desc.create(n[s], 32, CV_8U);
            for (int i = 0; i < n[s]; ++i) std::memcpy(desc.ptr(i), d.data() + ((size_t)s * cap + i) * 32, 32);
        };
        take(0, keypoints_1, descriptors_1);
        take(1, keypoints_2, descriptors_2);
    } else {
        std::thread threadOpencvORBLeft(OpencvORBDetector, left, LeftObjMask, std::ref(keypoints_1), std::ref(descriptors_1));
        std::thread threadOpencvORBRight(OpencvORBDetector, right, RightObjMask, std::ref(keypoints_2), std::ref(descriptors_2));
        threadOpencvORBLeft.join();
        threadOpencvORBRight.join();
    }
    mvTempObjKeys = keypoints_1;
    mvTempObjKeysRight = keypoints_2;
    mTempObjPointsDescriptors = descriptors_1.clone();
    mTempObjPointsDescriptorsRight = descriptors_2.clone();
}

void Frame::AssignDetObjFeasToGrid(const size_t &objOrder, const size_t &objFeaNum) // (:1863-1888)
{
    auto &grid = mvObjKeysGrid[objOrder];
    grid.assign(FRAME_GRID_COLS, std::vector<std::vector<std::size_t>>(FRAME_GRID_ROWS));
    for (size_t i = 0; i < objFeaNum; i++) {
        const cv::KeyPoint &kp = mvObjKeysUn[objOrder][i];
        int nGridPosX, nGridPosY;
        if (PosInGrid(kp, nGridPosX, nGridPosY)) grid[nGridPosX][nGridPosY].push_back(i);
    }
}

void Frame::AssignFeatures(std::vector<DetectionObject *> &vDetectionObjects) // (:759-977)
{
    const int nObj = (int)vDetectionObjects.size();
    if (nObj > 256 || mMaskImg.empty()) {
        std::fprintf(stderr, "pslot Frame::AssignFeatures: %s\n", mMaskImg.empty() ? "mMaskImg is empty" : "more than 256 detections");
        std::abort();
    }
    std::vector<int32_t> ids((size_t)std::max(nObj, 1), -1);
    for (int j = 0; j < nObj; ++j)
        if (vDetectionObjects[j]) ids[j] = vDetectionObjects[j]->mnObjectID; // a null detection matches no label (:839)
    pslot_orb *ctx = mpORBextractorLeft->Handle();
    auto partition = [&](const std::vector<cv::KeyPoint> &keys, bool first, const std::vector<float> *depth, std::vector<int32_t> &order,
                         std::vector<int32_t> &start, std::vector<int32_t> &effective) {
        const int n = (int)keys.size();
        std::vector<int32_t> cls((size_t)std::max(n, 1));
        order.assign((size_t)std::max(n, 1), 0);
        start.assign((size_t)nObj + 2, 0);
        effective.assign((size_t)std::max(nObj, 1), 0);
        const int rc = pslot_assign_features(ctx, reinterpret_cast<const pslot_keypoint *>(keys.data()), n,
                                             first ? mMaskImg.data : nullptr, mMaskImg.cols, mMaskImg.rows, (ptrdiff_t)mMaskImg.step,
                                             ids.data(), nObj, depth ? depth->data() : nullptr, 2 * mThDepth, cls.data(), order.data(),
                                             start.data(), depth ? effective.data() : nullptr);
        if (rc != PSLOT_OK) {
            std::fprintf(stderr, "pslot Frame::AssignFeatures: %s\n", pslot_last_error());
            std::abort();
        }
    };
    std::vector<int32_t> order, start, eff, oorder, ostart, oeff;
    partition(mvKeys, true, nullptr, order, start, eff);                // the frame's ORB features (:813-868)
    partition(mvTempObjKeys, false, &mvTempObjDepth, oorder, ostart, oeff); // the dense object features (:870-903), same mask

    auto gather_keys = [](const std::vector<cv::KeyPoint> &src, const int32_t *idx, int n) {
        std::vector<cv::KeyPoint> out((size_t)n);
        for (int r = 0; r < n; ++r) out[r] = src[idx[r]];
        return out;
    };
    auto gather_f = [](const std::vector<float> &src, const int32_t *idx, int n) {
        std::vector<float> out((size_t)n);
        for (int r = 0; r < n; ++r) out[r] = src[idx[r]];
        return out;
    };
    auto gather_rows = [](const cv::Mat &src, const int32_t *idx, int n) {
        cv::Mat out;
        if (n == 0) return out; // push_back never ran: an empty matrix, as in the reference
        out.create(n, 32, CV_8U);
        for (int r = 0; r < n; ++r) std::memcpy(out.ptr(r), src.ptr(idx[r]), 32);
        return out;
    };
    mvnEffectiveObjPointDepthNums.clear();
    for (int j = 0; j < nObj; ++j) { // (:914-948) every non-null detection, in order
        if (!vDetectionObjects[j]) continue;
        const int32_t *oi = order.data() + start[j + 1], *di = oorder.data() + ostart[j + 1];
        const int no = start[j + 2] - start[j + 1], nd = ostart[j + 2] - ostart[j + 1];
        mvDetectionObjects.push_back(vDetectionObjects[j]);
        mvObjKeys.push_back(gather_keys(mvTempObjKeys, di, nd));
        mvObjPointsDescriptors.push_back(gather_rows(mTempObjPointsDescriptors, di, nd));
        mvObjKeysUn.push_back(gather_keys(mvTempObjKeysUn, di, nd));
        mvObjPointDepth.push_back(gather_f(mvTempObjDepth, di, nd));
        mvuObjKeysRight.push_back(gather_f(mvuTempObjKeysRight, di, nd));
        mvOriKeys.push_back(gather_keys(mvKeys, oi, no));
        mOriDescriptors.push_back(gather_rows(mDescriptors, oi, no));
        mvOriKeysUn.push_back(gather_keys(mvKeysUn, oi, no));
        mvOriDepth.push_back(gather_f(mvDepth, oi, no));
        mvuOriRight.push_back(gather_f(mvuRight, oi, no));
        mvnEffectiveObjPointDepthNums.push_back((size_t)oeff[j]);
        vDetectionObjects[j]->InitInFrameMapObjectPointsOrder((size_t)nd);
    }
    // background features (:950-955)
    const int nb = start[1];
    std::vector<cv::KeyPoint> keys = gather_keys(mvKeys, order.data(), nb), keysUn = gather_keys(mvKeysUn, order.data(), nb);
    cv::Mat desc = gather_rows(mDescriptors, order.data(), nb);
    std::vector<float> depth = gather_f(mvDepth, order.data(), nb), uright = gather_f(mvuRight, order.data(), nb);
    mvKeys.swap(keys); mvKeysUn.swap(keysUn); mDescriptors = desc; mvDepth.swap(depth); mvuRight.swap(uright);
    mnDetObj = mvDetectionObjects.size();
    mvObjKeysGrid.resize(mnDetObj);
    mvbObjKeysOutlier.resize(mnDetObj);
    mvbObjKeysMatchedFlag.resize(mnDetObj);
    for (size_t i = 0; i < mnDetObj; i++) { // (:967-975)
        const size_t nFeaNum = mvObjKeysUn[i].size();
        mvbObjKeysOutlier[i] = std::vector<bool>(nFeaNum, false);
        mvbObjKeysMatchedFlag[i] = std::vector<bool>(nFeaNum, false);
        AssignDetObjFeasToGrid(i, nFeaNum);
    }
}

void Frame::UndistortObjKeyPoints() // (:2080-2110)
{
    if (mDistCoef.empty() || mDistCoef.at<float>(0) == 0.0) {
        mvTempObjKeysUn = mvTempObjKeys;
        return;
    }
    std::fprintf(stderr, "pslot Frame: distorted object keypoints are not supported (all shipped cameras are rectified)\n");
    std::abort();
}

} // namespace ORB_SLAM2
