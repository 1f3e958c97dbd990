// Frame.h — the stereo-construction path of PointSLOT's Frame (reference include/Frame.h:65-90, 94, 134, 213-214,
// 232-273, 303-325; src/Frame.cc:144-200, 203-760, 1636-1664, 1808-1861, 2027-2038, 2049-2110, 2112-2140, 2142-2316,
// 2318-2503) as a drop-in: same constructor signatures (all six stereo constructors), same member names, same outputs
// (mvKeys, mvKeysRight, mvKeysUn, mDescriptors, mDescriptorsRight, mvuRight, mvDepth, mGrid, N, mbf, mb, and the
// object-feature set mvTempObj* of ComputeObjStereoMatches).  Everything else in the reference's Frame (detections,
// masks, AssignFeatures, BoW, pose) consumes these members and is out of scope: the five object-mode constructors run
// the hot path exactly like the plain one and then hand over to PointSLOT's own object stage through ObjectStageHook;
// in a PointSLOT tree only the member functions below are replaced and the constructors stay as they are
// (INTEGRATION.md §2).
#ifndef FRAME_H
#define FRAME_H

#include <cstddef>
#include <vector>

#include "DetectionObject.h"
#include "ORBextractor.h"
#include "ORBVocabulary.h"
#include "compat/opencv_min.h"

namespace ORB_SLAM2
{
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

class MapPoint;      // only pointers are stored
class Detector;      // YOLO wrapper (reference include/YOLOdetector.h); opaque, only passed through
} // namespace ORB_SLAM2
namespace DS { class DeepSort; } // reference deepsort/include/deepsort.h; opaque
#ifndef PSLOT_USE_OPENCV
namespace cv { class MultiTracker; class Tracker; template <typename T> struct Ptr { T *p = nullptr; }; } // opencv_contrib tracking; opaque
#endif
namespace ORB_SLAM2
{

class Frame
{
public:
    Frame();

    // Constructor for stereo cameras (reference Frame.h:65 / Frame.cc:144-200).
    Frame(const cv::Mat &imLeft, const cv::Mat &imRight, const double &timeStamp, ORBextractor *extractorLeft,
          ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef, const float &bf,
          const float &thDepth);

    // The five object-mode stereo constructors (reference Frame.h:68-90; Frame.cc:203, 328, 428, 516, 660).  Their hot
    // path is the plain constructor's - scale info, ExtractORB on two threads, UndistortKeyPoints,
    // ComputeStereoMatches, statics, mb, AssignFeaturesToGrid - and runs here identically; the object stage in between
    // (detections, masks, AssignFeatures, ExtractObjORB: out of scope) is PointSLOT's own code, reached through
    // ObjectStageHook when one is installed.
    Frame(const cv::Mat &imLeft, const cv::Mat &imRight, cv::Mat &imColor, const double &timeStamp,
          ORBextractor *extractorLeft, ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef,
          const float &bf, const float &thDepth, Detector *YOLODetector);
    Frame(const cv::Mat &imLeft, const cv::Mat &imRight, cv::Mat &imColor, const double &timeStamp,
          ORBextractor *extractorLeft, ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef,
          const float &bf, const float &thDepth, Detector *YOLODetector, cv::MultiTracker *multiTracker,
          std::vector<cv::Ptr<cv::Tracker>> vTrackers);
    Frame(const cv::Mat &imLeft, const cv::Mat &imRight, cv::Mat &imColor, const double &timeStamp,
          ORBextractor *extractorLeft, ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef,
          const float &bf, const float &thDepth, cv::MultiTracker *multiTracker, std::vector<cv::Ptr<cv::Tracker>> vTrackers);
    Frame(const cv::Mat &imLeft, const cv::Mat &imRight, cv::Mat &imColor, const double &timeStamp,
          ORBextractor *extractorLeft, ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef,
          const float &bf, const float &thDepth, Detector *YOLODetector, DS::DeepSort *deepSort);
    Frame(const cv::Mat &imLeft, const cv::Mat &imRight, cv::Mat &imColor, const double &timeStamp,
          ORBextractor *extractorLeft, ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef,
          const float &bf, const float &thDepth, const int &SLOTMode);

    // Where PointSLOT's object stage plugs in (between ComputeStereoMatches and AssignFeaturesToGrid, as in the
    // reference constructors): called with the frame and the colour image; may fill mvTempObj* and call
    // ComputeObjStereoMatches().  Null by default.
    typedef void (*ObjectStageFn)(Frame &frame, const cv::Mat &imLeft, const cv::Mat &imRight, cv::Mat &imColor, int mode);
    static ObjectStageFn ObjectStageHook;

    // Extract ORB on the image. 0 for left image and 1 for right image.
    void ExtractORB(int flag, const cv::Mat &im);

    // Compute Bag of Words representation (reference Frame.h:94 / Frame.cc:2040-2047): mBowVec / mFeatVec of the frame's
    // descriptors (the tree walk runs on the device, the two containers are filled in DBoW2's own arithmetic).
    void ComputeBoW();

    // Dense object features (reference Frame.cc:2628-2665): OpenCV's own ORB on both views, restricted to the object
    // masks, on two threads; fills mvTempObjKeys / mvTempObjKeysRight and their descriptor matrices.  The reference
    // builds the two masks from mMaskImg / mMaskImgRight (:2633-2645, PointSLOT's own instance-mask members) and
    // calls this with them.
    void ExtractObjORB(cv::Mat left, cv::Mat right, cv::Mat LeftObjMask, cv::Mat RightObjMask);

    // Feature partition by the instance mask (reference Frame.h:209 / Frame.cc:759-977; runs after the two stereo matchers
    // in the object-mode constructors, :478, :648, :733): the frame's ORB features with mMaskImg == 0 stay in mvKeys /
    // mDescriptors / mvuRight / mvDepth / mvKeysUn, the others go to their detection's mvOri* lists; the dense object
    // features (mvTempObj*) go to their detection's mvObj* lists; null detections are dropped from mvDetectionObjects.
    // The classification and the push_back order of every list come from the device (pslot_assign_features, one call
    // per feature set, the mask uploaded once); the containers are then filled on the host.
    void AssignFeatures(std::vector<DetectionObject *> &vDetectionObjects);
    // mvObjKeysGrid[objOrder] from mvObjKeysUn[objOrder] (reference Frame.cc:1863-1888)
    void AssignDetObjFeasToGrid(const size_t &objOrder, const size_t &objFeaNum);

    // Stereo matching of the OBJECT feature set (reference Frame.cc:2318-2503: the same routine as
    // ComputeStereoMatches on mvTempObjKeys / mvTempObjKeysRight and their descriptors, reading the same two
    // pyramids); results in mvuTempObjKeysRight / mvTempObjDepth.
    void ComputeObjStereoMatches();
    // identity when the images are rectified (reference Frame.cc:2080-2110)
    void UndistortObjKeyPoints();

    // Search a match for each keypoint in the left image to a keypoint in the right image.
    // If there is a match, depth is computed and the right coordinate associated to the left keypoint is stored.
    void ComputeStereoMatches();

    // Compute the cell of a keypoint (return false if outside the grid) (reference Frame.cc:2027-2038)
    bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY);

    // Features of mGrid inside the square window |dx| < r, |dy| < r (reference Frame.cc:1808-1861)
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1,
                                          const int maxLevel = -1) const;
    // Extension: many windows at once on the device, against the grid AssignFeaturesToGrid left there; same
    // results as GetFeaturesInArea per window (the search the projection matchers issue per map point).
    struct AreaQuery { float x, y, r; int minLevel, maxLevel; };
    std::vector<std::vector<size_t>> GetFeaturesInAreas(const std::vector<AreaQuery> &queries) const;

public:
    ORBVocabulary *mpORBvocabulary;
    ORBextractor *mpORBextractorLeft, *mpORBextractorRight;

    double mTimeStamp;
    cv::Mat mK;
    static float fx, fy, cx, cy, invfx, invfy;
    cv::Mat mDistCoef;
    float mbf; // stereo baseline multiplied by fx
    float mb;  // stereo baseline in meters
    float mThDepth;

    int N; // number of keypoints

    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight; // negative value for monocular points
    std::vector<float> mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;

    // Bag of Words Vector structures (reference Frame.h:196-197)
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;

    // object feature set of the current detection pass (reference Frame.h:167-173)
    std::vector<cv::KeyPoint> mvTempObjKeys, mvTempObjKeysRight, mvTempObjKeysUn;
    std::vector<float> mvuTempObjKeysRight, mvTempObjDepth;
    cv::Mat mTempObjPointsDescriptors, mTempObjPointsDescriptorsRight;

    // instance mask of the left view (reference Frame.h:190): 0 background, 255 ignored, object id + 1 elsewhere
    cv::Mat mMaskImg;
    // per detected object, filled by AssignFeatures (reference Frame.h:151-164, 176-180, 206)
    std::vector<DetectionObject *> mvDetectionObjects;
    std::vector<std::vector<cv::KeyPoint>> mvObjKeys, mvObjKeysUn, mvOriKeys, mvOriKeysUn;
    std::vector<cv::Mat> mvObjPointsDescriptors, mOriDescriptors;
    std::vector<std::vector<float>> mvObjPointDepth, mvuObjKeysRight, mvuOriRight, mvOriDepth;
    std::vector<std::vector<bool>> mvbObjKeysOutlier, mvbObjKeysMatchedFlag;
    std::vector<std::vector<std::vector<std::vector<std::size_t>>>> mvObjKeysGrid;
    size_t mnDetObj = 0;
    std::vector<size_t> mvnEffectiveObjPointDepthNums; // vnEffectiveObjPointDepthNums of the reference (a local there, :775)

    // MapPoints associated to keypoints, NULL pointer if no association.
    std::vector<MapPoint *> mvpMapPoints;

    int mnScaleLevels;
    float mfScaleFactor;
    float mfLogScaleFactor;
    std::vector<float> mvScaleFactors;
    std::vector<float> mvInvScaleFactors;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    // Keypoints are assigned to cells in a grid to reduce matching complexity when projecting MapPoints.
    static float mfGridElementWidthInv;
    static float mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];

    // Undistorted Image Bounds (computed once).
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;

    static bool mbInitialComputations;

    // The device keeps ONE gridded keypoint set per extractor (the last AssignFeaturesToGrid).  The reference reads
    // F.mGrid / F.mvKeysUn of whichever frame it is handed, so before a device search the frame checks that the set on
    // the device is still its own (pslot_grid_token) and re-uploads it otherwise - copies of a frame, or an older frame
    // searched after a newer one was built on the same extractor, behave as in the reference.
    void EnsureDeviceGrid() const;
    mutable unsigned long long mnGridToken = 0;
    // identities of the feature sets the two extractions of this frame left on the device (pslot_orb_extract_token): while
    // they are current, ComputeStereoMatches matches them where they are instead of uploading mvKeys / mDescriptors again
    unsigned long long mnExtractTokenLeft = 0, mnExtractTokenRight = 0;

private:
    // the hot path shared by all six constructors
    void StereoFrontEnd(const cv::Mat &imLeft, const cv::Mat &imRight, cv::Mat &K, cv::Mat *imColor, int objectMode);
    void StereoMatchSets(const std::vector<cv::KeyPoint> &keysL, const cv::Mat &descL, const std::vector<cv::KeyPoint> &keysR,
                         const cv::Mat &descR, std::vector<float> &uRight, std::vector<float> &depth);
    // identity when the images are rectified (k1 == 0, all shipped YAMLs); reference Frame.cc:2049-2079
    void UndistortKeyPoints();
    // Computes image bounds for the undistorted image (reference Frame.cc:2112-2140)
    void ComputeImageBounds(const cv::Mat &imLeft);
    // Assign keypoints to the grid for speed up feature matching (reference Frame.cc:1636-1655); on the device
    void AssignFeaturesToGrid();
};

// cv::ORB::create(1000, 1.2, 8, 19)->detectAndCompute(im, ObjMask, kp, descriptor) (reference Frame.cc:2623-2627, same
// name and signature): runs on the device through pslot_cvorb_detect_and_compute, one context per calling thread.
// Keypoints come back ordered by (octave, y, x); cv::ORB's own order is unspecified (std::nth_element) and nothing
// downstream depends on it.
void OpencvORBDetector(cv::Mat im, cv::Mat ObjMask, std::vector<cv::KeyPoint> &kp, cv::Mat &descriptor);

} // namespace ORB_SLAM2

#endif
