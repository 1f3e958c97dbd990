// frame_stub.h - the smallest Frame / MapPoint / ORBmatcher class shells under which the REFERENCE's own member-function
// bodies (lifted verbatim at build time by extract_fn.py) compile: each class declares exactly the members those bodies
// read or write, with the reference's names and types (include/Frame.h:225-316, include/MapPoint.h:110-120,
// include/ORBmatcher.h:45-130).  No algorithm lives here.  Test infrastructure only.
#pragma once
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <iostream>
#include <set>
#include <sys/types.h> // u_int8_t (Frame::AssignFeatures reads the mask as at<u_int8_t>)
#include <vector>

#include "ORBextractor.h" // the reference's own header (-I <reference>/include): mvImagePyramid etc.

#define FRAME_GRID_ROWS 48 // include/Frame.h:40-41
#define FRAME_GRID_COLS 64

namespace ORB_SLAM2 {
using namespace std;

class MapPoint {
public:
    // tracking fields written by Frame::isInFrustum (include/MapPoint.h:110-116)
    float mTrackProjX, mTrackProjY, mTrackProjXR;
    bool mbTrackInView;
    int mnTrackScaleLevel;
    float mTrackViewCos;
    // state behind the accessors
    bool bad_;
    int nobs_;
    cv::Mat desc_, pos_;
    MapPoint() : mTrackProjX(0), mTrackProjY(0), mTrackProjXR(0), mbTrackInView(false), mnTrackScaleLevel(0),
                 mTrackViewCos(0), bad_(false), nobs_(0), mind_(0), maxd_(0), level_(0) {}
    // scale prediction state of the relocalisation search (include/MapPoint.h: GetMin/MaxDistanceInvariance, PredictScale)
    float mind_, maxd_;
    int level_;
    float GetMinDistanceInvariance() { return mind_; }
    float GetMaxDistanceInvariance() { return maxd_; }
    template <typename F> int PredictScale(const float &, F *) { return level_; }
    bool isBad() { return bad_; }
    int Observations() { return nobs_; }
    cv::Mat GetDescriptor() { return desc_.clone(); }
    cv::Mat GetWorldPos() { return pos_.clone(); }
    cv::Mat normal_;
    cv::Mat GetNormal() { return normal_.clone(); }
};

typedef MapPoint MapObjectPoint; // include/MapObjectPoint.h: the members the object overload reads are MapPoint's
struct DetectionObject { // include/DetectionObject.h:33, 40, 84: what the lifted functions read
    cv::Rect mrectBBox;
    int mnObjectID;
    size_t init_order_n_; // argument of the last InitInFrameMapObjectPointsOrder call
    DetectionObject() : mrectBBox(0, 0, 0, 0), mnObjectID(0), init_order_n_(0) {}
    void InitInFrameMapObjectPointsOrder(const size_t &nNum) { init_order_n_ = nNum; }
};
class MapObject;
extern int EnSLOTMode; // include/Parameters.h:75

class Frame {
public:
    // per detected object (include/Frame.h:151-163, 206, 222): the object overload of SearchByProjection reads these
    std::vector<DetectionObject *> mvDetectionObjects;
    std::vector<std::vector<cv::KeyPoint>> mvObjKeysUn;
    std::vector<std::vector<MapObjectPoint *>> mvpMapObjectPoints;
    vector<cv::Mat> mvObjPointsDescriptors;
    std::vector<std::vector<float>> mvuObjKeysRight;
    vector<vector<vector<std::vector<std::size_t>>>> mvObjKeysGrid;
    vector<size_t> GetObjectFeaturesInArea(const int &nDetObjOrder, const size_t &nDetObjFeaNum, const float &x, const float &y, const float &r, const int minLevel = -1, const int maxLevel = -1) const;
    bool isInBBox(const size_t &nOrder, const float &u, const float &v);
    // Frame::AssignFeatures and what it fills (include/Frame.h:153-190, 201, 209, 252)
    void AssignFeatures(vector<DetectionObject *> &vDetectionObjects);
    void AssignDetObjFeasToGrid(const size_t &objOrder, const size_t &objFeaNum);
    cv::Mat mMaskImg;
    float mThDepth;
    size_t mnDetObj;
    std::vector<std::vector<cv::KeyPoint>> mvObjKeys, mvOriKeys, mvOriKeysUn;
    std::vector<std::vector<float>> mvObjPointDepth, mvuOriRight, mvOriDepth;
    vector<cv::Mat> mOriDescriptors;
    std::vector<MapObject *> mvMapObjects;
    std::vector<std::vector<bool>> mvbObjKeysOutlier, mvbObjKeysMatchedFlag;
    vector<size_t> mvnNewConstructedObjOrders, mvTotalTrackedObjOrders;
    vector<pair<size_t, size_t>> mvInLastFrameTrackedObjOrders;
    void ComputeStereoMatches();
    void ComputeObjStereoMatches();
    void AssignFeaturesToGrid();
    vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1, const int maxLevel = -1) const;
    bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY);

    ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
    static float fx, fy, cx, cy;
    float mbf, mb;
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<cv::KeyPoint> mvTempObjKeys, mvTempObjKeysUn, mvTempObjKeysRight; // include/Frame.h:167-173
    vector<float> mvTempObjDepth, mvuTempObjKeysRight;
    cv::Mat mTempObjPointsDescriptors, mTempObjPointsDescriptorsRight;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    cv::Mat mTcw;
    vector<float> mvScaleFactors, mvInvScaleFactors;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    Frame() : mThDepth(0), mnDetObj(0), mpORBextractorLeft(nullptr), mpORBextractorRight(nullptr), mbf(0), mb(0), N(0) {}
};

class KeyFrame { // what the two KeyFrame overloads of ORBmatcher::SearchByProjection read (include/KeyFrame.h)
public:
    std::vector<MapPoint *> GetMapPointMatches() { return mvp_; }
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<MapPoint *> mvp_;
    // the Sim3 overload (ORBmatcher.cc:413-527): calibration, image bounds, scale factors, the grid and its lookup -
    // KeyFrame::GetFeaturesInArea / IsInImage are the reference's own bodies (src/KeyFrame.cc:665-709), lifted like the rest
    float fx, fy, cx, cy;
    int N, mnGridCols, mnGridRows, mnMinX, mnMinY, mnMaxX, mnMaxY;
    float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<float> mvScaleFactors;
    cv::Mat mDescriptors;
    std::vector<std::vector<std::vector<size_t>>> mGrid;
    vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r) const;
    bool IsInImage(const float &x, const float &y) const;
};

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
    int SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th = 3);
    int SearchByProjection(Frame &F, const size_t &nOrder, const std::vector<MapObjectPoint *> &vpMapPoints, const float th);
    int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono);
    int SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const std::set<MapPoint *> &sAlreadyFound, const float th, const int ORBdist);
    int SearchByProjection(KeyFrame *pKF, cv::Mat Scw, const std::vector<MapPoint *> &vpPoints, std::vector<MapPoint *> &vpMatched, int th);
    static const int TH_LOW, TH_HIGH, TH_HIGH_FORDYNAMIC, RADIUS_FORDYNAMIC, HISTO_LENGTH;
    float RadiusByViewingCos(const float &viewCos);
    void ComputeThreeMaxima(std::vector<int> *histo, const int L, int &ind1, int &ind2, int &ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};

} // namespace ORB_SLAM2
