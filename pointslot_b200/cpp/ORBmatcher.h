// ORBmatcher.h — the part of PointSLOT's include/ORBmatcher.h that this library covers:
// `static int DescriptorDistance(const cv::Mat&, const cv::Mat&)` (reference ORBmatcher.h:49,
// src/ORBmatcher.cc:2704-2720), the thresholds TH_LOW / TH_HIGH (ORBmatcher.cc:58-62) and the local-map
// `int SearchByProjection(Frame &F, const std::vector<MapPoint*> &vpMapPoints, const float th)`
// (ORBmatcher.cc:68-155).  The remaining members of the reference class (the other
// SearchByProjection overloads, SearchByBoW, Fuse, ...) stay in PointSLOT's own ORBmatcher.cc; INTEGRATION.md shows
// the patches that make the reference's definitions forward to the C ABI.
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <vector>

#include "Frame.h"
#include "MapPoint.h"
#include "compat/opencv_min.h"

namespace ORB_SLAM2
{

class ORBmatcher
{
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);

    // Search matches between Frame keypoints and projected MapPoints. Returns number of matches
    // Used to track the local map (Tracking)
    int SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th = 3);

    // Computes the Hamming distance between two ORB descriptors
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);

    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;

protected:
    float mfNNratio;
    bool mbCheckOrientation;
};

} // namespace ORB_SLAM2

#endif
