// ORBmatcher.h — the part of PointSLOT's include/ORBmatcher.h that lies on the hot path:
// `static int DescriptorDistance(const cv::Mat&, const cv::Mat&)` (reference ORBmatcher.h:49,
// src/ORBmatcher.cc:2704-2720) and the thresholds TH_LOW / TH_HIGH (ORBmatcher.cc:58-62).
// The remaining members of the reference class (SearchByProjection, SearchByBoW, Fuse, ...) are outside
// the scope of this library and stay in PointSLOT's own ORBmatcher.cc; INTEGRATION.md shows the 3-line
// patch that makes the reference's definition forward to pslot_hamming256.
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include "compat/opencv_min.h"

namespace ORB_SLAM2
{

class ORBmatcher
{
public:
    // Computes the Hamming distance between two ORB descriptors
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);

    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
};

} // namespace ORB_SLAM2

#endif
