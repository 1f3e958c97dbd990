// MapPoint.h — stand-in for the members of PointSLOT's MapPoint that ORBmatcher::SearchByProjection reads
// (reference include/MapPoint.h:110-120, src/MapPoint.cc isBad / Observations / GetDescriptor), so that the drop-in
// matcher can be compiled and tested here, where PointSLOT itself cannot be built.  In a PointSLOT tree the
// project's own MapPoint.h is used instead; the matcher touches nothing else of the class.
#ifndef MAPPOINT_H
#define MAPPOINT_H

#include "compat/opencv_min.h"

namespace ORB_SLAM2
{

class MapPoint
{
public:
    // Variables used by the tracking
    float mTrackProjX = 0.f;
    float mTrackProjY = 0.f;
    float mTrackProjXR = 0.f;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    float mTrackViewCos = 0.f;

    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }

    // stand-in state (private in the reference, set by the mapping threads)
    bool mbBad = false;
    int nObs = 0;
    cv::Mat mDescriptor;
};

} // namespace ORB_SLAM2

#endif
