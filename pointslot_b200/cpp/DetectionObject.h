// DetectionObject.h — stand-in for the members of PointSLOT's DetectionObject that Frame::AssignFeatures reads
// (reference include/DetectionObject.h:33 mnObjectID, :40 mrectBBox, :84 InitInFrameMapObjectPointsOrder), so that the
// drop-in Frame can be compiled and tested here, where PointSLOT itself cannot be built.  In a PointSLOT tree the project's
// own DetectionObject.h is used instead; the drop-in touches nothing else of the class.
#ifndef DETECTIONOBJECT_H
#define DETECTIONOBJECT_H

#include <cstddef>

namespace ORB_SLAM2
{

class DetectionObject
{
public:
    int mnObjectID = 0;
    void InitInFrameMapObjectPointsOrder(const size_t &nNum) { mnInFrameOrderSize = nNum; }

    // stand-in state (a vector of that size in the reference)
    size_t mnInFrameOrderSize = 0;
};

} // namespace ORB_SLAM2

#endif
