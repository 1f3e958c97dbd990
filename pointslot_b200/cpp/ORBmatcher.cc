// ORBmatcher.cc — ORBmatcher::DescriptorDistance over the C ABI (reference src/ORBmatcher.cc:2704-2720).
#include "ORBmatcher.h"

#include "../../include/pslot_orb.h"

namespace ORB_SLAM2
{

const int ORBmatcher::TH_HIGH = 100; // reference ORBmatcher.cc:58-62
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
{
    return pslot_hamming256(a.ptr<unsigned char>(), b.ptr<unsigned char>());
}

} // namespace ORB_SLAM2
