// ORBmatcher.cc — ORBmatcher::DescriptorDistance (reference src/ORBmatcher.cc:2704-2720) and the local-map
// ORBmatcher::SearchByProjection (:68-155) over the C ABI.
#include "ORBmatcher.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/pslot_orb.h"

namespace ORB_SLAM2
{

const int ORBmatcher::TH_HIGH = 100; // reference ORBmatcher.cc:58-62
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

int ORBmatcher::SearchByProjection(Frame &F, const std::vector<MapPoint *> &vpMapPoints, const float th)
{
    // Flatten what the reference loop reads from the MapPoint and Frame objects (:75-120) ...
    const int nMP = (int)vpMapPoints.size(), N = F.N;
    if (nMP == 0 || N == 0) return 0;
    std::vector<pslot_map_point> mps(nMP);
    std::vector<unsigned char> mpDesc((size_t)nMP * 32), occupied(N, 0);
    for (int i = 0; i < nMP; ++i) {
        MapPoint *pMP = vpMapPoints[i];
        const bool searched = pMP->mbTrackInView && !pMP->isBad();
        mps[i].proj_x = pMP->mTrackProjX; mps[i].proj_y = pMP->mTrackProjY; mps[i].proj_xr = pMP->mTrackProjXR;
        mps[i].view_cos = pMP->mTrackViewCos;
        mps[i].level = pMP->mnTrackScaleLevel;
        mps[i].flags = (searched ? 1 : 0) | (pMP->Observations() > 0 ? 2 : 0);
        if (searched) {
            const cv::Mat d = pMP->GetDescriptor();
            std::memcpy(&mpDesc[(size_t)i * 32], d.ptr<unsigned char>(), 32);
        }
    }
    if ((int)F.mvpMapPoints.size() < N) F.mvpMapPoints.resize(N, static_cast<MapPoint *>(nullptr));
    for (int idx = 0; idx < N; ++idx)
        occupied[idx] = F.mvpMapPoints[idx] && F.mvpMapPoints[idx]->Observations() > 0; // (:111-113)
    std::vector<unsigned char> fdesc;
    const unsigned char *pd = F.mDescriptors.data;
    if (F.mDescriptors.step != 32) {
        fdesc.resize((size_t)N * 32);
        for (int i = 0; i < N; ++i) std::memcpy(&fdesc[(size_t)i * 32], F.mDescriptors.ptr(i), 32);
        pd = fdesc.data();
    }
    // ... run the search on the device, against F's own grid: if another frame has been gridded on the extractor since
    // (or F is a copy), the frame re-uploads its keypoints first (the reference reads F.mGrid / F.mvKeysUn directly) ...
    F.EnsureDeviceGrid();
    std::vector<int32_t> match(nMP);
    int nmatches = 0;
    int rc = pslot_search_by_projection(F.mpORBextractorLeft->Handle(), mps.data(), mpDesc.data(), nMP, pd,
                                        F.mvuRight.data(), occupied.data(), N, F.mvScaleFactors.data(),
                                        (int)F.mvScaleFactors.size(), th, mfNNratio, match.data(), &nmatches);
    if (rc != PSLOT_OK) {
        std::fprintf(stderr, "pslot ORBmatcher: pslot_search_by_projection failed: %s\n", pslot_last_error());
        std::abort();
    }
    // ... and apply F.mvpMapPoints[bestIdx] = pMP in map-point order (:148)
    for (int i = 0; i < nMP; ++i)
        if (match[i] >= 0) F.mvpMapPoints[match[i]] = vpMapPoints[i];
    return nmatches;
}

int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
{
    return pslot_hamming256(a.ptr<unsigned char>(), b.ptr<unsigned char>());
}

} // namespace ORB_SLAM2
