// ref_frame_glue.cc - C entry points around the REFERENCE's own Frame::ComputeStereoMatches, AssignFeaturesToGrid,
// GetFeaturesInArea, PosInGrid (src/Frame.cc) and ORBmatcher::SearchByProjection x2, DescriptorDistance
// (src/ORBmatcher.cc).  The function bodies are the reference's text, lifted at build time into $(PIECES)/*.inc by
// extract_fn.py and compiled here inside the class shells of frame_stub.h.  Test infrastructure only.
#include <cstdint>
#include <cstring>
#include <map>

#include "frame_stub.h"

using namespace std;

namespace ORB_SLAM2 {
int EnSLOTMode = 4; // src/Parameters.cc; Frame::AssignFeatures returns early in mode 1 only
float Frame::fx, Frame::fy, Frame::cx, Frame::cy;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
#include "matcher_consts.inc"
#include "frame_fns.inc"
#include "keyframe_fns.inc"
#include "matcher_fns.inc"
} // namespace ORB_SLAM2

using ORB_SLAM2::Frame;
using ORB_SLAM2::KeyFrame;
using ORB_SLAM2::MapPoint;
using ORB_SLAM2::ORBextractor;
using ORB_SLAM2::ORBmatcher;

static void set_keys(vector<cv::KeyPoint> &dst, const void *kps, int n)
{
    static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint layout");
    dst.resize((size_t)n);
    if (n) memcpy(dst.data(), kps, sizeof(cv::KeyPoint) * (size_t)n);
}
static cv::Mat desc_mat(const uint8_t *d, int n)
{
    cv::Mat m(n > 0 ? n : 1, 32, CV_8UC1);
    if (n) memcpy(m.data, d, (size_t)n * 32);
    return m;
}

extern "C" {

// Frame::ComputeStereoMatches (Frame.cc:2142-2316) on the pyramids the two reference extractors hold from their last
// operator().  mb is an explicit argument: the reference reads it before the constructor assigns it (rule C.2).
int reff_stereo(void *extL, void *extR, const void *kL, int nL, const uint8_t *dL, const void *kR, int nR,
                const uint8_t *dR, float mbf, float mb, float *uright, float *depth)
{
    Frame F;
    F.mpORBextractorLeft = static_cast<ORBextractor *>(extL);
    F.mpORBextractorRight = static_cast<ORBextractor *>(extR);
    F.mvScaleFactors = F.mpORBextractorLeft->GetScaleFactors();
    F.mvInvScaleFactors = F.mpORBextractorLeft->GetInverseScaleFactors();
    set_keys(F.mvKeys, kL, nL); set_keys(F.mvKeysRight, kR, nR);
    F.N = nL; F.mbf = mbf; F.mb = mb;
    F.mDescriptors = desc_mat(dL, nL); F.mDescriptorsRight = desc_mat(dR, nR);
    F.ComputeStereoMatches();
    memcpy(uright, F.mvuRight.data(), sizeof(float) * (size_t)nL);
    memcpy(depth, F.mvDepth.data(), sizeof(float) * (size_t)nL);
    return 0;
}

// Frame::ComputeObjStereoMatches (Frame.cc:2318-2503) on an object feature set
int reff_obj_stereo(void *extL, void *extR, const void *kL, int nL, const uint8_t *dL, const void *kR, int nR,
                    const uint8_t *dR, float mbf, float mb, float *uright, float *depth)
{
    Frame F;
    F.mpORBextractorLeft = static_cast<ORBextractor *>(extL);
    F.mpORBextractorRight = static_cast<ORBextractor *>(extR);
    F.mvScaleFactors = F.mpORBextractorLeft->GetScaleFactors();
    F.mvInvScaleFactors = F.mpORBextractorLeft->GetInverseScaleFactors();
    set_keys(F.mvTempObjKeys, kL, nL); set_keys(F.mvTempObjKeysRight, kR, nR);
    F.mvTempObjKeysUn = F.mvTempObjKeys;
    F.mbf = mbf; F.mb = mb;
    F.mTempObjPointsDescriptors = desc_mat(dL, nL); F.mTempObjPointsDescriptorsRight = desc_mat(dR, nR);
    F.ComputeObjStereoMatches();
    if ((int)F.mvuTempObjKeysRight.size() != nL) return -1;
    memcpy(uright, F.mvuTempObjKeysRight.data(), sizeof(float) * (size_t)nL);
    memcpy(depth, F.mvTempObjDepth.data(), sizeof(float) * (size_t)nL);
    return 0;
}

int reff_hamming256(const uint8_t *a, const uint8_t *b)
{
    alignas(4) uint8_t ba[32], bb[32];
    memcpy(ba, a, 32); memcpy(bb, b, 32);
    cv::Mat ma(1, 32, CV_8UC1, ba), mb_(1, 32, CV_8UC1, bb);
    return ORBmatcher::DescriptorDistance(ma, mb_);
}

// ---- a stub Frame that lives across calls: grid + projection searches --------------------------------------------------
void *reff_frame_create(const void *kps_un, int n, const uint8_t *desc, const float *uright, const float *scale_factors,
                        int nlevels, float min_x, float max_x, float min_y, float max_y, float inv_w, float inv_h)
{
    Frame *F = new Frame;
    set_keys(F->mvKeysUn, kps_un, n);
    F->mvKeys = F->mvKeysUn;
    F->N = n;
    F->mDescriptors = desc_mat(desc, n);
    F->mvuRight.assign(uright, uright + n);
    F->mvScaleFactors.assign(scale_factors, scale_factors + nlevels);
    F->mvpMapPoints.assign((size_t)n, nullptr);
    F->mvbOutlier.assign((size_t)n, false);
    Frame::mnMinX = min_x; Frame::mnMaxX = max_x; Frame::mnMinY = min_y; Frame::mnMaxY = max_y;
    Frame::mfGridElementWidthInv = inv_w; Frame::mfGridElementHeightInv = inv_h;
    F->AssignFeaturesToGrid();
    return F;
}
void reff_frame_destroy(void *h) { delete static_cast<Frame *>(h); }

// mGrid flattened to CSR, cell = i * FRAME_GRID_ROWS + j; returns the number of indices
int reff_frame_grid(void *h, int *cell_start, int *indices)
{
    Frame &F = *static_cast<Frame *>(h);
    int k = 0;
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) {
            cell_start[i * FRAME_GRID_ROWS + j] = k;
            for (size_t v : F.mGrid[i][j]) indices[k++] = (int)v;
        }
    cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = k;
    return k;
}

int reff_frame_features_in_area(void *h, float x, float y, float r, int min_level, int max_level, int cap, int *out)
{
    Frame &F = *static_cast<Frame *>(h);
    const vector<size_t> v = F.GetFeaturesInArea(x, y, r, min_level, max_level);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int)v[i];
    return (int)v.size();
}

// the MapPoint fields of oracle/orb_oracle.h:orc_map_point
struct reff_map_point { float proj_x, proj_y, proj_xr, view_cos; int32_t level, flags; };

// ORBmatcher::SearchByProjection(F, vpMapPoints, th) (ORBmatcher.cc:68-155).  slot_of_feature[idx] = index of the map
// point F.mvpMapPoints[idx] holds afterwards, -2 for a pre-existing association (occupied), -1 for none.
int reff_search_by_projection(void *h, const reff_map_point *mps, const uint8_t *mp_desc, int n_mp,
                              const uint8_t *occupied, float th, float nn_ratio, int *slot_of_feature)
{
    Frame &F = *static_cast<Frame *>(h);
    vector<MapPoint> store((size_t)n_mp);
    vector<MapPoint *> v((size_t)n_mp);
    map<MapPoint *, int> index;
    for (int i = 0; i < n_mp; ++i) {
        MapPoint &p = store[(size_t)i];
        p.mTrackProjX = mps[i].proj_x; p.mTrackProjY = mps[i].proj_y; p.mTrackProjXR = mps[i].proj_xr;
        p.mTrackViewCos = mps[i].view_cos; p.mnTrackScaleLevel = mps[i].level;
        p.mbTrackInView = (mps[i].flags & 1) != 0; p.bad_ = false; p.nobs_ = (mps[i].flags & 2) ? 1 : 0;
        p.desc_ = desc_mat(mp_desc + (size_t)i * 32, 1);
        v[(size_t)i] = &p; index[&p] = i;
    }
    MapPoint held; held.nobs_ = 1;
    for (int i = 0; i < F.N; ++i) F.mvpMapPoints[(size_t)i] = occupied && occupied[i] ? &held : nullptr;
    ORBmatcher m(nn_ratio, true);
    const int n = m.SearchByProjection(F, v, th);
    for (int i = 0; i < F.N; ++i) {
        MapPoint *p = F.mvpMapPoints[(size_t)i];
        slot_of_feature[i] = !p ? -1 : p == &held ? -2 : index[p];
        F.mvpMapPoints[(size_t)i] = nullptr;
    }
    return n;
}

// ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1613-1756) with identity rotations:
// CurrentFrame.mTcw = I, LastFrame.mTcw = [I | (0, 0, last_tz)], so x3Dc == x3Dw exactly and tlc.z == last_tz.
int reff_search_by_projection_last(void *h, const void *last_keys_un, int n_last, const float *world_xyz,
                                   const uint8_t *last_desc, const uint8_t *last_flags /* bit0 has point, bit1 outlier, bit2 obs>0 */,
                                   float last_tz, float fx, float fy, float cx, float cy, float mbf, float mb,
                                   const uint8_t *occupied, float th, int mono, int check_orientation, int *slot_of_feature)
{
    Frame &C = *static_cast<Frame *>(h);
    Frame L;
    Frame::fx = fx; Frame::fy = fy; Frame::cx = cx; Frame::cy = cy;
    C.mbf = mbf; C.mb = mb;
    C.mTcw = cv::Mat::eye(4, 4, CV_32F);
    L.mTcw = cv::Mat::eye(4, 4, CV_32F); L.mTcw.at<float>(2, 3) = last_tz;
    set_keys(L.mvKeysUn, last_keys_un, n_last); L.mvKeys = L.mvKeysUn; L.N = n_last;
    vector<MapPoint> store((size_t)n_last);
    map<MapPoint *, int> index;
    L.mvpMapPoints.assign((size_t)n_last, nullptr); L.mvbOutlier.assign((size_t)n_last, false);
    for (int i = 0; i < n_last; ++i) {
        MapPoint &p = store[(size_t)i];
        p.pos_ = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; ++k) p.pos_.at<float>(k) = world_xyz[3 * i + k];
        p.desc_ = desc_mat(last_desc + (size_t)i * 32, 1);
        p.nobs_ = (last_flags[i] & 4) ? 1 : 0;
        index[&p] = i;
        if (last_flags[i] & 1) L.mvpMapPoints[(size_t)i] = &p;
        L.mvbOutlier[(size_t)i] = (last_flags[i] & 2) != 0;
    }
    MapPoint held; held.nobs_ = 1;
    for (int i = 0; i < C.N; ++i) C.mvpMapPoints[(size_t)i] = occupied && occupied[i] ? &held : nullptr;
    ORBmatcher m(0.9f, check_orientation != 0);
    const int n = m.SearchByProjection(C, L, th, mono != 0);
    for (int i = 0; i < C.N; ++i) {
        MapPoint *p = C.mvpMapPoints[(size_t)i];
        slot_of_feature[i] = !p ? -1 : p == &held ? -2 : index[p];
        C.mvpMapPoints[(size_t)i] = nullptr;
    }
    return n;
}

// ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:2529-2654) with
// CurrentFrame.mTcw = I: x3Dc == x3Dw and Ow == 0 exactly, so dist3D = |x3Dw|.  kf_flags: bit 0 the key frame holds a map
// point at i, bit 1 it is bad, bit 2 it is in sAlreadyFound.
int reff_search_by_projection_kf(void *h, const void *kf_keys_un, int n_kf, const float *world_xyz, const uint8_t *kf_desc,
                                 const uint8_t *kf_flags, const int *level, const float *min_dist, const float *max_dist,
                                 float fx, float fy, float cx, float cy, const uint8_t *occupied, float th, int orb_dist,
                                 int check_orientation, int *slot_of_feature)
{
    Frame &C = *static_cast<Frame *>(h);
    Frame::fx = fx; Frame::fy = fy; Frame::cx = cx; Frame::cy = cy;
    C.mTcw = cv::Mat::eye(4, 4, CV_32F);
    KeyFrame K;
    set_keys(K.mvKeysUn, kf_keys_un, n_kf);
    vector<MapPoint> store((size_t)n_kf);
    map<MapPoint *, int> index;
    set<MapPoint *> found;
    K.mvp_.assign((size_t)n_kf, nullptr);
    for (int i = 0; i < n_kf; ++i) {
        MapPoint &p = store[(size_t)i];
        p.pos_ = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; ++k) p.pos_.at<float>(k) = world_xyz[3 * i + k];
        p.desc_ = desc_mat(kf_desc + (size_t)i * 32, 1);
        p.bad_ = (kf_flags[i] & 2) != 0;
        p.level_ = level[i]; p.mind_ = min_dist[i]; p.maxd_ = max_dist[i];
        index[&p] = i;
        if (kf_flags[i] & 1) K.mvp_[(size_t)i] = &p;
        if (kf_flags[i] & 4) found.insert(&p);
    }
    MapPoint held;
    for (int i = 0; i < C.N; ++i) C.mvpMapPoints[(size_t)i] = occupied && occupied[i] ? &held : nullptr;
    ORBmatcher m(0.9f, check_orientation != 0);
    const int n = m.SearchByProjection(C, &K, found, th, orb_dist);
    for (int i = 0; i < C.N; ++i) {
        MapPoint *p = C.mvpMapPoints[(size_t)i];
        slot_of_feature[i] = !p ? -1 : p == &held ? -2 : index[p];
        C.mvpMapPoints[(size_t)i] = nullptr;
    }
    return n;
}

// ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) (ORBmatcher.cc:413-527, loop closing): the key frame is
// built from the stub Frame's keypoints / descriptors / grid (KeyFrame's constructor copies Frame::mGrid, src/KeyFrame.cc).
// pt_flags bit 1: the point is bad.  pre_matched[idx] = index of the point already held by vpMatched[idx], -2 for some other
// point, -1 for none; slot_of_feature[idx] = the same encoding afterwards.
int reff_search_by_projection_sim3(void *h, const float *scw16, const float *world_xyz, const float *normals,
                                   const uint8_t *pt_desc, const uint8_t *pt_flags, const int *level, const float *min_dist,
                                   const float *max_dist, int n_pts, float fx, float fy, float cx, float cy,
                                   const int *pre_matched, int th, int *slot_of_feature)
{
    Frame &F = *static_cast<Frame *>(h);
    KeyFrame K;
    K.mvKeysUn = F.mvKeysUn; K.N = F.N; K.mDescriptors = F.mDescriptors; K.mvScaleFactors = F.mvScaleFactors;
    K.fx = fx; K.fy = fy; K.cx = cx; K.cy = cy;
    K.mnGridCols = FRAME_GRID_COLS; K.mnGridRows = FRAME_GRID_ROWS;
    K.mnMinX = (int)Frame::mnMinX; K.mnMinY = (int)Frame::mnMinY; K.mnMaxX = (int)Frame::mnMaxX; K.mnMaxY = (int)Frame::mnMaxY;
    K.mfGridElementWidthInv = Frame::mfGridElementWidthInv; K.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    K.mGrid.resize(FRAME_GRID_COLS);
    for (int i = 0; i < FRAME_GRID_COLS; ++i) {
        K.mGrid[i].resize(FRAME_GRID_ROWS);
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) K.mGrid[i][j] = F.mGrid[i][j];
    }
    cv::Mat Scw(4, 4, CV_32F);
    for (int i = 0; i < 16; ++i) Scw.at<float>(i / 4, i % 4) = scw16[i];
    vector<MapPoint> store((size_t)n_pts);
    vector<MapPoint *> v((size_t)n_pts);
    map<MapPoint *, int> index;
    for (int i = 0; i < n_pts; ++i) {
        MapPoint &p = store[(size_t)i];
        p.pos_ = cv::Mat(3, 1, CV_32F); p.normal_ = cv::Mat(3, 1, CV_32F);
        for (int k = 0; k < 3; ++k) { p.pos_.at<float>(k) = world_xyz[3 * i + k]; p.normal_.at<float>(k) = normals[3 * i + k]; }
        p.desc_ = desc_mat(pt_desc + (size_t)i * 32, 1);
        p.bad_ = (pt_flags[i] & 2) != 0;
        p.level_ = level[i]; p.mind_ = min_dist[i]; p.maxd_ = max_dist[i];
        v[(size_t)i] = &p; index[&p] = i;
    }
    MapPoint held;
    vector<MapPoint *> matched((size_t)F.N, nullptr);
    for (int i = 0; i < F.N; ++i) matched[(size_t)i] = pre_matched[i] >= 0 ? &store[(size_t)pre_matched[i]] : pre_matched[i] == -2 ? &held : nullptr;
    ORBmatcher m(0.75f, true);
    const int n = m.SearchByProjection(&K, Scw, v, matched, th);
    for (int i = 0; i < F.N; ++i) {
        MapPoint *p = matched[(size_t)i];
        slot_of_feature[i] = !p ? -1 : p == &held ? -2 : index[p];
    }
    return n;
}

// ORBmatcher::SearchByProjection(F, nOrder, vpMapObjectPoints, th) (ORBmatcher.cc:157-246): the stub Frame's keypoints,
// descriptors, uRight and grid become detected object 0's feature set (mvObjKeysGrid[0] = the grid the reference's own
// AssignFeaturesToGrid built); bbox = (x, y, w, h) of the detection.  Encoding of slot_of_feature as above.
int reff_search_by_projection_obj(void *h, const reff_map_point *mps, const uint8_t *mp_desc, int n_mp, const uint8_t *occupied,
                                  const int *bbox, float th, float nn_ratio, int *slot_of_feature)
{
    Frame &F = *static_cast<Frame *>(h);
    ORB_SLAM2::DetectionObject det;
    det.mrectBBox = cv::Rect(bbox[0], bbox[1], bbox[2], bbox[3]);
    F.mvDetectionObjects.assign(1, &det);
    F.mvObjKeysUn.assign(1, F.mvKeysUn);
    F.mvObjPointsDescriptors.assign(1, F.mDescriptors);
    F.mvuObjKeysRight.assign(1, F.mvuRight);
    F.mvObjKeysGrid.assign(1, vector<vector<vector<size_t>>>(FRAME_GRID_COLS, vector<vector<size_t>>(FRAME_GRID_ROWS)));
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) F.mvObjKeysGrid[0][i][j] = F.mGrid[i][j];
    vector<MapPoint> store((size_t)n_mp);
    vector<MapPoint *> v((size_t)n_mp);
    map<MapPoint *, int> index;
    for (int i = 0; i < n_mp; ++i) {
        MapPoint &p = store[(size_t)i];
        p.mTrackProjX = mps[i].proj_x; p.mTrackProjY = mps[i].proj_y; p.mTrackProjXR = mps[i].proj_xr;
        p.mTrackViewCos = mps[i].view_cos; p.mnTrackScaleLevel = mps[i].level;
        p.mbTrackInView = (mps[i].flags & 1) != 0; p.bad_ = false; p.nobs_ = (mps[i].flags & 2) ? 1 : 0;
        p.desc_ = desc_mat(mp_desc + (size_t)i * 32, 1);
        v[(size_t)i] = &p; index[&p] = i;
    }
    MapPoint held; held.nobs_ = 1;
    F.mvpMapObjectPoints.assign(1, vector<MapPoint *>((size_t)F.N, nullptr));
    for (int i = 0; i < F.N; ++i) F.mvpMapObjectPoints[0][(size_t)i] = occupied && occupied[i] ? &held : nullptr;
    ORBmatcher m(nn_ratio, true);
    const int n = m.SearchByProjection(F, (size_t)0, v, th);
    for (int i = 0; i < F.N; ++i) {
        MapPoint *p = F.mvpMapObjectPoints[0][(size_t)i];
        slot_of_feature[i] = !p ? -1 : p == &held ? -2 : index[p];
    }
    F.mvDetectionObjects.clear();
    return n;
}

// Frame::AssignFeatures (Frame.cc:759-977) on a stub frame.  Every feature carries its index in KeyPoint::class_id (the
// function copies whole KeyPoints), so the partitions come back as index lists; that the descriptor rows, mvuRight,
// mvDepth and mvKeysUn travelled with their keypoints is checked here (return -2 otherwise).  Outputs:
//   bg_idx / *bg_n                         indices of the frame's ORB features left in mvKeys (the background)
//   det_of / *n_det                        original index j of every detection kept in mvDetectionObjects
//   ori_idx, ori_start[n_det + 1]          mvOriKeys[k]: the ORB features inside detection k's mask
//   obj_idx, obj_start[n_det + 1]          mvObjKeys[k]: the dense object features of detection k
//   init_n[k]                              argument of InitInFrameMapObjectPointsOrder for detection k
//   grid_cs[k][64 * 48 + 1], grid_idx      mvObjKeysGrid[k] flattened (AssignDetObjFeasToGrid), indices LOCAL to mvObjKeys[k]
int reff_assign_features(const void *kps, int n, const uint8_t *desc, const float *uright, const float *depth,
                         const void *okps, int on, const uint8_t *odesc, const float *ouright, const float *odepth,
                         const uint8_t *mask, int w, int h, const int *object_ids, int n_objects, float th_depth,
                         float min_x, float max_x, float min_y, float max_y, float inv_w, float inv_h,
                         int *bg_idx, int *bg_n, int *det_of, int *n_det, int *ori_idx, int *ori_start, int *obj_idx,
                         int *obj_start, int *init_n, int *grid_cs, int *grid_idx)
{
    Frame F;
    set_keys(F.mvKeys, kps, n);
    for (int i = 0; i < n; ++i) F.mvKeys[i].class_id = i;
    F.mvKeysUn = F.mvKeys;
    F.N = n;
    F.mDescriptors = desc_mat(desc, n);
    F.mvuRight.assign(uright, uright + n);
    F.mvDepth.assign(depth, depth + n);
    set_keys(F.mvTempObjKeys, okps, on);
    for (int i = 0; i < on; ++i) F.mvTempObjKeys[i].class_id = i;
    F.mvTempObjKeysUn = F.mvTempObjKeys;
    F.mTempObjPointsDescriptors = desc_mat(odesc, on);
    F.mvuTempObjKeysRight.assign(ouright, ouright + on);
    F.mvTempObjDepth.assign(odepth, odepth + on);
    F.mMaskImg = cv::Mat(h, w, CV_8UC1, const_cast<uint8_t *>(mask));
    F.mThDepth = th_depth;
    Frame::mnMinX = min_x; Frame::mnMaxX = max_x; Frame::mnMinY = min_y; Frame::mnMaxY = max_y;
    Frame::mfGridElementWidthInv = inv_w; Frame::mfGridElementHeightInv = inv_h;
    vector<ORB_SLAM2::DetectionObject> pool((size_t)n_objects);
    vector<ORB_SLAM2::DetectionObject *> dets((size_t)n_objects, nullptr);
    for (int j = 0; j < n_objects; ++j)
        if (object_ids[j] >= 0) { pool[j].mnObjectID = object_ids[j]; dets[j] = &pool[j]; }
    F.AssignFeatures(dets);

    auto same_row = [](const cv::Mat &m, int r, const uint8_t *d, int id) { return memcmp(m.ptr(r), d + (size_t)id * 32, 32) == 0; };
    *bg_n = (int)F.mvKeys.size();
    if (F.mDescriptors.rows != *bg_n && *bg_n) return -2;
    for (int r = 0; r < *bg_n; ++r) {
        const int id = F.mvKeys[r].class_id;
        bg_idx[r] = id;
        if (!same_row(F.mDescriptors, r, desc, id) || F.mvuRight[r] != uright[id] || F.mvDepth[r] != depth[id] || F.mvKeysUn[r].class_id != id) return -2;
    }
    *n_det = (int)F.mvDetectionObjects.size();
    if ((size_t)*n_det != F.mnDetObj || F.mvObjKeysGrid.size() != F.mnDetObj) return -2;
    int ko = 0, kd = 0, kg = 0;
    for (int k = 0; k < *n_det; ++k) {
        det_of[k] = (int)(F.mvDetectionObjects[k] - pool.data());
        init_n[k] = (int)F.mvDetectionObjects[k]->init_order_n_;
        ori_start[k] = ko;
        for (size_t r = 0; r < F.mvOriKeys[k].size(); ++r) {
            const int id = F.mvOriKeys[k][r].class_id;
            ori_idx[ko++] = id;
            if (!same_row(F.mOriDescriptors[k], (int)r, desc, id) || F.mvuOriRight[k][r] != uright[id] || F.mvOriDepth[k][r] != depth[id] ||
                F.mvOriKeysUn[k][r].class_id != id) return -2;
        }
        obj_start[k] = kd;
        for (size_t r = 0; r < F.mvObjKeys[k].size(); ++r) {
            const int id = F.mvObjKeys[k][r].class_id;
            obj_idx[kd++] = id;
            if (!same_row(F.mvObjPointsDescriptors[k], (int)r, odesc, id) || F.mvuObjKeysRight[k][r] != ouright[id] ||
                F.mvObjPointDepth[k][r] != odepth[id] || F.mvObjKeysUn[k][r].class_id != id) return -2;
        }
        if (F.mvpMapObjectPoints[k].size() != F.mvObjKeys[k].size() || F.mvbObjKeysOutlier[k].size() != F.mvObjKeys[k].size()) return -2;
        int *cs = grid_cs + (size_t)k * (FRAME_GRID_COLS * FRAME_GRID_ROWS + 1);
        const int g0 = kg;
        for (int i = 0; i < FRAME_GRID_COLS; ++i)
            for (int j = 0; j < FRAME_GRID_ROWS; ++j) {
                cs[i * FRAME_GRID_ROWS + j] = kg - g0;
                for (size_t v : F.mvObjKeysGrid[k][i][j]) grid_idx[kg++] = (int)v;
            }
        cs[FRAME_GRID_COLS * FRAME_GRID_ROWS] = kg - g0;
    }
    ori_start[*n_det] = ko;
    obj_start[*n_det] = kd;
    return 0;
}

} // extern "C"
