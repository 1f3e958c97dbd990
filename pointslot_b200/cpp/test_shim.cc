// test_shim.cc — exercises the drop-in classes exactly as Tracking::GrabImageStereo does
// (reference src/Tracking.cc:385-388, 1045): two ORBextractor objects, one Frame per stereo pair.
// usage: test_shim <in.bin> <out.bin> [vocabulary.txt]
//   in : int32 W,H,nfeatures,nlevels,iniTh,minTh; float32 scaleFactor,bf,fx; then W*H left bytes, W*H right bytes
//   out: int32 N, Nr; N keypoints (28 B), N*32 desc, Nr keypoints, Nr*32 desc, N float uRight, N float depth,
//        int32 pyramid level-1 w,h + bytes (checks SyncImagePyramid), int32 DescriptorDistance(desc0, desc1),
//        mGrid as 64*48 x (int32 size, int32 entries...), then 3 windows x (int32 n, entries) from GetFeaturesInArea
//        and the same 3 from GetFeaturesInAreas; then int32 nMP, nMP x (4 float + 2 int32 map point record as
//        pslot_map_point), nMP x 32 descriptor bytes, int32 nmatches of ORBmatcher(0.8).SearchByProjection(frame, mps, 3),
//        N x int32 map-point index held by frame.mvpMapPoints[idx] afterwards (-1 = none);
//        then the object-mode frame built in between on the same extractors from both images rolled down 7 rows
//        (SLOTMode constructor + ObjectStageHook): int32 nObjL, nObjR, nObjL float mvuTempObjKeysRight, nObjL float
//        mvTempObjDepth, int32 N of that frame; then Frame::ExtractObjORB on the first frame's images with two-rectangle
//        object masks: int32 nObjL, nObjR, nObjL keypoints + nObjL*32 desc, nObjR keypoints + nObjR*32 desc; with a third
//        argument (vocabulary text file) Frame::ComputeBoW: int32 nWords, nFv, nWords x (int32 id, double value),
//        nFv x (int32 node, int32 count, count x uint32 feature index)
// The second frame is built BEFORE the area / projection queries on the first one, so those queries only pass when the
// first frame notices that the device grid is no longer its own (Frame::EnsureDeviceGrid).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "Frame.h"
#include "ORBmatcher.h"

using namespace ORB_SLAM2;

// stands in for PointSLOT's object stage: every third left / second right keypoint becomes an "object feature"
static void object_stage(Frame &F, const cv::Mat &, const cv::Mat &, cv::Mat &, int)
{
    std::vector<int> il, ir;
    for (int i = 0; i < (int)F.mvKeys.size(); i += 3) il.push_back(i);
    for (int i = 0; i < (int)F.mvKeysRight.size(); i += 2) ir.push_back(i);
    F.mvTempObjKeys.clear(); F.mvTempObjKeysRight.clear();
    F.mTempObjPointsDescriptors.create((int)il.size(), 32, CV_8U);
    F.mTempObjPointsDescriptorsRight.create((int)ir.size(), 32, CV_8U);
    for (size_t k = 0; k < il.size(); ++k) {
        F.mvTempObjKeys.push_back(F.mvKeys[il[k]]);
        std::memcpy(F.mTempObjPointsDescriptors.ptr((int)k), F.mDescriptors.ptr(il[k]), 32);
    }
    for (size_t k = 0; k < ir.size(); ++k) {
        F.mvTempObjKeysRight.push_back(F.mvKeysRight[ir[k]]);
        std::memcpy(F.mTempObjPointsDescriptorsRight.ptr((int)k), F.mDescriptorsRight.ptr(ir[k]), 32);
    }
    F.UndistortObjKeyPoints();
    F.ComputeObjStereoMatches();
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int hdr[6];
    float fl[3];
    if (std::fread(hdr, 4, 6, f) != 6 || std::fread(fl, 4, 3, f) != 3) return 4;
    const int W = hdr[0], H = hdr[1];
    std::vector<unsigned char> L((size_t)W * H), R((size_t)W * H);
    if (std::fread(L.data(), 1, L.size(), f) != L.size() || std::fread(R.data(), 1, R.size(), f) != R.size()) return 5;
    std::fclose(f);

    ORBextractor *left = new ORBextractor(hdr[2], fl[0], hdr[3], hdr[4], hdr[5]);
    ORBextractor *right = new ORBextractor(hdr[2], fl[0], hdr[3], hdr[4], hdr[5]);
    cv::Mat imL(H, W, CV_8UC1, L.data()), imR(H, W, CV_8UC1, R.data());
    cv::Mat K(3, 3, CV_32F);
    for (int i = 0; i < 9; ++i) K.at<float>(i / 3, i % 3) = 0.f;
    K.at<float>(0, 0) = fl[2]; K.at<float>(1, 1) = fl[2]; K.at<float>(0, 2) = W / 2.f; K.at<float>(1, 2) = H / 2.f; K.at<float>(2, 2) = 1.f;
    cv::Mat dist(4, 1, CV_32F);
    for (int i = 0; i < 4; ++i) dist.at<float>(i) = 0.f;
    const float bf = fl[1], thDepth = 35.f;

    Frame frame;
    for (int rep = 0; rep < 2; ++rep) // twice: contexts are reused frame after frame
        frame = Frame(imL, imR, 0.1 * rep, left, right, nullptr, K, dist, bf, thDepth);

    FILE *o = std::fopen(argv[2], "wb");
    if (!o) return 6;
    const int N = frame.N, Nr = (int)frame.mvKeysRight.size();
    std::fwrite(&N, 4, 1, o); std::fwrite(&Nr, 4, 1, o);
    std::fwrite(frame.mvKeys.data(), sizeof(cv::KeyPoint), N, o);
    for (int i = 0; i < N; ++i) std::fwrite(frame.mDescriptors.ptr(i), 1, 32, o);
    std::fwrite(frame.mvKeysRight.data(), sizeof(cv::KeyPoint), Nr, o);
    for (int i = 0; i < Nr; ++i) std::fwrite(frame.mDescriptorsRight.ptr(i), 1, 32, o);
    std::fwrite(frame.mvuRight.data(), 4, N, o);
    std::fwrite(frame.mvDepth.data(), 4, N, o);
    const cv::Mat p1 = left->mvImagePyramid[1]; // no SyncImagePyramid(): the member fills itself on first host access

    // a second frame on the same extractors (object-mode constructor), from the images rolled down by 7 rows
    std::vector<unsigned char> L2(L.size()), R2(R.size());
    for (int y = 0; y < H; ++y) {
        std::memcpy(&L2[(size_t)((y + 7) % H) * W], &L[(size_t)y * W], W);
        std::memcpy(&R2[(size_t)((y + 7) % H) * W], &R[(size_t)y * W], W);
    }
    cv::Mat imL2(H, W, CV_8UC1, L2.data()), imR2(H, W, CV_8UC1, R2.data()), imColor;
    Frame::ObjectStageHook = object_stage;
    const int slotMode = 4;
    Frame other(imL2, imR2, imColor, 0.3, left, right, nullptr, K, dist, bf, thDepth, slotMode);
    std::fwrite(&p1.cols, 4, 1, o); std::fwrite(&p1.rows, 4, 1, o);
    for (int r = 0; r < p1.rows; ++r) std::fwrite(p1.ptr(r), 1, p1.cols, o);
    const int dd = N >= 2 ? ORBmatcher::DescriptorDistance(frame.mDescriptors.row(0), frame.mDescriptors.row(1)) : -1;
    std::fwrite(&dd, 4, 1, o);
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) {
            const int n = (int)frame.mGrid[i][j].size();
            std::fwrite(&n, 4, 1, o);
            for (size_t v : frame.mGrid[i][j]) { const int iv = (int)v; std::fwrite(&iv, 4, 1, o); }
        }
    const std::vector<Frame::AreaQuery> qs = {{W * 0.5f, H * 0.5f, 40.f, -1, -1}, {W * 0.25f, H * 0.7f, 25.f, 1, 4}, {5.f, 5.f, 90.f, 0, -1}};
    std::vector<std::vector<size_t>> res;
    for (const auto &q : qs) res.push_back(frame.GetFeaturesInArea(q.x, q.y, q.r, q.minLevel, q.maxLevel));
    for (const auto &r : frame.GetFeaturesInAreas(qs)) res.push_back(r);
    for (const auto &r : res) {
        const int n = (int)r.size();
        std::fwrite(&n, 4, 1, o);
        for (size_t v : r) { const int iv = (int)v; std::fwrite(&iv, 4, 1, o); }
    }
    // local-map projection search: map points that re-observe the frame's own features with deterministic noise
    const int nMP = N + N / 2;
    std::vector<MapPoint> pool(nMP);
    std::vector<MapPoint *> vp(nMP);
    unsigned lcg = 12345u;
    auto rnd = [&lcg]() { lcg = lcg * 1664525u + 1013904223u; return (lcg >> 8) & 0xFFFF; };
    for (int i = 0; i < nMP; ++i) {
        const int src = (int)(rnd() % (unsigned)N);
        MapPoint &m = pool[i];
        m.mTrackProjX = frame.mvKeysUn[src].pt.x + ((int)(rnd() % 9) - 4) * 0.5f;
        m.mTrackProjY = frame.mvKeysUn[src].pt.y + ((int)(rnd() % 9) - 4) * 0.5f;
        m.mTrackProjXR = frame.mvuRight[src] > 0 ? frame.mvuRight[src] + ((int)(rnd() % 5) - 2) * 0.5f : m.mTrackProjX - 6.f;
        m.mbTrackInView = rnd() % 10 != 0;
        m.mbBad = rnd() % 25 == 0;
        m.nObs = rnd() % 5 == 0 ? 0 : 3;
        m.mnTrackScaleLevel = std::min(hdr[3] - 1, frame.mvKeysUn[src].octave + (int)(rnd() % 2));
        m.mTrackViewCos = rnd() % 2 ? 0.9991f : 0.95f;
        m.mDescriptor = frame.mDescriptors.row(src).clone();
        for (int b = (int)(rnd() % 40); b > 0; --b) { const unsigned bit = rnd() % 256; m.mDescriptor.ptr<unsigned char>()[bit >> 3] ^= (unsigned char)(1u << (bit & 7)); }
        vp[i] = &m;
    }
    std::fwrite(&nMP, 4, 1, o);
    for (int i = 0; i < nMP; ++i) {
        const MapPoint &m = pool[i];
        const float f4[4] = {m.mTrackProjX, m.mTrackProjY, m.mTrackProjXR, m.mTrackViewCos};
        const int i2[2] = {m.mnTrackScaleLevel, ((m.mbTrackInView && !m.mbBad) ? 1 : 0) | (m.nObs > 0 ? 2 : 0)};
        std::fwrite(f4, 4, 4, o); std::fwrite(i2, 4, 2, o);
    }
    for (int i = 0; i < nMP; ++i) std::fwrite(pool[i].mDescriptor.ptr<unsigned char>(), 1, 32, o);
    ORBmatcher matcher(0.8);
    const int nmatches = matcher.SearchByProjection(frame, vp, 3);
    std::fwrite(&nmatches, 4, 1, o);
    for (int idx = 0; idx < N; ++idx) {
        const int who = frame.mvpMapPoints[idx] ? (int)(frame.mvpMapPoints[idx] - pool.data()) : -1;
        std::fwrite(&who, 4, 1, o);
    }
    const int nObjL = (int)other.mvTempObjKeys.size(), nObjR = (int)other.mvTempObjKeysRight.size();
    std::fwrite(&nObjL, 4, 1, o); std::fwrite(&nObjR, 4, 1, o);
    std::fwrite(other.mvuTempObjKeysRight.data(), 4, nObjL, o);
    std::fwrite(other.mvTempObjDepth.data(), 4, nObjL, o);
    std::fwrite(&other.N, 4, 1, o);
    // dense object features (Frame::ExtractObjORB -> OpencvORBDetector, reference Frame.cc:2623-2665): both views on two
    // threads, object masks = two rectangles
    cv::Mat maskL(H, W, CV_8UC1), maskR(H, W, CV_8UC1);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const bool in = (y >= H / 4 && y < 3 * H / 4 && x >= W / 6 && x < W / 2) || (y >= H / 8 && y < H / 2 && x >= 2 * W / 3 && x < W - 40);
            maskL.ptr(y)[x] = in ? 255 : 0;
            maskR.ptr(y)[x] = in ? 255 : 0;
        }
    frame.ExtractObjORB(imL, imR, maskL, maskR);
    const int nOL = (int)frame.mvTempObjKeys.size(), nOR = (int)frame.mvTempObjKeysRight.size();
    { // the free function with the reference's name (single-image entry point) gives the batched call's left result
        std::vector<cv::KeyPoint> kp1;
        cv::Mat d1;
        OpencvORBDetector(imL, maskL, kp1, d1);
        if ((int)kp1.size() != nOL || (nOL && std::memcmp(static_cast<const void *>(kp1.data()), static_cast<const void *>(frame.mvTempObjKeys.data()), sizeof(cv::KeyPoint) * nOL)))
            return 8;
        for (int i = 0; i < nOL; ++i)
            if (std::memcmp(d1.ptr(i), frame.mTempObjPointsDescriptors.ptr(i), 32)) return 8;
    }
    std::fwrite(&nOL, 4, 1, o); std::fwrite(&nOR, 4, 1, o);
    std::fwrite(frame.mvTempObjKeys.data(), sizeof(cv::KeyPoint), nOL, o);
    for (int i = 0; i < nOL; ++i) std::fwrite(frame.mTempObjPointsDescriptors.ptr(i), 1, 32, o);
    std::fwrite(frame.mvTempObjKeysRight.data(), sizeof(cv::KeyPoint), nOR, o);
    for (int i = 0; i < nOR; ++i) std::fwrite(frame.mTempObjPointsDescriptorsRight.ptr(i), 1, 32, o);
    // Frame::ComputeBoW (reference Frame.cc:2040-2047) with the vocabulary file given as the third argument
    int nWords = -1, nFv = 0;
    if (argc > 3) {
        ORBVocabulary voc;
        if (!voc.loadFromTextFile(argv[3])) return 7;
        frame.mpORBvocabulary = &voc;
        frame.ComputeBoW();
        frame.mpORBvocabulary = nullptr;
        nWords = (int)frame.mBowVec.size(); nFv = (int)frame.mFeatVec.size();
        std::fwrite(&nWords, 4, 1, o); std::fwrite(&nFv, 4, 1, o);
        for (const auto &e : frame.mBowVec) { const int id = (int)e.first; std::fwrite(&id, 4, 1, o); std::fwrite(&e.second, 8, 1, o); }
        for (const auto &e : frame.mFeatVec) {
            const int id = (int)e.first, cnt = (int)e.second.size();
            std::fwrite(&id, 4, 1, o); std::fwrite(&cnt, 4, 1, o);
            std::fwrite(e.second.data(), 4, cnt, o);
        }
    }
    // Frame::AssignFeatures (reference Frame.cc:759-977) with the instance mask and detection ids given as the fourth
    // argument: [int nObj][int ids[nObj] (negative: null detection)][H * W mask bytes].  As in the object-mode constructors
    // (:725-733) the dense features get their stereo depth first.
    if (argc > 4) {
        FILE *mf = std::fopen(argv[4], "rb");
        int nObj = 0;
        if (!mf || std::fread(&nObj, 4, 1, mf) != 1 || nObj < 0 || nObj > 256) return 9;
        std::vector<int> ids((size_t)nObj);
        std::vector<unsigned char> mbytes((size_t)W * H);
        if ((nObj && std::fread(ids.data(), 4, nObj, mf) != (size_t)nObj) || std::fread(mbytes.data(), 1, mbytes.size(), mf) != mbytes.size()) return 9;
        std::fclose(mf);
        std::vector<DetectionObject> pool((size_t)nObj);
        std::vector<DetectionObject *> dets((size_t)nObj, nullptr);
        for (int j = 0; j < nObj; ++j)
            if (ids[j] >= 0) { pool[j].mnObjectID = ids[j]; dets[j] = &pool[j]; }
        frame.UndistortObjKeyPoints();
        frame.ComputeObjStereoMatches();
        const int nOL2 = (int)frame.mvTempObjKeys.size();
        std::fwrite(&nOL2, 4, 1, o);
        std::fwrite(frame.mvuTempObjKeysRight.data(), 4, nOL2, o);
        std::fwrite(frame.mvTempObjDepth.data(), 4, nOL2, o);
        frame.mMaskImg = cv::Mat(H, W, CV_8UC1, mbytes.data());
        frame.AssignFeatures(dets);
        const int nKept = (int)frame.mnDetObj, nb = (int)frame.mvKeys.size();
        std::fwrite(&nKept, 4, 1, o); std::fwrite(&nb, 4, 1, o);
        auto dump_set = [&](const std::vector<cv::KeyPoint> &k, const std::vector<cv::KeyPoint> &kun, const cv::Mat &d,
                            const std::vector<float> &ur, const std::vector<float> &dp) {
            const int n = (int)k.size();
            if ((int)kun.size() != n || (int)ur.size() != n || (int)dp.size() != n || (n ? d.rows : 0) != n) std::exit(10);
            std::fwrite(k.data(), sizeof(cv::KeyPoint), n, o);
            std::fwrite(kun.data(), sizeof(cv::KeyPoint), n, o);
            for (int i = 0; i < n; ++i) std::fwrite(d.ptr(i), 1, 32, o);
            std::fwrite(ur.data(), 4, n, o);
            std::fwrite(dp.data(), 4, n, o);
        };
        dump_set(frame.mvKeys, frame.mvKeysUn, frame.mDescriptors, frame.mvuRight, frame.mvDepth);
        for (int k = 0; k < nKept; ++k) {
            const int id = frame.mvDetectionObjects[k]->mnObjectID, nOri = (int)frame.mvOriKeys[k].size(), nDense = (int)frame.mvObjKeys[k].size();
            const int eff = (int)frame.mvnEffectiveObjPointDepthNums[k], init = (int)frame.mvDetectionObjects[k]->mnInFrameOrderSize;
            int gridTotal = 0, gridChecksum = 0;
            for (int i = 0; i < FRAME_GRID_COLS; ++i)
                for (int j = 0; j < FRAME_GRID_ROWS; ++j)
                    for (size_t v : frame.mvObjKeysGrid[k][i][j]) { ++gridTotal; gridChecksum += (int)v * (i * FRAME_GRID_ROWS + j + 1); }
            const int head[7] = {id, nOri, nDense, eff, init, gridTotal, gridChecksum};
            std::fwrite(head, 4, 7, o);
            dump_set(frame.mvOriKeys[k], frame.mvOriKeysUn[k], frame.mOriDescriptors[k], frame.mvuOriRight[k], frame.mvOriDepth[k]);
            dump_set(frame.mvObjKeys[k], frame.mvObjKeysUn[k], frame.mvObjPointsDescriptors[k], frame.mvuObjKeysRight[k], frame.mvObjPointDepth[k]);
        }
    }
    std::fclose(o);
    std::printf("shim ok: N=%d Nr=%d matched=%d mb=%g projection matches=%d\n", N, Nr,
                (int)std::count_if(frame.mvuRight.begin(), frame.mvuRight.end(), [](float v) { return v >= 0; }), frame.mb, nmatches);
    delete left; delete right;
    return 0;
}
