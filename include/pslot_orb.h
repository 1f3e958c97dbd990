/*
 * pslot_orb.h — C ABI of libpslot_orb.so: PointSLOT's per-frame stereo ORB front end on B200.
 *
 * The reference (pkzhou/PointSLOT) has no FFI layer; its boundary for this path is the C++ class
 * surface that Tracking / Frame compile against.  Each entry point below names the reference
 * interface it replaces (paths under /root/reference).  The C++ drop-in classes that keep the
 * reference signatures verbatim and forward here live in pointslot_b200/cpp/ (see INTEGRATION.md).
 *
 * Conventions: every function returns 0 on success and a negative pslot_status on failure, never
 * throws, and leaves a thread-local message retrievable with pslot_last_error().  Plain pointers and
 * sizes only.  There is no CPU fallback: without a CUDA device every compute call fails with
 * PSLOT_ERR_CUDA.  Contexts are not re-entrant; distinct contexts may be used from distinct host
 * threads concurrently (each owns a CUDA stream), as Frame.cc:161-164 does with its two extractors.
 */
#ifndef PSLOT_ORB_H
#define PSLOT_ORB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSLOT_ABI_VERSION 1

typedef enum pslot_status {
    PSLOT_OK = 0,
    PSLOT_ERR_ARG = -1,      /* bad argument (null pointer, non-positive size, unsupported parameter) */
    PSLOT_ERR_CAPACITY = -2, /* caller's output capacity too small; *n holds the required count */
    PSLOT_ERR_CUDA = -3,     /* CUDA runtime / driver error (no device, launch failure, OOM) */
    PSLOT_ERR_SHAPE = -4,    /* image too small for the pyramid (coarsest level needs >= 62 x 62 px) */
    PSLOT_ERR_STATE = -5     /* call sequence error (e.g. stereo before extraction) */
} pslot_status;

/* Layout-identical to cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id. */
typedef struct pslot_keypoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} pslot_keypoint;

typedef struct pslot_orb pslot_orb;     /* one extractor == one ORB_SLAM2::ORBextractor object */
typedef struct pslot_batch pslot_batch; /* throughput context: batches of independent stereo pairs */

const char *pslot_last_error(void);
int pslot_abi_version(void);
/* Number of visible CUDA devices, or a negative pslot_status. */
int pslot_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * Extractor — replaces class ORBextractor (include/ORBextractor.h:45-111, src/ORBextractor.cc).
 * --------------------------------------------------------------------------------------------- */

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (ORBextractor.cc:410-470).  device = CUDA ordinal. */
int pslot_orb_create(pslot_orb **ctx, int nfeatures, float scale_factor, int nlevels,
                     int ini_th_fast, int min_th_fast, int device);
void pslot_orb_destroy(pslot_orb *ctx);

/* GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (ORBextractor.h:63-83) and mnFeaturesPerLevel (:103).
 * Each non-null array receives nlevels entries. */
int pslot_orb_levels(const pslot_orb *ctx);
int pslot_orb_tables(const pslot_orb *ctx, float *scale, float *inv_scale, float *sigma2,
                     float *inv_sigma2, int32_t *features_per_level);
/* Upper bound on the number of keypoints one image can yield (the reference overshoots nfeatures by
 * up to 3 per level, ORBextractor.cc:669-737): size kps / desc32 with this. */
int pslot_orb_max_keypoints(const pslot_orb *ctx);

/* ORBextractor::operator()(image, mask, keypoints, descriptors); the mask is ignored there too
 * (ORBextractor.cc:1043-1105).  img: CV_8UC1, h rows of w bytes, row stride `pitch` bytes (host
 * memory).  Writes *n keypoints (level-ascending, quadtree list order inside a level) and *n x 32
 * descriptor bytes.  Synchronous.  Empty image (w<=0 || h<=0 || !img): *n = 0, returns PSLOT_OK and
 * leaves kps/desc32 untouched (:1046-1047). */
int pslot_orb_extract(pslot_orb *ctx, const uint8_t *img, int w, int h, ptrdiff_t pitch,
                      pslot_keypoint *kps, uint8_t *desc32, int cap, int *n);

/* Colour ingest: the cvtColor(RGB/BGR(A) -> GRAY) that Tracking::GrabImageStereo applies
 * before building the Frame (src/Tracking.cc:1014-1039), fused with the upload.  img: interleaved 8-bit,
 * channels = 3 or 4 (alpha ignored), rgb_order != 0 for RGB(A) (Camera.RGB: 1), 0 for BGR(A).  Arithmetic =
 * OpenCV 4.13: Y = (R*9798 + G*19235 + B*3735 + 2^14) >> 15.  channels = 1 is pslot_orb_extract. */
int pslot_orb_extract_color(pslot_orb *ctx, const uint8_t *img, int w, int h, ptrdiff_t pitch, int channels,
                            int rgb_order, pslot_keypoint *kps, uint8_t *desc32, int cap, int *n);
/* The same extraction in two halves, so that ONE host thread can have both views in flight (the reference spawns two
 * threads per frame for that, Frame.cc:161-164): _begin uploads the image and enqueues every kernel and the result
 * download on the context's stream, _end waits and hands the results over exactly as pslot_orb_extract does.  One
 * extraction per context at a time; the image may be released once _begin returns. */
int pslot_orb_extract_begin(pslot_orb *ctx, const uint8_t *img, int w, int h, ptrdiff_t pitch);
int pslot_orb_extract_end(pslot_orb *ctx, pslot_keypoint *kps, uint8_t *desc32, int cap, int *n);
/* Identity of the feature set the context's last extraction left on the device (0: none); changes with every
 * extraction.  pslot_stereo_match_resident = pslot_stereo_match on the two sets as they lie on the device - what
 * Frame::ComputeStereoMatches does in the Frame constructors, right after the two extractions (reference
 * src/Frame.cc:161-173): nothing is uploaded.  The caller passes the tokens it took after the extractions;
 * PSLOT_ERR_STATE when either context has moved on (use pslot_stereo_match with the host arrays then).  *n_l = number of
 * left keypoints = entries written to u_right / depth (PSLOT_ERR_CAPACITY with *n_l = needed when cap is smaller). */
uint64_t pslot_orb_extract_token(const pslot_orb *ctx);
int pslot_stereo_match_resident(pslot_orb *ctx_l, pslot_orb *ctx_r, uint64_t token_l, uint64_t token_r, float bf,
                                float max_d, float *u_right, float *depth, int cap, int *n_l);

/* Backing store of the public member std::vector<cv::Mat> mvImagePyramid (ORBextractor.h:85,
 * filled by ComputePyramid, ORBextractor.cc:1107-1132).  Copies level `level` of the most recent
 * extraction to host memory (dst may be NULL to query the size only).  The 19-px replicated frame of
 * the reference is not materialised: no consumer reads it. */
int pslot_orb_level(pslot_orb *ctx, int level, uint8_t *dst, ptrdiff_t pitch, int *w, int *h);

/* ---------------------------------------------------------------------------------------------
 * Stereo matcher — replaces Frame::ComputeStereoMatches (src/Frame.cc:2142-2316).  Also covers
 * Frame::ComputeObjStereoMatches (Frame.cc:2318-2503), which is the same routine on the object
 * keypoint subset, because the keypoint/descriptor sets are explicit arguments.
 *
 * Reads the un-blurred pyramids of the most recent extraction held by ctx_l / ctx_r (the reference
 * reads mpORBextractorLeft/Right->mvImagePyramid, Frame.cc:2238,2256).  bf = mbf.  max_d replaces
 * the reference's `mbf/mb` with `mb` uninitialised at that point (Frame.cc:2172-2174 vs :197):
 * pass mbf/(mbf/fx) for the intended behaviour, INFINITY for "mb == 0".
 * u_right / depth receive n_l floats each (-1 = no match), i.e. mvuRight / mvDepth.
 * --------------------------------------------------------------------------------------------- */
int pslot_stereo_match(pslot_orb *ctx_l, pslot_orb *ctx_r,
                       const pslot_keypoint *kps_l, int n_l, const uint8_t *desc_l,
                       const pslot_keypoint *kps_r, int n_r, const uint8_t *desc_r,
                       float bf, float max_d, float *u_right, float *depth);

/* Page-locked (pinned) host memory for the images handed to pslot_batch_submit and for its output arrays: transfers
 * from pageable memory are staged by the driver and run at a fraction of the PCIe rate.  mode 0: plain pinned pages;
 * 1: write-combined pages (slow for the CPU to read back: input buffers only); 2: a transparent-huge-page backed mapping
 * registered with the driver (fewer translation entries per transfer on virtualised hosts). */
int pslot_host_alloc(void **ptr, size_t bytes, int mode);
int pslot_host_free(void *ptr);

/* static int ORBmatcher::DescriptorDistance(const cv::Mat&, const cv::Mat&) (ORBmatcher.cc:2704-2720):
 * Hamming distance of two 32-byte descriptors.  Host-side (one popcount pair is not device work);
 * the device matcher uses the same definition inside the stereo kernel. */
int pslot_hamming256(const void *a, const void *b);

/* ---------------------------------------------------------------------------------------------
 * Feature grid — void Frame::AssignFeaturesToGrid() with bool Frame::PosInGrid(kp, posX,
 * posY) (Frame.cc:1636-1655, 2027-2038) and its consumer vector<size_t> Frame::GetFeaturesInArea(x, y, r,
 * minLevel, maxLevel) (Frame.cc:1808-1861).
 * mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS] is returned in CSR form: cell = i * rows + j for mGrid[i][j],
 * cell_start[cols * rows + 1] offsets into `indices`, which lists the keypoint indices of each cell in
 * push_back (= ascending) order; keypoints outside the grid appear nowhere.
 * spec: min_x = mnMinX, min_y = mnMinY, inv_w = mfGridElementWidthInv, inv_h = mfGridElementHeightInv
 * (Frame.cc:184-185), cols / rows = FRAME_GRID_COLS / FRAME_GRID_ROWS (Frame.h:40-41). */
typedef struct pslot_grid_spec {
    float min_x, min_y, inv_w, inv_h;
    int32_t cols, rows;
} pslot_grid_spec;
typedef struct pslot_area_query {
    float x, y, r;
    int32_t min_level, max_level; /* as GetFeaturesInArea: levels are checked iff min_level > 0 || max_level >= 0 */
} pslot_area_query;
/* kps_un = mvKeysUn (n entries, host).  cell_start / indices (host, [cols*rows+1] / [n]) may be NULL to keep the
 * grid on the device only; *n_in_grid (optional) receives the number of keypoints inside the grid.  The grid and
 * the keypoints stay on the device for pslot_features_in_area. */
int pslot_assign_grid(pslot_orb *ctx, const pslot_keypoint *kps_un, int n, const pslot_grid_spec *spec,
                      int32_t *cell_start, int32_t *indices, int *n_in_grid);
/* Identity of the keypoint set currently gridded on `ctx`: a process-unique non-zero number assigned by every
 * successful pslot_assign_grid (0 = nothing gridded).  pslot_features_in_area and the projection searches run against
 * whatever set was gridded LAST on the context; a caller that keeps several frames per extractor (the reference reads
 * F.mGrid / F.mvKeysUn of the frame it is handed, ORBmatcher.cc:92) stores the token with each frame and re-grids when
 * it no longer matches - that is what the C++ Frame / ORBmatcher drop-ins do (Frame::EnsureDeviceGrid). */
uint64_t pslot_grid_token(const pslot_orb *ctx);
/* nq area queries against the grid of the last pslot_assign_grid on ctx.  counts[q] = number of features found
 * (may exceed cap); indices[q * cap + k], k < min(counts[q], cap), in the reference's visiting order. */
int pslot_features_in_area(pslot_orb *ctx, const pslot_area_query *queries, int nq, int cap, int32_t *counts,
                           int32_t *indices);

/* ---------------------------------------------------------------------------------------------
 * Local-map projection search — int ORBmatcher::SearchByProjection(Frame &F, const
 * vector<MapPoint*> &vpMapPoints, const float th) (include/ORBmatcher.h, src/ORBmatcher.cc:68-155), called by
 * Tracking::SearchLocalPoints (src/Tracking.cc:3354).
 * The frame side is what pslot_assign_grid left on ctx (mvKeysUn + mGrid) plus the arrays below; a map point is the
 * set of MapPoint fields the routine reads (include/MapPoint.h:110-120). */
typedef struct pslot_map_point {
    float proj_x, proj_y, proj_xr, view_cos; /* mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos */
    int32_t level;                           /* mnTrackScaleLevel */
    int32_t flags;                           /* bit 0: mbTrackInView && !isBad()  (searched at all)
                                                bit 1: Observations() > 0        (its match blocks later points) */
} pslot_map_point;
/* mp_desc [n_mp*32] = pMP->GetDescriptor(); frame_desc [n*32] = F.mDescriptors, u_right [n] = F.mvuRight,
 * occupied [n] (may be NULL) != 0 where F.mvpMapPoints[idx] already holds a point with Observations() > 0;
 * n must equal the keypoint count of the last pslot_assign_grid; scale_factors [nlevels] = F.mvScaleFactors;
 * nn_ratio = mfNNratio.  match_of_mp[i] receives the feature index assigned to map point i (the reference's
 * F.mvpMapPoints[bestIdx] = pMP, to be applied in map-point order) or -1; *nmatches the return value. */
int pslot_search_by_projection(pslot_orb *ctx, const pslot_map_point *mps, const uint8_t *mp_desc, int n_mp,
                               const uint8_t *frame_desc, const float *u_right, const uint8_t *occupied, int n,
                               const float *scale_factors, int nlevels, float th, float nn_ratio,
                               int32_t *match_of_mp, int *nmatches);

/* Replaces `int ORBmatcher::SearchByProjection(Frame &F, const size_t &nOrder, const vector<MapObjectPoint*> &vpMapPoints,
 * const float th)` (reference src/ORBmatcher.cc:157-246): the same search over ONE detected object's feature set.  Grid
 * that set first - pslot_assign_grid on F.mvObjKeysUn[nOrder] reproduces F.mvObjKeysGrid[nOrder], and
 * Frame::GetObjectFeaturesInArea (src/Frame.cc:1892-1951) is GetFeaturesInArea on it - then pass
 * F.mvObjPointsDescriptors[nOrder], F.mvuObjKeysRight[nOrder] and occupied[idx] = (F.mvpMapObjectPoints[nOrder][idx] has
 * observations) || !F.isInBBox(nOrder, keypoint) (:206-211).  The overload's own constants: a fixed 5 px search window
 * (:179), levels nPredictedLevel - 1 .. + 1, TH_HIGH_FORDYNAMIC = 130 (:236); the stereo test keeps r * scale (:215). */
int pslot_search_by_projection_obj(pslot_orb *ctx, const pslot_map_point *mps, const uint8_t *mp_desc, int n_mp,
                                   const uint8_t *frame_desc, const float *u_right, const uint8_t *occupied, int n,
                                   const float *scale_factors, int nlevels, float th, float nn_ratio,
                                   int32_t *match_of_mp, int *nmatches);
/* int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
 * (src/ORBmatcher.cc:1613-1756), called by Tracking::TrackWithMotionModel (src/Tracking.cc:3047, 3053).  The caller
 * keeps the reference's projection lines (:1623-1676: pose algebra on cv::Mat, bounds test, level window) and hands
 * over one record per LastFrame keypoint that carries a map point; the search, the assignment order dependence
 * and the rotation-histogram consistency check (:1720-1753, ComputeThreeMaxima :2658-2699) happen here. */
typedef struct pslot_proj_point {
    float u, v;      /* projection into the current frame (:1656-1657) */
    float ur;        /* u - CurrentFrame.mbf * invzc (:1690) */
    float radius;    /* th * CurrentFrame.mvScaleFactors[nLastOctave] (:1667) */
    int32_t min_level, max_level; /* the level arguments of the GetFeaturesInArea call taken at :1671-1676
                                     (forward: nLastOctave, -1; backward: 0, nLastOctave; else nLastOctave-1, nLastOctave+1) */
    float angle;     /* LastFrame.mvKeysUn[i].angle (:1722) */
    int32_t flags;   /* bit 0: the point reaches the search (pMP && !mvbOutlier[i] && invzc >= 0 && inside the bounds)
                        bit 1: pMP->Observations() > 0 */
} pslot_proj_point;
/* match_of_pt[i] = bestIdx2 of record i or -1; rejected[i] != 0 where the rotation check removed the match again
 * (CurrentFrame.mvpMapPoints[match_of_pt[i]] = NULL after all assignments have been applied in order);
 * *nmatches = the reference's return value.  check_orientation = mbCheckOrientation.  Frame arrays as above. */
int pslot_search_by_projection_last(pslot_orb *ctx, const pslot_proj_point *pts, const uint8_t *pt_desc, int n_pts,
                                    const uint8_t *frame_desc, const float *u_right, const uint8_t *occupied, int n,
                                    int check_orientation, int32_t *match_of_pt, uint8_t *rejected, int *nmatches);

/* Replaces `int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound,
 * const float th, const int ORBdist)` (reference src/ORBmatcher.cc:2529-2654; relocalisation, called from
 * Tracking::Relocalization) from the GetFeaturesInArea call on.  The caller keeps the projection lines (:2551-2581) and
 * passes one pslot_proj_point per pKF->GetMapPointMatches() entry: u, v, radius = th * mvScaleFactors[nPredictedLevel],
 * min_level / max_level = nPredictedLevel -+ 1, angle = pKF->mvKeysUn[i].angle, flags bit 0 = the point reaches the search
 * (pMP && !isBad && !sAlreadyFound.count(pMP), inside the image bounds and the distance range); `ur` and flags bit 1 are
 * ignored.  Differences from the last-frame overload, all the reference's: no stereo test, ANY association blocks a
 * feature (occupied[idx] = CurrentFrame.mvpMapPoints[idx] != NULL) and so does every match made on the way, acceptance at
 * bestDist <= orb_dist, rotation bins of 30 degrees (`factor = 1.0f/HISTO_LENGTH`, :2541).
 * The loop-closing overload `SearchByProjection(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints,
 * vector<MapPoint*> &vpMatched, int th)` (:413-527) is the same search on the key frame's gridded keypoints
 * (KeyFrame::GetFeaturesInArea, src/KeyFrame.cc:665-704, is Frame's lookup without the level arguments) with
 * min_level / max_level = nPredictedLevel - 1, nPredictedLevel (:499-502), orb_dist = TH_LOW, check_orientation = 0 and
 * occupied[idx] = vpMatched[idx] != NULL; flags bit 0 = the point passed :438-480. */
int pslot_search_by_projection_kf(pslot_orb *ctx, const pslot_proj_point *pts, const uint8_t *pt_desc, int n_pts,
                                  const uint8_t *frame_desc, const uint8_t *occupied, int n, int orb_dist,
                                  int check_orientation, int32_t *match_of_pt, uint8_t *rejected, int *nmatches);
/* Number of device sweeps the last projection search needed to reproduce the sequential result (>= 2). */
int pslot_debug_projection_sweeps(const pslot_orb *ctx);

/* ---------------------------------------------------------------------------------------------
 * Throughput API — independent stereo pairs of one geometry, processed `max_pairs` at a time on one
 * GPU stream: both extractions (Frame.cc:161-164) + ComputeStereoMatches (:173) per pair, with no
 * host round trip in between.  This is the data-parallel form of the Frame stereo constructor path
 * used for batch / multi-GPU processing; pairs never interact.
 * --------------------------------------------------------------------------------------------- */
/* stream: a cudaStream_t to run on (e.g. torch.cuda.Stream().cuda_stream), or NULL to create one. */
int pslot_batch_create(pslot_batch **b, int nfeatures, float scale_factor, int nlevels,
                       int ini_th_fast, int min_th_fast, int width, int height, int max_pairs,
                       int device, void *stream);
void pslot_batch_destroy(pslot_batch *b);
/* Route the host->device uploads of this context through `stream` (a cudaStream_t shared by all the batch
 * contexts of a GPU; NULL restores per-context uploads).  Concurrent uploads from several streams interleave on
 * the DMA engine and lose ~35 % of the PCIe rate; one upload stream keeps every transfer at full rate while the
 * contexts' own streams overlap the kernels.  Ordering is enforced with events inside pslot_batch_submit. */
int pslot_batch_set_copy_stream(pslot_batch *b, void *stream);
int pslot_batch_max_keypoints(const pslot_batch *b); /* per-image capacity K of the output arrays */
void *pslot_batch_stream(const pslot_batch *b);      /* the cudaStream_t all work is enqueued on */

/* Host-memory outputs of one batch; image i of view v owns rows [ (v*npairs+i)*K, ... +K ).
 *   counts  [2*npairs]          keypoints per image (left images first, then right images)
 *   kps     [2*npairs*K]        pslot_keypoint
 *   desc    [2*npairs*K*32]     descriptors
 *   u_right [npairs*K], depth [npairs*K]   per LEFT keypoint, -1 = none
 * Pointers may be pageable or pinned (pinned makes the copies asynchronous). */
typedef struct pslot_batch_out {
    int32_t *counts;
    pslot_keypoint *kps;
    uint8_t *desc;
    float *u_right;
    float *depth;
} pslot_batch_out;

/* Enqueue one batch: left / right point to npairs images each (image stride `image_stride` bytes,
 * row stride `pitch` bytes), in HOST memory when host_input != 0, else in DEVICE memory.  With
 * out == NULL the results stay on the device (kernel-only timing).  Asynchronous with respect to
 * the host; call pslot_batch_wait before reading `out` or reusing the input buffers. */
int pslot_batch_submit(pslot_batch *b, const uint8_t *left, const uint8_t *right, ptrdiff_t pitch,
                       ptrdiff_t image_stride, int npairs, int host_input, float bf, float max_d,
                       const pslot_batch_out *out);
/* Same for interleaved colour frames in HOST memory (channels 3 or 4; see pslot_orb_extract_color). */
int pslot_batch_submit_color(pslot_batch *b, const uint8_t *left, const uint8_t *right, ptrdiff_t pitch,
                             ptrdiff_t image_stride, int npairs, int channels, int rgb_order, float bf,
                             float max_d, const pslot_batch_out *out);
/* AssignFeaturesToGrid for the LEFT keypoints of every pair of the last submit (rectified input: mvKeysUn ==
 * mvKeys), enqueued on the batch stream.  cell_start [npairs * (cols*rows+1)] / indices [npairs * K] are host
 * buffers filled asynchronously (valid after pslot_batch_wait); either may be NULL. */
int pslot_batch_assign_grid(pslot_batch *b, const pslot_grid_spec *spec, int32_t *cell_start, int32_t *indices);
int pslot_batch_wait(pslot_batch *b);
/* Number of kernels pslot_batch_submit launches per call (for launch accounting). */
int pslot_batch_kernel_launches(const pslot_batch *b);
/* Device pointers of the last batch's results (same layout as pslot_batch_out), valid until the next
 * submit; for callers that keep the features on the GPU. */
int pslot_batch_device_out(const pslot_batch *b, pslot_batch_out *dev);

/* Per-stage device timing of a submit (CUDA events on the batch stream, recorded only when enabled).
 * Stages: 0 ingest copy (H2D or D2D), 1 pyramid (nlevels-1 launches), 2 FAST cells, 3 quadtree,
 * 4 orientation+blur+descriptor, 5 stereo match, 6 stereo reject, 7 result copy (D2H).
 * pslot_batch_stage_times is valid after pslot_batch_wait; ms receives PSLOT_NUM_STAGES floats. */
#define PSLOT_NUM_STAGES 8
int pslot_batch_set_profiling(pslot_batch *b, int enable);
int pslot_batch_stage_times(pslot_batch *b, float *ms);
/* Work counters of the last batch (read back after pslot_batch_wait; not on the timed path):
 * stats[0] FAST candidates over all images, [1] keypoints over all images, [2] left keypoints that
 * reached the SAD refinement (best Hamming < 75), [3] stereo matches that survived the median test. */
int pslot_batch_stats(pslot_batch *b, int64_t stats[4]);
/* Stereo search counters accumulated by the submits made while profiling was enabled, then reset:
 * out[0] = row-band candidates scanned, i.e. the sum over left keypoints of |vRowIndices[row]| (Frame.cc:2188),
 * out[1] = descriptor pairs compared (Frame.cc:2217).  With pslot_batch_stats they give the measured c_row and
 * c-bar of SURVEY.md §8d. */
int pslot_batch_stereo_counters(pslot_batch *b, int64_t out[2]);

/* ---------------------------------------------------------------------------------------------
 * OpenCV's own ORB for the dense object features.
 * Replaces `cv::ORB::create(1000, 1.2, 8, 19)->detectAndCompute(im, ObjMask, kp, descriptor)`
 * (reference src/Frame.cc:2623-2627 OpencvORBDetector, run for both views by Frame::ExtractObjORB :2628-2665):
 * HARRIS_SCORE, patchSize 31, WTA_K 2, firstLevel 0.  `mask` may be NULL; non-zero = keep, on every pyramid level (cv::ORB
 * reads its mask as zero / non-zero; the reference builds 0 / 255 masks, Frame.cc:2637-2646).  edge_threshold >= 19.
 * Output: cv::KeyPoint-compatible records and 32-byte descriptors ordered by (octave, y, x) - cv::ORB's own order is
 * whatever std::nth_element leaves and is not part of the contract; the SET and every field are OpenCV 4.13's, bit for
 * bit (tests compare with cv2.ORB_create directly).  Returns PSLOT_ERR_CAPACITY with *n = needed when cap is too small.
 * One context per calling thread (ExtractObjORB runs left and right on two threads).
 * --------------------------------------------------------------------------------------------- */
typedef struct pslot_cvorb pslot_cvorb;
int pslot_cvorb_create(pslot_cvorb **ctx, int nfeatures, float scale_factor, int nlevels, int edge_threshold,
                       int fast_threshold, int device);
void pslot_cvorb_destroy(pslot_cvorb *ctx);
int pslot_cvorb_detect_and_compute(pslot_cvorb *ctx, const uint8_t *img, int w, int h, ptrdiff_t pitch,
                                   const uint8_t *mask, ptrdiff_t mask_pitch, pslot_keypoint *kps, uint8_t *desc32,
                                   int cap, int *n);
/* Batched form: n_images images of one size through ONE set of launches and one result round trip - the left and the
 * right view of Frame::ExtractObjORB (reference src/Frame.cc:2648-2651 runs them on two threads), or the object frames
 * of a whole sequence.  imgs[s] / masks[s] are the s-th image / object mask (`masks` may be NULL, and so may any
 * masks[s]: that image is searched everywhere); all share w, h and the two pitches.  Results of image s go to
 * kps[s * cap ..] / desc32[(s * cap) * 32 ..], counts[s] keypoints; each image's result is exactly what the single-image
 * call returns for it.  PSLOT_ERR_CAPACITY when some counts[s] > cap (counts are filled in: retry with the largest). */
int pslot_cvorb_detect_and_compute_batch(pslot_cvorb *ctx, int n_images, const uint8_t *const *imgs, int w, int h,
                                         ptrdiff_t pitch, const uint8_t *const *masks, ptrdiff_t mask_pitch,
                                         pslot_keypoint *kps, uint8_t *desc32, int cap, int *counts);
/* The same on images that are ALREADY on the device: images first_image .. first_image + n_images - 1 of the last
 * pslot_batch_submit on `b` (0 .. npairs-1 = the left views, npairs .. 2*npairs-1 = the right views), read from level 0
 * of the batch's pyramids - in object mode (Frame.cc:706-713) both extractors look at the same two images, so they need
 * not cross PCIe twice; only the masks (at most) are uploaded.  Call after pslot_batch_wait and before the next submit
 * on `b`.  Results as pslot_cvorb_detect_and_compute_batch. */
int pslot_cvorb_detect_and_compute_resident(pslot_cvorb *ctx, pslot_batch *b, int first_image, int n_images,
                                            const uint8_t *const *masks, ptrdiff_t mask_pitch, pslot_keypoint *kps,
                                            uint8_t *desc32, int cap, int *counts);
/* Frame::ComputeObjStereoMatches (reference src/Frame.cc:2318-2503) for every pair of the batch at once: the stereo
 * matcher on the cv::ORB features that pslot_cvorb_detect_and_compute_resident(ctx, b, 0, 2 * npairs, ...) just left on
 * the device (left views = images 0 .. npairs-1, right views = npairs .. 2*npairs-1), reading the pyramids of `b`'s last
 * submit.  u_right / depth: npairs rows of `cap` floats; row p holds one entry per left cv::ORB keypoint of pair p, in the
 * order the resident call returned them.  PSLOT_ERR_STATE unless the context holds exactly those features. */
int pslot_cvorb_stereo_match_resident(pslot_cvorb *ctx, pslot_batch *b, float bf, float max_d, float *u_right, float *depth,
                                      int cap);

/* ---------------------------------------------------------------------------------------------
 * Bag of words (second half).
 * Replaces `void Frame::ComputeBoW()` (reference src/Frame.cc:2040-2047):
 *     mpORBvocabulary->transform(Converter::toDescriptorVector(mDescriptors), mBowVec, mFeatVec, 4)
 * with ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (include/ORBVocabulary.h:30; transform:
 * Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1132-1201, 1223-1266).
 *
 * pslot_vocab_create takes the tree as flat arrays - what loadFromTextFile (:1389-1474) builds from Vocabulary/ORBvoc.txt:
 * node 0 = root, children of node i = child[child_start[i] .. child_start[i+1]) in file order, a node without children is
 * a leaf with word_id >= 0 and its weight.  scoring: 0 L1_NORM .. 5 DOT_PRODUCT, weighting: 0 TF_IDF 1 TF 2 IDF 3 BINARY
 * (BowVector.h:36-57).
 * pslot_bow_transform: the tree walk on the device, one result pair per descriptor: the leaf reached and the node at
 * level L - levelsup (desc32 = n x 32 host bytes; or NULL with `ctx` = an extractor whose last pslot_orb_extract left
 * its n descriptors on the device - they are read there).
 * pslot_bow_vectors: mBowVec as (word id ascending, value) and mFeatVec as CSR (node id ascending, start, feature
 * indices in feature order) from those pairs, in DBoW2's own double arithmetic (addWeight accumulation in feature
 * order, L1 / L2 normalisation in id order) - values are bit-identical to the reference's.  Arrays hold n entries
 * (fv_start n + 1).
 * --------------------------------------------------------------------------------------------- */
typedef struct pslot_vocab pslot_vocab;
int pslot_vocab_create(pslot_vocab **voc, int n_nodes, int k, int L, int scoring, int weighting, const uint8_t *desc32,
                       const int32_t *child_start, const int32_t *child, const int32_t *word_id, const double *weight,
                       int device);
void pslot_vocab_destroy(pslot_vocab *voc);
int pslot_bow_transform(pslot_vocab *voc, pslot_orb *ctx, const uint8_t *desc32, int n, int levelsup, int32_t *leaf_node,
                        int32_t *nid);
int pslot_bow_vectors(const pslot_vocab *voc, const int32_t *leaf_node, const int32_t *nid, int n, int32_t *bow_id,
                      double *bow_val, int *n_words, int32_t *fv_node, int32_t *fv_start, int32_t *fv_feat, int *n_fv_nodes);

/* ---------------------------------------------------------------------------------------------
 * Feature partition by the instance mask (the step after the stereo matchers in the object-mode Frame constructors).
 * Replaces the classification + push_back loops of `void Frame::AssignFeatures(vector<DetectionObject*>&)`
 * (reference src/Frame.cc:759-977; called :478, :648, :733): feature i is looked up at mMaskImg((int)pt.y, (int)pt.x);
 * label 0 -> background, 255 -> dropped, L in 1..254 -> the first non-null detection j whose id equals L - 1, with
 * id = mnObjectID > 255 ? mnObjectID - 255 : mnObjectID (:836-846), else dropped.  The same routine serves the frame's
 * own ORB features (-> mvKeys / mvOriKeys[j] ...) and the dense cv::ORB object features (-> mvObjKeys[j] ...; pass
 * `depth` = mvTempObjDepth and `depth_limit` = 2 * mThDepth to get vnEffectiveObjPointDepthNums, :893-896).
 *   object_ids[j] = vDetectionObjects[j]->mnObjectID, or any negative value for a NULL detection;
 *   cls[i]        = -1 dropped | 0 background | 1 + j;
 *   order[]       = feature indices grouped by class (background, detection 0, 1, ...), original order inside a class -
 *                   exactly the reference's push_back order; class c owns order[class_start[c] .. class_start[c + 1]);
 *                   class_start has n_objects + 2 entries, the last = number of kept features;
 *   effective[j]  = features of detection j with 0 < depth < depth_limit (NULL, or `depth` NULL: not computed).
 * `mask` may be NULL to reuse the mask the previous call uploaded (the two calls of one frame share mMaskImg).
 * The per-object grids (AssignDetObjFeasToGrid, :1863-1888) are pslot_assign_grid on each object's slice.
 * pslot_batch_assign_features: the same for the LEFT keypoints of every pair of the last submit, which are still on the
 * device (masks: one per pair; object_ids: n_objects per pair); cls / order hold K entries per pair (K = the batch's
 * keypoint capacity), class_start n_objects + 2 per pair; results are valid after pslot_batch_wait.
 * --------------------------------------------------------------------------------------------- */
int pslot_assign_features(pslot_orb *ctx, const pslot_keypoint *kps, int n, const uint8_t *mask, int w, int h,
                          ptrdiff_t mask_pitch, const int32_t *object_ids, int n_objects, const float *depth,
                          float depth_limit, int32_t *cls, int32_t *order, int32_t *class_start, int32_t *effective);
int pslot_batch_assign_features(pslot_batch *b, const uint8_t *masks, ptrdiff_t mask_pitch, ptrdiff_t mask_image_stride,
                                const int32_t *object_ids, int n_objects, int32_t *cls, int32_t *order,
                                int32_t *class_start);

/* Host-clock phases of the last (batched) call in ms: [0] uploads enqueued, [1] kernels + downloads enqueued, [2] waiting
 * for the device, [3] results copied to the caller's arrays. */
int pslot_cvorb_last_timing(const pslot_cvorb *ctx, double out_ms[4]);

/* Stage accessor of the last pslot_cvorb_detect_and_compute (image 0 of a batch), for the parity tests: what = 0 the pyramid level (w*h bytes,
 * tightly packed), 1 its raw FAST score map (0 where not a corner, and outside the band [edge-1, dim-edge] that the
 * border filter and the maxima test can see), 2 the level's candidates after the maxima test, the mask
 * and the border filter (packed u32 y:12 | x:12 | score:8, unordered), 3 the candidates retainBest(2*quota) kept (same
 * packing, unordered; their Harris responses go to `aux`).  *w, *h = level size; *n = entries written. */
int pslot_cvorb_debug(pslot_cvorb *ctx, int what, int level, void *out, float *aux, size_t cap_bytes, int *w, int *h, int *n);

/* ---------------------------------------------------------------------------------------------
 * Stage accessors for the parity tests (not needed by a drop-in caller).
 * --------------------------------------------------------------------------------------------- */
/* FAST candidates handed to the quadtree for `level` of the last pslot_orb_extract: window-relative
 * x, y and the FAST response, in the reference's order (cells row-major, raster inside a cell;
 * ORBextractor.cc:789-829). */
int pslot_orb_debug_candidates(pslot_orb *ctx, int level, int cap, int32_t *x, int32_t *y,
                               int32_t *resp, int *n);
/* Keypoints of `level` after DistributeOctTree + border offset (level coordinates, list order;
 * ORBextractor.cc:831-847) with their IC_Angle orientation (:850-852). */
int pslot_orb_debug_level_keypoints(pslot_orb *ctx, int level, int cap, int32_t *x, int32_t *y,
                                    int32_t *resp, float *angle, int *n);
/* Per-left-keypoint intermediates of the last pslot_stereo_match: best Hamming distance (100 = none),
 * index of the chosen right keypoint (-1 = none), SAD of accepted matches before the median test (-1). */
int pslot_stereo_debug(pslot_orb *ctx_l, int n_l, int32_t *best_dist, int32_t *best_idx, int32_t *best_sad);

/* The descriptor steering coefficients a = (float)cos(angle*pi/180), b = (float)sin(...) exactly as the
 * device computes them (ORBextractor.cc:112-113), for n host angles in degrees; lets a test sweep every
 * float in [0, 360] against libm. */
int pslot_debug_steer(int device, const float *angle_deg, int n, float *a, float *b);

/* layout checks (plain C: negative array size on mismatch) */
typedef char pslot_check_keypoint[(sizeof(pslot_keypoint) == 28) ? 1 : -1];
typedef char pslot_check_grid_spec[(sizeof(pslot_grid_spec) == 24) ? 1 : -1];
typedef char pslot_check_area_query[(sizeof(pslot_area_query) == 20) ? 1 : -1];
typedef char pslot_check_map_point[(sizeof(pslot_map_point) == 24) ? 1 : -1];
typedef char pslot_check_proj_point[(sizeof(pslot_proj_point) == 32) ? 1 : -1];

#ifdef __cplusplus
}
#endif
#endif /* PSLOT_ORB_H */
