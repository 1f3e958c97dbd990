/*
 * orb_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of PointSLOT's per-frame stereo ORB front end, used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker
 * and the timed CPU baseline.  The shipped library (libpslot_orb.so) never links or calls it.
 *
 * PARITY STATUS: pinned against the reference's own code compiled in this container (oracle/_ref,
 * `make -C oracle ref`, tests/test_oracle_reference.py): byte-identical extractor output and bit-equal
 * stereo results versus the reference binary built with a monotonic allocator; versus the stock-malloc
 * build the only delta is the heap-address tie of ORBextractor.cc:684 (0.84 % of the keypoint set,
 * profiles/r02_ref_delta.txt).  The primitives restated here (resize INTER_LINEAR, FAST 9-16 + NMS,
 * GaussianBlur 7x7 s=2, fastAtan2, cvRound) are pinned against OpenCV 4.13.0 through cv2 in
 * tests/test_oracle_cv2.py, cosf/sinf against glibc 2.39.  Explicit rules replace the reference's two
 * UB / address-dependent spots (DESIGN.md "Parity rules"): quadtree ties -> later-created node first;
 * maxD explicit.
 *
 * Reference files restated (paths under /root/reference):
 *   src/ORBextractor.cc:77-147, 410-853, 1034-1132   include/ORBextractor.h
 *   src/Frame.cc:2142-2316                           src/ORBmatcher.cc:2704-2720
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* layout-identical to cv::KeyPoint (28 bytes) */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orc_keypoint;

typedef struct orc_extractor orc_extractor;

orc_extractor *orc_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th);
void orc_destroy(orc_extractor *e);

/* tables computed by the constructor (ORBextractor.cc:410-470) */
const float *orc_scale_factors(const orc_extractor *e);
const float *orc_inv_scale_factors(const orc_extractor *e);
const float *orc_level_sigma2(const orc_extractor *e);
const float *orc_inv_level_sigma2(const orc_extractor *e);
const int *orc_features_per_level(const orc_extractor *e);
const int *orc_umax(const orc_extractor *e);

/* ORBextractor::operator() (ORBextractor.cc:1043-1105). Returns 0, or -1 if cap is too small
 * (n is still set to the required count).  Empty image (w<=0||h<=0) -> n=0, return 0. */
int orc_extract(orc_extractor *e, const uint8_t *img, int w, int h, ptrdiff_t pitch,
                orc_keypoint *kps, uint8_t *desc32, int cap, int *n);

/* stage intermediates of the most recent orc_extract */
int orc_level_size(const orc_extractor *e, int level, int *w, int *h);
const uint8_t *orc_level_ptr(const orc_extractor *e, int level);   /* borderless, pitch = w */
const uint8_t *orc_blurred_ptr(const orc_extractor *e, int level); /* NULL if level had no keypoints */
/* FAST candidates handed to DistributeOctTree, window-relative coords, reference order */
int orc_level_candidates(const orc_extractor *e, int level, int cap, int *x, int *y, int *resp);
/* keypoints after DistributeOctTree + border offset, level coordinates, list order */
int orc_level_keypoints(const orc_extractor *e, int level, int cap, int *x, int *y, int *resp, float *angle);
/* Allocator (heap-address) sensitivity of DistributeOctTree's fine phase, ORBextractor.cc:684 (rule C.1):
 * bit 0: the cut of :731 came at a node that belongs to a run of equal-size nodes (the retained SET depends on the tie rule);
 * bit 1: two equal-size nodes were both split in one run (the list ORDER depends on the tie rule).
 * Sound against the stock-allocator build of the reference: no bit 0 -> same set, no flag -> byte-identical level. */
int orc_level_tie_sensitive(const orc_extractor *e, int level);

/* Frame::ComputeStereoMatches (Frame.cc:2142-2316) on the pyramids held by eL / eR.
 * uright/depth: nL floats (-1 = no match).  Optional debug arrays (may be NULL), nL entries:
 *   best_dist (Hamming of the chosen candidate, 100 if none), best_idx (right index or -1),
 *   best_sad (SAD of accepted matches before outlier rejection, -1 otherwise). */
int orc_stereo_match(const orc_extractor *eL, const orc_extractor *eR,
                     const orc_keypoint *kL, int nL, const uint8_t *dL,
                     const orc_keypoint *kR, int nR, const uint8_t *dR,
                     float bf, float maxD, float *uright, float *depth,
                     int *best_dist, int *best_idx, int *best_sad);

/* ORBmatcher::DescriptorDistance (ORBmatcher.cc:2704-2720) */
int orc_hamming256(const void *a, const void *b);

/* restated OpenCV primitives, exported for the known-answer tests */
void orc_resize_linear_u8(const uint8_t *src, int sw, int sh, ptrdiff_t spitch,
                          uint8_t *dst, int dw, int dh, ptrdiff_t dpitch);
int orc_fast9_16(const uint8_t *img, int w, int h, ptrdiff_t pitch, int threshold,
                 int cap, int *x, int *y, int *resp);
void orc_gaussian7x7_s2(const uint8_t *src, int w, int h, ptrdiff_t spitch, uint8_t *dst, ptrdiff_t dpitch);
float orc_fast_atan2(float y, float x);
int orc_cv_round(double v);
/* cv::cvtColor(RGB/BGR(A) -> GRAY) for 8-bit images as OpenCV 4.13 computes it (called at Tracking.cc:1014-1039) */
void orc_cvt_gray(const uint8_t *src, int w, int h, ptrdiff_t spitch, int channels, int rgb_order, uint8_t *dst, ptrdiff_t dpitch);
/* glibc >= 2.28 cosf / sinf (the functions `(float)cos(float)` resolves to in ORBextractor.cc:112-113), |x| < 120 */
float orc_glibc_cosf(float x);
float orc_glibc_sinf(float x);
void orc_glibc_sincosf_n(const float *x, int n, float *c, float *s);
/* steering coefficients of computeOrbDescriptor (ORBextractor.cc:112-113) for n angles in degrees */
void orc_steer_coeffs(const float *angle_deg, int n, float *a, float *b);
/* DistributeOctTree alone (ORBextractor.cc:539-763); in/out arrays of n / cap ints */
int orc_distribute_octtree(const int *x, const int *y, const int *resp, int n,
                           int minX, int maxX, int minY, int maxY, int N,
                           int cap, int *ox, int *oy, int *oresp, int *tie_sensitive);

/* Frame::AssignFeaturesToGrid + PosInGrid (Frame.cc:1636-1655, 2027-2038): mGrid in CSR form, cell = i * rows + j;
 * returns the number of keypoints that fell inside the grid (-1: bad arguments) */
int orc_assign_grid(const orc_keypoint *kps, int n, float min_x, float min_y, float inv_w, float inv_h, int cols,
                    int rows, int *cell_start, int *indices);
/* Frame::AssignFeatures (Frame.cc:759-977), its classification + push_back loops: feature i -> mask((int)pt.y, (int)pt.x);
 * 0 -> background, 255 -> dropped, L -> the first detection j (object_ids[j] >= 0) with
 * (object_ids[j] > 255 ? object_ids[j] - 255 : object_ids[j]) == L - 1 (:836-846), else dropped.  cls[i] = -1 | 0 | 1 + j;
 * order = indices grouped by class in push_back order; class_start[n_objects + 2]; effective[j] (depth != NULL) = features
 * of detection j with 0 < depth < depth_limit (:893-896).  Returns the number of kept features (-1: bad arguments). */
int orc_assign_features(const orc_keypoint *kps, int n, const uint8_t *mask, int w, int h, int pitch, const int *object_ids,
                        int n_objects, const float *depth, float depth_limit, int *cls, int *order, int *class_start,
                        int *effective);
/* Frame::GetFeaturesInArea (Frame.cc:1808-1861) on that CSR grid; returns the number found (may exceed cap) */
int orc_features_in_area(const orc_keypoint *kps, const int *cell_start, const int *indices, float min_x, float min_y,
                         float inv_w, float inv_h, int cols, int rows, float x, float y, float r, int min_level,
                         int max_level, int cap, int *out);

/* the MapPoint fields ORBmatcher::SearchByProjection reads (MapPoint.h:110-120) */
typedef struct orc_map_point {
    float proj_x, proj_y, proj_xr, view_cos; /* mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos */
    int32_t level;                           /* mnTrackScaleLevel */
    int32_t flags;                           /* bit 0: mbTrackInView && !isBad(); bit 1: Observations() > 0 */
} orc_map_point;
/* ORBmatcher::SearchByProjection(F, vpMapPoints, th) (ORBmatcher.cc:68-155) on the CSR grid of orc_assign_grid;
 * occupied[idx] != 0 where F.mvpMapPoints[idx] already holds a point with observations; match_of_mp[i] = bestIdx
 * of map point i or -1; returns nmatches */
int orc_search_by_projection(const orc_keypoint *kps_un, const uint8_t *frame_desc, const float *u_right, int n,
                             const int *cell_start, const int *indices, float min_x, float min_y, float inv_w,
                             float inv_h, int cols, int rows, const orc_map_point *mps, const uint8_t *mp_desc,
                             int n_mp, const uint8_t *occupied, const float *scale_factors, float th, float nn_ratio,
                             int *match_of_mp);

/* the object-point overload (ORBmatcher.cc:157-246) on one detected object's feature set: fixed 5 px window, levels
 * [L - 1, L + 1], TH_HIGH_FORDYNAMIC; `occupied` also carries !isInBBox(keypoint) */
int orc_search_by_projection_obj(const orc_keypoint *kps_un, const uint8_t *frame_desc, const float *u_right, int n,
                                 const int *cell_start, const int *indices, float min_x, float min_y, float inv_w,
                                 float inv_h, int cols, int rows, const orc_map_point *mps, const uint8_t *mp_desc,
                                 int n_mp, const uint8_t *occupied, const float *scale_factors, float th, float nn_ratio,
                                 int *match_of_mp);

/* one LastFrame map point after the reference's projection lines (ORBmatcher.cc:1640-1676) */
typedef struct orc_proj_point {
    float u, v, ur, radius;       /* projection, u - mbf * invzc, th * mvScaleFactors[nLastOctave] */
    int32_t min_level, max_level; /* level arguments of the GetFeaturesInArea call taken at :1671-1676 */
    float angle;                  /* LastFrame.mvKeysUn[i].angle */
    int32_t flags;                /* bit 0: reaches the search; bit 1: Observations() > 0 */
} orc_proj_point;
/* the rest of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1671-1756);
 * rejected[i] = 1 where the rotation check nulled the match again; returns nmatches (-1: angle out of range) */
int orc_search_by_projection_last(const orc_keypoint *kps_un, const uint8_t *frame_desc, const float *u_right, int n,
                                  const int *cell_start, const int *indices, float min_x, float min_y, float inv_w,
                                  float inv_h, int cols, int rows, const orc_proj_point *pts, const uint8_t *pt_desc,
                                  int n_pts, const uint8_t *occupied, int check_orientation, int *match_of_pt,
                                  uint8_t *rejected);

/* the rest of ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:2583-2654,
 * relocalisation): no stereo check, any association blocks a feature, acceptance at orb_dist, rotation bins of 30 degrees;
 * pts[i].ur is ignored */
int orc_search_by_projection_kf(const orc_keypoint *kps_un, const uint8_t *frame_desc, int n, const int *cell_start,
                                const int *indices, float min_x, float min_y, float inv_w, float inv_h, int cols, int rows,
                                const orc_proj_point *pts, const uint8_t *pt_desc, int n_pts, const uint8_t *occupied,
                                int orb_dist, int check_orientation, int *match_of_pt, uint8_t *rejected);

/* cv::ORB::create(nfeatures, scaleFactor, nlevels, edgeThreshold)->detectAndCompute(image, mask) (reference
 * src/Frame.cc:2623-2627; OpenCV 4.13 orb.cpp), HARRIS score, patch 31, WTA_K 2; mask may be NULL; output ordered by
 * (octave, y, x); returns 0, -1 (bad arguments) or -2 (cap too small) */
int orc_cvorb_detect_and_compute(const uint8_t *img, int w, int h, ptrdiff_t pitch, const uint8_t *mask,
                                 ptrdiff_t mask_pitch, int nfeatures, float scale_factor, int nlevels,
                                 int edge_threshold, int fast_threshold, orc_keypoint *kps, uint8_t *desc32, int cap,
                                 int *n_out);

/* Frame::ComputeBoW (reference src/Frame.cc:2040-2047) = ORBVocabulary::transform(descriptors, mBowVec, mFeatVec, 4), i.e.
 * DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>::transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1132-1201,
 * 1223-1266; un-vendored algorithm = the vendored copy in the reference tree).  The vocabulary is a flat tree: node 0 is
 * the root, children of node i are child[child_start[i] .. child_start[i + 1]) in file order, a node without children is
 * a leaf and carries word_id >= 0 and its weight. */
typedef struct orc_vocab {
    int32_t n_nodes, k, L, scoring, weighting; /* scoring: 0 L1 1 L2 2 CHI 3 KL 4 BHATT 5 DOT; weighting: 0 TF_IDF 1 TF 2 IDF 3 BINARY */
    const uint8_t *desc;        /* [n_nodes][32] */
    const int32_t *child_start; /* [n_nodes + 1] */
    const int32_t *child;       /* node ids */
    const int32_t *word_id;     /* [n_nodes], -1 for inner nodes */
    const double *weight;       /* [n_nodes] */
} orc_vocab;
/* per feature: the leaf reached (TemplatedVocabulary.h:1223-1266: first child with the strictly smallest Hamming distance
 * at every level) and the node at level L - levelsup on the way (0 when that level is <= 0; the leaf itself when the walk
 * ends above that level, where the reference leaves *nid unset) */
void orc_bow_transform(const orc_vocab *v, const uint8_t *desc, int n, int levelsup, int32_t *leaf_node, int32_t *nid);
/* the two containers of transform(features, v, fv, levelsup) (:1132-1201) from the per-feature results, in the reference's
 * arithmetic: BowVector = (word id ascending, value) with addWeight / addIfNotExist accumulation in feature order and the
 * scoring's normalisation; FeatureVector = (node id ascending -> feature indices in feature order) as CSR.
 * Returns the number of words; *n_fv_nodes = FeatureVector size.  Arrays must hold n entries (fv_start n + 1). */
int orc_bow_vectors(const orc_vocab *v, const int32_t *leaf_node, const int32_t *nid, int n, int32_t *bow_id, double *bow_val,
                    int32_t *fv_node, int32_t *fv_start, int32_t *fv_feat, int *n_fv_nodes);

#ifdef __cplusplus
}
#endif
#endif
