// pslot_internal.h — geometry tables and kernel launchers shared by the .cu files of libpslot_orb.so.
//
// Device data layout (all per "slot" = one image of a batch; a single extractor is a 1-slot engine):
//   pyramid     level-major: level l is one array [S][h_l][pitch_l] of u8 (pitch_l % 128 == 0), so the
//               level-0 upload of a whole batch is one 2-D copy and a (x, y, slot) tensor map per level
//               can address every tile.  No 19-px border is stored (never read).
//   cell_count  [S][cells]  FAST survivors per 30-px cell (all levels concatenated)
//   cell_cand   [S][cand_per_slot] u32  per-cell slots of packed candidates (x:12 | y:12 | resp:8,
//               window-relative coordinates), raster order inside a cell
//   lkp         [S][lkp_per_slot] u32   per-level keypoints chosen by the quadtree (level image coords)
//   lkp_count   [S][nlevels]
//   out_*       final keypoints / descriptors / counts, level-ascending
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <utility>

#include "../../include/pslot_orb.h"

#define PSLOT_MAX_LEVELS 16
#define PSLOT_EDGE 19       // EDGE_THRESHOLD, ORBextractor.cc:74
#define PSLOT_MINB 16       // minBorderX/Y = EDGE_THRESHOLD-3, ORBextractor.cc:773-774
#define PSLOT_HALF_PATCH 15 // HALF_PATCH_SIZE, ORBextractor.cc:73

struct LevelGeom {
    int w, h, pitch;                    // level image (ORBextractor.cc:1111-1112)
    int max_bx, max_by;                 // maxBorderX/Y = dim - 16 (:775-776)
    int n_cols, n_rows, w_cell, h_cell; // FAST cell grid (:784-787)
    int cell_base;                      // first cell of this level inside a slot
    int cand_stride;                    // candidate slots per cell (upper bound of strict 3x3 maxima)
    int cand_base;                      // u32 offset of this level's candidate slots inside a slot
    int quota;                          // mnFeaturesPerLevel[level] (:435-446)
    int kp_base, kp_cap;                // this level's segment of the lkp buffer
    int n_ini;                          // quadtree roots (:543)
    float hx;                           // root width (:545)
    float scale, inv_scale;             // mvScaleFactor / mvInvScaleFactor (:421,429)
    int patch_size;                     // (int)(31 * scale) (:834)
    int tab_x, tab_y;                   // offsets into the resize tap table
    int box_w, box_h;                   // TMA box of one FAST cell tile: align16(wCell + 30) x (hCell + 6)
};

struct Geom {
    int nlevels, ini_th, min_th;
    int cells_per_slot, cand_per_slot, lkp_per_slot, kp_cap_total;
    int max_level_cap; // max kp_cap over levels (sizes the quadtree node tables)
    LevelGeom lv[PSLOT_MAX_LEVELS];
};

struct PyrPtrs {
    uint8_t *lv[PSLOT_MAX_LEVELS];            // base of [S][h][pitch]
    unsigned long long slot_stride[PSLOT_MAX_LEVELS]; // h * pitch
};

// one bilinear tap of cv::resize INTER_LINEAR (8-bit): source indices and 11-bit coefficients
struct ResizeTap {
    uint16_t s0, s1;
    int16_t c0, c1;
};

// packed candidate / level keypoint
__host__ __device__ inline uint32_t pslot_pack(int x, int y, int r) { return ((uint32_t)y << 20) | ((uint32_t)x << 8) | (uint32_t)r; }
__host__ __device__ inline int pslot_px(uint32_t p) { return (int)((p >> 8) & 0xFFFu); }
__host__ __device__ inline int pslot_py(uint32_t p) { return (int)(p >> 20); }
__host__ __device__ inline int pslot_pr(uint32_t p) { return (int)(p & 0xFFu); }

// bits of the host-mapped error word a kernel can raise (Engine::h_err); decoded by check_err_flag in engine.cu
enum : int {
    PSLOT_FLAG_QT_OVERFLOW = 1,   // quadtree node table overflow (internal capacity bound violated)
    PSLOT_FLAG_STEREO_NR = 2,     // more right keypoints than the stereo kernel staged (would have been truncated)
    PSLOT_FLAG_STEREO_OCTAVE = 4, // a right keypoint's octave lies outside [0, nlevels)
    PSLOT_FLAG_PYR_TILE = 8,      // a resize tile's source rectangle exceeds the staging buffer (level left unwritten)
    PSLOT_FLAG_CVORB_CAP = 16,    // cv::ORB path: a candidate list or the on-chip sort of one level overflowed
};

struct StereoArgs {
    // left / right feature sets: per pair p, keypoints at kps + slot*K
    const pslot_keypoint *kps_l, *kps_r;
    const uint8_t *desc_l, *desc_r;
    const int *n_l, *n_r; // per-pair counts
    int K;                // row stride (keypoints) between pairs
    PyrPtrs pyr_l, pyr_r; // pyramids; pair p is slot p of each
    float bf, max_d;
    float *u_right, *depth;          // [pairs][K]
    int *best_dist, *best_idx, *sad; // [pairs][K]
    int *err_flag;                   // host-mapped error word (PSLOT_FLAG_*), may be null
    unsigned long long *counters;    // optional work counters: [0] row-band candidates scanned (sum over left keypoints of
                                     // |vRowIndices[row]|), [1] descriptors compared; null = not counted
};

// ---- launchers (each enqueues on `st`; returns the number of kernels launched) -----------------
struct PyrMaps;
int launch_pyramid(const Geom &g, const PyrPtrs &pyr, const PyrMaps &maps, const ResizeTap *taps, int nslots, int *err_flag,
                   cudaStream_t st);
// source box (bytes x rows) that covers every 128 x 64 output tile of a level's resize; the x origin of a tile's box is
// rounded down to 16 bytes
void pyramid_src_box(const ResizeTap *xt, int dw, const ResizeTap *yt, int dh, int *box_w, int *box_h);
#ifndef PYR_TH
#define PYR_TH 64
#endif
constexpr int PYR_TILE_W = 128, PYR_TILE_H = PYR_TH;
// true when every 128 x 64 output tile of the level's resize fits the kernel's staging buffer (checked on the host so
// that the device-side guard can never fire silently)
bool pyramid_tiles_fit(const ResizeTap *xt, int dw, const ResizeTap *yt, int dh);
struct TmapSet;
// One FAST cell of a slot, precomputed on the host (ORBextractor.cc:789-806): everything k_fast_cells needs, in one
// 16-byte record that a warp fetches with a single load instead of deriving it from the level tables.
struct FastCell {
    uint16_t ini_x, ini_y;  // cell sub-image origin (level coordinates)
    uint8_t tw, th;         // sub-image size incl. the 3-px FAST margin; tw == 0: the cell tests no pixel
    uint8_t level, box_h;   // pyramid level; rows of the TMA box
    uint32_t cand_off;      // first candidate slot of the cell inside a slot's candidate array
    uint16_t cand_stride;   // candidate slots of the cell
    uint16_t box_w;         // TMA box width = shared-memory row stride (bytes, multiple of 16)
};
static_assert(sizeof(FastCell) == 16, "FastCell is loaded as one uint4");
void make_fast_cells(const Geom &g, FastCell *out); // out[g.cells_per_slot]
int launch_fast(const Geom &g, const TmapSet &maps, const FastCell *cells, int *cell_count, uint32_t *cell_cand, int nslots,
                int *work_counter, cudaStream_t st);
int launch_quadtree(const Geom &g, const int *cell_count, const uint32_t *cell_cand, int *cell_prefix,
                    uint32_t *qt_keys, uint16_t *qt_nodes, uint32_t *lkp, int *lkp_count, int *err_flag,
                    int nslots, cudaStream_t st, cudaStream_t aux = nullptr, cudaEvent_t ev_fork = nullptr,
                    cudaEvent_t ev_join = nullptr);
int launch_describe(const Geom &g, const PyrPtrs &pyr, const uint32_t *lkp, const int *lkp_count,
                    pslot_keypoint *out_kps, uint8_t *out_desc, int *out_count, int nslots, cudaStream_t st);
int launch_stereo(const Geom &g, const StereoArgs &a, int npairs, int max_n_r, cudaStream_t st,
                  cudaEvent_t between = nullptr);
size_t quadtree_smem_bytes(const Geom &g, int *key_cap);
int launch_gray_ingest(const uint8_t *src, size_t src_image_stride, int channels, int rgb, uint8_t *dst, int w, int h,
                       int dpitch, unsigned long long dstride, int nimages, cudaStream_t st);
// Opt a kernel in to `bytes` of dynamic shared memory.  The attribute is per function and per device, and contexts with
// different geometries launch from different host threads (Frame drives its two extractors concurrently), so it is
// only ever RAISED, under a lock: setting the exact size before each launch would let one thread lower it between
// another thread's set and launch.
template <typename K> inline void ensure_dyn_smem(K kernel, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> high;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    size_t &h = high[std::make_pair((const void *)kernel, dev)];
    if (bytes > h) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        h = bytes;
    }
}

// ---- cv::ORB (OpenCV's own extractor, cvorb.cu) -------------------------------------------------------------------
struct CvLevel {
    int w, h, pitch;         // cvRound(dim / scale), rows padded to 128 bytes
    unsigned long long off;  // byte offset of the level inside the pyramid / mask / score buffers
    int quota;               // nfeatures per level (orb.cpp computeKeyPoints)
    float scale;             // (float)pow(scaleFactor, level)
    int tab_x, tab_y;        // INTER_LINEAR_EXACT tap tables (source index, 8-bit fraction)
    int cand_base, cand_cap; // segment of the candidate / kept arrays
    int kp_base;             // segment of the final per-level keypoint lists (CV_SORT_CAP entries each)
    int tile_base, tiles_x;  // 128 px x 32 row tiles of the scored band (rows from edge - 1), numbered level after level
};
struct CvGeom {
    int nlevels, edge, fast_th;
    float gk[4];             // getGaussianKernel(7, 2, CV_32F)[0..3]
    CvLevel lv[PSLOT_MAX_LEVELS];
    // batch form: slot s (one image of the call) owns [s * stride, (s + 1) * stride) of every per-image buffer
    unsigned long long slot_bytes; // pyramid / mask / score
    int slot_cand;                 // cand / kept / kept_resp entries
    int slot_kp;                   // final_kp / final_resp / out_kps / out_desc entries (CV_SORT_CAP per level)
    int describe_blocks;           // grid.x of k_cv_describe (grid-stride over a level's keypoints)
    int n_tiles;                   // sum of the levels' tiles: grid.x of the score / maxima kernels
};
constexpr int CV_SORT_CAP = 4096; // keypoints one level can keep after retainBest(quota) incl. ties (on-chip sort)
enum : int { CV_CNT_CAND = 0, CV_CNT_KEPT = PSLOT_MAX_LEVELS, CV_CNT_FINAL = 2 * PSLOT_MAX_LEVELS, CV_CNT_TOTAL = 3 * PSLOT_MAX_LEVELS };
struct CvBuffers {
    uint8_t *pyr, *mask, *score;
    const int2 *taps;
    uint32_t *cand, *kept, *final_kp;
    float *kept_resp, *final_resp;
    int *counters, *hist, *err;
    pslot_keypoint *out_kps;
    uint8_t *out_desc;
    int *out_count;
};
int launch_cvorb(const CvGeom &g, const CvBuffers &b, int slots, bool has_mask, cudaStream_t st);

int launch_bow_transform(const uint8_t *node_desc, const int *child_start, const int *child, const uint8_t *feat, int n,
                         int nid_level, int *leaf_node, int *nid, cudaStream_t st);

typedef pslot_grid_spec GridSpec;
size_t grid_smem_bytes(const GridSpec &s);
int launch_assign_grid(const pslot_keypoint *kps, const int *counts, int fixed_n, size_t stride, const GridSpec &s,
                       int *cell_start, int *indices, int nimages, cudaStream_t st);
int launch_features_in_area(const pslot_keypoint *kps, const int *cell_start, const int *indices, const GridSpec &s,
                            const pslot_area_query *queries, int nq, int cap, int *out_count, int *out, cudaStream_t st);
// one projected map point as the search kernel sees it: the window the reference passes to GetFeaturesInArea and the
// right-image test, already evaluated by the caller
struct ProjPoint {
    float x, y, xr, radius, er_limit;
    int min_level, max_level;
    int flags; // bit 0: searched; bit 1: Observations() > 0
};
struct ProjArgs {
    const pslot_keypoint *kps;      // mvKeysUn of the gridded frame
    const int *cell_start, *indices;
    GridSpec s;
    const ProjPoint *pts;
    const uint8_t *pt_desc;
    int n_pts;
    const uint8_t *frame_desc;
    const float *u_right;
    const uint8_t *occupied;
    const int *claim_prev;
    int *claim_next;
    const int *match_prev;
    int *match_next;
    int *changed;
    int mode;       // 0: local map (best + second best, ratio test); 1: last frame / key frame (best only)
    float nn_ratio;
    int th_dist;    // acceptance: bestDist <= th_dist (TH_HIGH = 100, or the relocalisation search's ORBdist)
};
int launch_search_by_projection(const ProjArgs &a, cudaStream_t st);
// Frame::AssignFeatures (partition.cu): classify by the instance mask + stable partition, one CTA per image
struct PartArgs {
    const pslot_keypoint *kps; const int *counts; int fixed_n; size_t kp_stride; // image i: kps + i * kp_stride, counts[i] or fixed_n features
    const uint8_t *mask; int w, h, pitch; size_t mask_stride;                   // instance mask (0 background, 255 ignored, id + 1)
    const int *object_ids; int n_objects; size_t ids_stride;                    // mnObjectID per detection (negative: null); stride 0 = shared
    const float *depth; size_t depth_stride; float depth_limit;                 // optional: effective[j] counts 0 < depth < limit
    int *cls, *order; size_t out_stride;
    int *class_start;                                                           // n_objects + 2 per image
    int *effective;                                                             // n_objects per image, or null
};
int launch_assign_features(const PartArgs &a, int nimages, cudaStream_t st);
void launch_debug_steer(const float *angle_deg, int n, float *a, float *b, cudaStream_t st);
