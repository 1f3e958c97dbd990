// partition.cu — Frame::AssignFeatures (reference src/Frame.cc:759-977): the step of the object-mode Frame constructors
// that follows ComputeStereoMatches / ComputeObjStereoMatches (:725-733).  It looks every feature up in the instance mask
// (mMaskImg: 0 = background, 255 = ignored, 1..254 = object id + 1) and re-partitions the frame's arrays:
//     label 0                         -> the background set (mvKeys / mDescriptors / mvuRight / mvDepth / mvKeysUn afterwards)
//     label L in 1..254               -> the set of the FIRST non-null detection j whose id equals L - 1, where
//                                        id = mnObjectID > 255 ? mnObjectID - 255 : mnObjectID (:836-846); none: dropped
//     label 255                       -> dropped
// each set in the features' original order (push_back in a forward loop).  The same routine runs on the frame's own ORB
// features (-> mvOriKeys[j] ...) and on the dense cv::ORB object features (-> mvObjKeys[j] ..., plus the count of
// features with 0 < depth < 2 * mThDepth, :893-896), so it is one classify + stable-partition primitive:
//     cls[i]   = -1 dropped | 0 background | 1 + j
//     order[]  = feature indices grouped by class (background first, then detection 0, 1, ...), original order inside
//     class_start[c] .. class_start[c + 1] = the slice of order[] that holds class c  (n_objects + 2 entries)
// Integer / index work only, hence bit-exact.  One CTA per image (a frame has a few thousand features): label table and
// class histogram in shared memory, then one warp scatters in groups of 32 consecutive features (match.any ranks inside a
// group, running offsets between groups), which reproduces the push_back order without sorting.  Latency-bound.
#include <climits>

#include "pslot_internal.h"

namespace {

constexpr int PT = 256;          // threads per CTA
constexpr int MAX_CLASSES = 258; // background + up to 256 detections (+ 1 for the closing offset)

__global__ void __launch_bounds__(PT)
k_assign_features(const PartArgs a)
{
    __shared__ int s_tab[256];         // mask label -> first matching detection (INT_MAX: none)
    __shared__ int s_cnt[MAX_CLASSES]; // class counts, then running offsets
    __shared__ int s_eff[256];         // per detection: features with 0 < depth < depth_limit
    const int img = blockIdx.x, tid = threadIdx.x;
    const int n = a.counts ? a.counts[img] : a.fixed_n;
    const pslot_keypoint *kps = a.kps + (size_t)img * a.kp_stride;
    const uint8_t *mask = a.mask + (size_t)img * a.mask_stride;
    const int *ids = a.object_ids + (size_t)img * a.ids_stride;
    const float *depth = a.depth ? a.depth + (size_t)img * a.depth_stride : nullptr;
    int *cls = a.cls + (size_t)img * a.out_stride, *order = a.order + (size_t)img * a.out_stride;
    int *class_start = a.class_start + (size_t)img * (a.n_objects + 2);
    int *effective = a.effective ? a.effective + (size_t)img * a.n_objects : nullptr;

    s_tab[tid] = INT_MAX;
    s_eff[tid] = 0;
    for (int c = tid; c < MAX_CLASSES; c += PT) s_cnt[c] = 0;
    __syncthreads();
    for (int j = tid; j < a.n_objects; j += PT) {
        const int oid = ids[j]; // negative: a null detection (vDetectionObjects[j] == NULL, :839) - matches no label
        const int id = oid > 255 ? oid - 255 : oid;
        if (id >= 0 && id <= 253) atomicMin(&s_tab[id + 1], j); // the reference's inner loop stops at the first match
    }
    __syncthreads();
    for (int i = tid; i < n; i += PT) {
        const pslot_keypoint k = kps[i];
        const int x = (int)k.x, y = (int)k.y; // int x = mvKeys[i].pt.x (:816-817): truncation
        int c = -1;
        if (x >= 0 && y >= 0 && x < a.w && y < a.h) { // the reference reads the mask unchecked; keypoints lie inside the image
            const int label = mask[(size_t)y * a.pitch + x];
            if (label == 0) c = 0;
            else if (label != 255 && s_tab[label] != INT_MAX) c = 1 + s_tab[label];
        }
        cls[i] = c;
        if (c >= 0) atomicAdd(&s_cnt[c], 1);
        if (c >= 1 && depth) {
            const float d = depth[i];
            if (d > 0.f && d < a.depth_limit) atomicAdd(&s_eff[c - 1], 1);
        }
    }
    __syncthreads();
    if (tid < 32) {
        // exclusive scan over the classes by one warp, 32 at a time
        const int nc = a.n_objects + 1;
        int carry = 0;
        for (int base = 0; base < nc; base += 32) {
            const int c = base + tid;
            const int v = c < nc ? s_cnt[c] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (tid >= o) inc += t;
            }
            if (c < nc) {
                s_cnt[c] = carry + inc - v;
                class_start[c] = carry + inc - v;
            }
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (tid == 0) class_start[nc] = carry;
        __syncwarp();
        // stable scatter: groups of 32 consecutive features; inside a group the rank among equal classes comes from
        // match.any, between groups from the running offsets
        const unsigned lt = (1u << tid) - 1u;
        for (int base = 0; base < n; base += 32) {
            const int i = base + tid;
            const int c = i < n ? cls[i] : -2;
            const unsigned m = __match_any_sync(0xffffffffu, c);
            if (c >= 0) {
                const int off = s_cnt[c];
                order[off + __popc(m & lt)] = i;
            }
            __syncwarp();
            if (c >= 0 && (m & lt) == 0) s_cnt[c] += __popc(m); // the group's first lane of this class
            __syncwarp();
        }
    }
    if (effective)
        for (int j = tid; j < a.n_objects; j += PT) effective[j] = s_eff[j];
}

} // namespace

int launch_assign_features(const PartArgs &a, int nimages, cudaStream_t st)
{
    k_assign_features<<<nimages, PT, 0, st>>>(a);
    return 1;
}
