// bow.cu — Frame::ComputeBoW (reference src/Frame.cc:2040-2047): mpORBvocabulary->transform(descriptors, mBowVec, mFeatVec, 4),
// i.e. DBoW2's TemplatedVocabulary<FORB::TDescriptor, FORB>::transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:
// 1223-1266 per feature): walk the vocabulary tree from the root, at every level take the FIRST child with the strictly
// smallest Hamming distance (FORB::distance, FORB.cpp:79-98), stop at a leaf; remember the node passed at level
// L - levelsup for the FeatureVector.
//
// One warp per feature.  The feature's 256 bits sit in registers; lane c loads child c's centre (two 16-byte loads,
// L2-resident: the ORB-SLAM vocabulary is 35 MB), XOR + popcount, and one __reduce_min_sync over (distance << 16 | child
// index) picks the child - ties go to the lower index, as the reference's `d < best_d` scan does.  Six dependent levels
// per feature, a few thousand features per frame: latency-bound by construction, a few microseconds per frame.
#include "pslot_internal.h"

namespace {

__global__ void __launch_bounds__(128)
k_bow_transform(const uint4 *__restrict__ node_desc, const int *__restrict__ child_start, const int *__restrict__ child,
                const uint4 *__restrict__ feat, int n, int nid_level, int *__restrict__ leaf_node, int *__restrict__ nid_out)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const uint4 f0 = __ldg(feat + 2 * (size_t)i), f1 = __ldg(feat + 2 * (size_t)i + 1);
    int node = 0, level = 0, at = nid_level <= 0 ? 0 : -1; // TemplatedVocabulary.h:1232-1233
    int cs = __ldg(child_start), ce = __ldg(child_start + 1);
    while (ce > cs) { // :1238-1261
        ++level;
        unsigned best = 0xFFFFFFFFu;
        for (int base = cs; base < ce; base += 32) { // more than 32 children: chunks, the earlier chunk wins ties
            const int c = base + lane;
            unsigned key = 0xFFFFFFFFu;
            if (c < ce) {
                const int id = __ldg(child + c);
                const uint4 a = __ldg(node_desc + 2 * (size_t)id), b = __ldg(node_desc + 2 * (size_t)id + 1);
                const int d = __popc(a.x ^ f0.x) + __popc(a.y ^ f0.y) + __popc(a.z ^ f0.z) + __popc(a.w ^ f0.w) +
                              __popc(b.x ^ f1.x) + __popc(b.y ^ f1.y) + __popc(b.z ^ f1.z) + __popc(b.w ^ f1.w);
                key = ((unsigned)d << 16) | (unsigned)(c - cs);
            }
            best = min(best, __reduce_min_sync(0xffffffffu, key));
        }
        node = __ldg(child + cs + (int)(best & 0xFFFFu));
        if (level == nid_level) at = node;
        cs = __ldg(child_start + node);
        ce = __ldg(child_start + node + 1);
    }
    if (lane == 0) {
        leaf_node[i] = node;
        nid_out[i] = at >= 0 ? at : node; // a leaf above level L - levelsup: the reference leaves *nid unset
    }
}

} // namespace

int launch_bow_transform(const uint8_t *node_desc, const int *child_start, const int *child, const uint8_t *feat, int n,
                         int nid_level, int *leaf_node, int *nid, cudaStream_t st)
{
    if (n <= 0) return 0;
    k_bow_transform<<<(n + 3) / 4, 128, 0, st>>>(reinterpret_cast<const uint4 *>(node_desc), child_start, child,
                                                  reinterpret_cast<const uint4 *>(feat), n, nid_level, leaf_node, nid);
    return 1;
}
