// grid.cu — the 64 x 48 feature grid of Frame: Frame::PosInGrid (reference
// src/Frame.cc:2027-2038), Frame::AssignFeaturesToGrid (:1636-1655) and the lookup that consumes it,
// Frame::GetFeaturesInArea (:1808-1861).
//
// mGrid[i][j] (a vector of keypoint indices per cell, filled by push_back in keypoint order) becomes a CSR table
// on the device: cell = i * rows + j, cell_start[cols * rows + 1], indices[] ascending inside a cell.
//   k_assign_grid      one CTA per image: shared-memory histogram of the cell of every keypoint, block scan, then
//                      a chunk-ordered scatter with shared atomics and one short insertion pass per cell, which
//                      leaves each cell's list in ascending keypoint order; deterministic.  (A single-warp stable
//                      scatter with match.any ranks was 3x slower: match.any serialises over the distinct cells.)
//   k_features_in_area one warp per query: the cell rectangle is walked in the reference's order (ix outer, iy
//                      inner, list order inside a cell) and the hits are compacted with ballots, so the output
//                      order equals the reference's vIndices.
//   k_search_by_projection  ORBmatcher::SearchByProjection (local-map and last-frame variants)
//                      on top of the same window walk; see the kernel's comment for how the reference's order
//                      dependence is kept.
// Float arithmetic as the reference writes it, products and differences kept separate (no FMA).
// Bound: latency / HBM (28 B read + 4 B written per keypoint); algorithmic bytes = n_k * 32 + 4 * (cols * rows + 1).
#include "pslot_internal.h"

namespace {

__device__ __forceinline__ int pos_in_grid(const pslot_keypoint &kp, const GridSpec &s)
{
    // posX = round((kp.pt.x - mnMinX) * mfGridElementWidthInv) (:2029-2030); C round() of a float value == roundf
    const float rx = roundf(__fmul_rn(__fsub_rn(kp.x, s.min_x), s.inv_w));
    const float ry = roundf(__fmul_rn(__fsub_rn(kp.y, s.min_y), s.inv_h));
    if (!(rx >= 0.0f && rx < (float)s.cols && ry >= 0.0f && ry < (float)s.rows)) return -1; // (:2033)
    return (int)rx * s.rows + (int)ry;
}

// cell rectangle of a search window (:1813-1827); the float -> int conversions are clamped first (the reference's are
// undefined out of range)
struct AreaCells { int x0, x1, y0, y1; bool ok; };
__device__ __forceinline__ AreaCells area_cells(const GridSpec &s, float x, float y, float r)
{
    const float big = 1.0e9f;
    const float fx0 = floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, s.min_x), r), s.inv_w));
    const float fx1 = ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, s.min_x), r), s.inv_w));
    const float fy0 = floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, s.min_y), r), s.inv_h));
    const float fy1 = ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, s.min_y), r), s.inv_h));
    AreaCells A;
    A.x0 = max(0, (int)fminf(fmaxf(fx0, -big), big));
    A.x1 = min(s.cols - 1, (int)fminf(fmaxf(fx1, -big), big));
    A.y0 = max(0, (int)fminf(fmaxf(fy0, -big), big));
    A.y1 = min(s.rows - 1, (int)fminf(fmaxf(fy1, -big), big));
    A.ok = A.x0 < s.cols && A.x1 >= 0 && A.y0 < s.rows && A.y1 >= 0 && fx0 == fx0 && fx1 == fx1 && fy0 == fy0 && fy1 == fy1;
    return A;
}

constexpr int GT = 256;

__global__ void __launch_bounds__(GT)
k_assign_grid(const pslot_keypoint *__restrict__ kps, const int *__restrict__ counts, int fixed_n, size_t stride,
              GridSpec s, int *__restrict__ cell_start, int *__restrict__ indices, int cache_cap)
{
    extern __shared__ int sm[];
    const int C = s.cols * s.rows;
    int *start = sm;          // [C + 1]
    int *fill = sm + C + 1;   // [C]
    // cell of every keypoint (0xFFFF = outside), computed once
    uint16_t *cell_of = reinterpret_cast<uint16_t *>(fill + C); // [cache_cap]; C <= 16384 fits 16 bits
    __shared__ int warp_sums[GT / 32];

    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const pslot_keypoint *k = kps + (size_t)img * stride;
    const int n = counts ? min(counts[img], (int)stride) : fixed_n;
    int *cs = cell_start + (size_t)img * (C + 1);
    int *idx = indices + (size_t)img * stride;

    for (int c = tid; c <= C; c += GT) start[c] = 0;
    __syncthreads();
    const bool cached = n <= cache_cap;
    for (int i = tid; i < n; i += GT) {
        const int c = pos_in_grid(k[i], s);
        if (c >= 0) atomicAdd(&start[c + 1], 1);
        if (cached) cell_of[i] = (uint16_t)c; // -1 -> 0xFFFF
    }
    __syncthreads();
    // inclusive scan of start[1..C] in place: a contiguous chunk per thread, then a scan of the chunk sums
    const int per = (C + GT - 1) / GT;
    const int lo = 1 + tid * per, hi = min(lo + per, C + 1);
    int sum = 0;
    for (int c = lo; c < hi; ++c) sum += start[c];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int offset = incl - sum;
    for (int w = 0; w < warp; ++w) offset += warp_sums[w];
    for (int c = lo; c < hi; ++c) {
        offset += start[c];
        start[c] = offset;
    }
    __syncthreads();
    for (int c = tid; c <= C; c += GT) {
        cs[c] = start[c];
        if (c < C) fill[c] = start[c];
    }
    __syncthreads();
    // Scatter in chunks of GT consecutive keypoints with a barrier between chunks: the shared atomics may permute the
    // entries of one chunk inside a cell, but chunks land in index order, so every list is sorted up to swaps among the
    // few same-chunk neighbours ...
    for (int base = 0; base < n; base += GT) {
        const int i = base + tid;
        if (i < n) {
            int c;
            if (cached) { const int v = cell_of[i]; c = v == 0xFFFF ? -1 : v; }
            else c = pos_in_grid(k[i], s);
            if (c >= 0) idx[atomicAdd(&fill[c], 1)] = i;
        }
        __syncthreads(); // also makes the block's global writes visible to it
    }
    // ... which one insertion pass per cell (a thread per cell; lists are a handful of entries, displacements bounded by
    // the same-chunk population) turns into the ascending = push_back order of the reference (:1647-1654).
    for (int c = tid; c < C; c += GT) {
        const int a = start[c], b = start[c + 1];
        for (int i = a + 1; i < b; ++i) {
            const int v = idx[i];
            int j = i - 1;
            while (j >= a && idx[j] > v) { idx[j + 1] = idx[j]; --j; }
            idx[j + 1] = v;
        }
    }
}

__global__ void __launch_bounds__(256)
k_features_in_area(const pslot_keypoint *__restrict__ kps, const int *__restrict__ cell_start,
                   const int *__restrict__ indices, GridSpec s, const pslot_area_query *__restrict__ queries, int nq,
                   int cap, int *__restrict__ out_count, int *__restrict__ out)
{
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nq) return;
    const pslot_area_query Q = queries[q];
    const float x = Q.x, y = Q.y, r = Q.r;
    int n = 0;
    const AreaCells A = area_cells(s, x, y, r);
    const int nMinCellX = A.x0, nMaxCellX = A.x1, nMinCellY = A.y0, nMaxCellY = A.y1;
    const bool ok = A.ok;
    if (ok) {
        const bool check_levels = Q.min_level > 0 || Q.max_level >= 0;
        int *o = out + (size_t)q * cap;
        for (int ix = nMinCellX; ix <= nMaxCellX; ++ix) {
            // the cells (ix, nMinCellY..nMaxCellY) are adjacent in the CSR table: one contiguous index range
            const int a = cell_start[ix * s.rows + nMinCellY], b = cell_start[ix * s.rows + nMaxCellY + 1];
            for (int base = a; base < b; base += 32) {
                const int j = base + lane;
                bool hit = false;
                int id = 0;
                if (j < b) {
                    id = indices[j];
                    const pslot_keypoint kp = kps[id];
                    hit = !(check_levels && (kp.octave < Q.min_level || (Q.max_level >= 0 && kp.octave > Q.max_level)));
                    hit = hit && fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r;
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                const int p = n + __popc(m & ((1u << lane) - 1u));
                if (hit && p < cap) o[p] = id;
                n += __popc(m);
            }
        }
    }
    if (lane == 0) out_count[q] = n;
}

// ORBmatcher::SearchByProjection, one warp per projected map point; two variants share the window walk:
//   mode 0  (F, vpMapPoints, th)                 local map  (reference src/ORBmatcher.cc:68-155): best + second best,
//                                                ratio test when both lie in the same scale level;
//   mode 1  (CurrentFrame, LastFrame, th, bMono) last frame (:1613-1756): best only.
// The per-point window (centre, radius, level range, right-image coordinate and its tolerance) is computed by the
// caller exactly as the reference computes it; the kernel starts at GetFeaturesInArea.
// Both reference loops are order dependent: once a map point with observations has been written into
// F.mvpMapPoints[bestIdx] (:148, :1715), later map points skip that feature (:111-113, :1684-1686).  Restated as a
// fixed point: a sweep computes every point's match with "feature idx is taken for point i" <=> occupied[idx] or
// some point j < i with observations matched idx in the PREVIOUS sweep (claim_prev[idx] = the smallest such j).
// Point 0 is right after one sweep, point k after k + 1 at the latest, and a sweep that changes no match is the
// sequential result; in practice the dependency chains are a few links long.  Inside a point the candidates are
// visited in the reference's order (GetFeaturesInArea order), 32 at a time: distances in parallel, then the
// bookkeeping replayed in order, because bestLevel2 depends on the visiting order among equal distances.
__global__ void __launch_bounds__(256)
k_search_by_projection(const __grid_constant__ ProjArgs a)
{
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= a.n_pts) return;
    const ProjPoint P = a.pts[i];
    int match = -1;
    if (P.flags & 1) { // the reference reaches the search for this point
        const float rad = P.radius;
        const uint4 *md = reinterpret_cast<const uint4 *>(a.pt_desc + (size_t)i * 32);
        const uint4 m0 = __ldg(md), m1 = __ldg(md + 1);
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        const AreaCells A = area_cells(a.s, P.x, P.y, rad);
        if (A.ok) {
            const bool check_levels = P.min_level > 0 || P.max_level >= 0; // GetFeaturesInArea (:1829)
            for (int ix = A.x0; ix <= A.x1; ++ix) {
                const int c0 = a.cell_start[ix * a.s.rows + A.y0], c1 = a.cell_start[ix * a.s.rows + A.y1 + 1];
                for (int base = c0; base < c1; base += 32) {
                    const int j = base + lane;
                    bool use = false;
                    int idx = 0, dist = 0, oct = 0;
                    if (j < c1) {
                        idx = a.indices[j];
                        const pslot_keypoint kp = a.kps[idx];
                        oct = kp.octave;
                        // GetFeaturesInArea (:1837-1852)
                        use = !(check_levels && (oct < P.min_level || (P.max_level >= 0 && oct > P.max_level))) &&
                              fabsf(__fsub_rn(kp.x, P.x)) < rad && fabsf(__fsub_rn(kp.y, P.y)) < rad;
                        // F.mvpMapPoints[idx] with observations, as of the previous sweep
                        use = use && !(a.occupied[idx] || a.claim_prev[idx] < i);
                        if (use) {
                            const float ur = a.u_right[idx];
                            if (ur > 0) use = !(fabsf(__fsub_rn(P.xr, ur)) > P.er_limit); // (:115-120, :1688-1694)
                        }
                        if (use) {
                            const uint4 *fd = reinterpret_cast<const uint4 *>(a.frame_desc + (size_t)idx * 32);
                            const uint4 f0 = __ldg(fd), f1 = __ldg(fd + 1);
                            dist = __popc(m0.x ^ f0.x) + __popc(m0.y ^ f0.y) + __popc(m0.z ^ f0.z) + __popc(m0.w ^ f0.w) +
                                   __popc(m1.x ^ f1.x) + __popc(m1.y ^ f1.y) + __popc(m1.z ^ f1.z) + __popc(m1.w ^ f1.w);
                        }
                    }
                    // replay (:126-140, :1700-1704) over the usable candidates in visiting order
                    for (unsigned m = __ballot_sync(0xffffffffu, use); m; m &= m - 1) {
                        const int src = __ffs((int)m) - 1;
                        const int d = __shfl_sync(0xffffffffu, dist, src), lv = __shfl_sync(0xffffffffu, oct, src);
                        const int id = __shfl_sync(0xffffffffu, idx, src);
                        if (d < bestDist) {
                            bestDist2 = bestDist; bestDist = d;
                            bestLevel2 = bestLevel; bestLevel = lv;
                            bestIdx = id;
                        } else if (d < bestDist2) {
                            bestLevel2 = lv; bestDist2 = d;
                        }
                    }
                }
            }
        }
        // bestDist <= TH_HIGH (or ORBdist); local map only: ratio test against the second best if both are in the same level (:144-147)
        if (bestDist <= a.th_dist &&
            !(a.mode == 0 && bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(a.nn_ratio, (float)bestDist2)))
            match = bestIdx;
    }
    if (lane == 0) {
        a.match_next[i] = match;
        if (match != a.match_prev[i]) *a.changed = 1;
        if (match >= 0 && (P.flags & 2)) atomicMin(&a.claim_next[match], i); // Observations() > 0: blocks later points
    }
}

} // namespace

int launch_search_by_projection(const ProjArgs &a, cudaStream_t st)
{
    k_search_by_projection<<<(a.n_pts + 7) / 8, 256, 0, st>>>(a);
    return 1;
}

size_t grid_smem_bytes(const GridSpec &s) { return sizeof(int) * (2 * (size_t)s.cols * s.rows + 1); }

int launch_assign_grid(const pslot_keypoint *kps, const int *counts, int fixed_n, size_t stride, const GridSpec &s,
                       int *cell_start, int *indices, int nimages, cudaStream_t st)
{
    const size_t tables = grid_smem_bytes(s);
    // keypoint -> cell cache: room for the largest image of the launch if shared memory allows (else recomputed)
    const size_t want = counts ? stride : (size_t)fixed_n;
    const size_t room = (200 * 1024 - tables) / 2;
    const int cache_cap = (int)(want <= room ? want : 0);
    const size_t smem = tables + 2 * (size_t)cache_cap + 16;
    ensure_dyn_smem(k_assign_grid, smem);
    k_assign_grid<<<nimages, GT, smem, st>>>(kps, counts, fixed_n, stride, s, cell_start, indices, cache_cap);
    return 1;
}

int launch_features_in_area(const pslot_keypoint *kps, const int *cell_start, const int *indices, const GridSpec &s,
                            const pslot_area_query *queries, int nq, int cap, int *out_count, int *out, cudaStream_t st)
{
    k_features_in_area<<<(nq + 7) / 8, 256, 0, st>>>(kps, cell_start, indices, s, queries, nq, cap, out_count, out);
    return 1;
}
