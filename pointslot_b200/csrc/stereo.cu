// stereo.cu — K7+K8 fused and K9: Frame::ComputeStereoMatches (reference src/Frame.cc:2142-2316).
//
// k_stereo_match, one warp per LEFT keypoint, CTAs tiled over (chunk of left keypoints, pair):
//   * the reference's row table vRowIndices (:2152-2169) is replaced by its defining predicate: right
//     keypoint iR is a candidate of left row r  <=>  floor(yR - 2*s[oct]) <= r <= ceil(yR + 2*s[oct]),
//     the same set.  The band limits, octave and x of all right keypoints of the pair are staged in shared
//     memory once per CTA and bucketed by 16-row bins of their band (count, scan, scatter in shared memory), so
//     a left keypoint at row r scans only bin r >> 4 and applies the exact predicate; if the lists do not fit
//     (very many right keys, or bands spanning many bins) it scans every right key instead;
//   * lanes stride over the candidates, gate on band / octave+-1 / uR in [uL-maxD, uL] (:2200-2209)
//     and evaluate the 256-bit Hamming distance (ORBmatcher.cc:2704-2720) with __popc over 8 words;
//     the first strict minimum below TH_HIGH=100 (:2198,2219) in ascending iR is a warp argmin over
//     (dist << 16 | iR), whatever order the candidates were visited in;
//   * if best < thOrbDist=75 (:2228): 11x11 SAD at 11 offsets in the un-blurred pyramid level of the
//     left keypoint (:2231-2268), first strict minimum, parabola (:2273-2281), disparity/depth
//     (:2284-2298).  Float steps are single-rounded __f*_rn operations in the reference's order.
// k_stereo_reject, one CTA per pair: median of the accepted SADs by two-level radix select on the
// 16-bit SAD value, then reject SAD >= 1.5f*1.4f*median (:2302-2315).
//
// Integer parts are exact; the float tail is a handful of IEEE single operations -> bit-equal to the
// oracle in practice, tolerance 1e-3 px per north_star.
#include <math.h>

#include "pslot_internal.h"

namespace {

constexpr int SW = 8; // warps per CTA
constexpr int TH_HIGH = 100, TH_LOW = 50;

struct RightKey {
    short minr, maxr;
    int octave;
    float x;
};

struct SadSmem {
    uint8_t left[11 * 12];
    uint8_t right[11 * 24];
    int sad[12];
};

constexpr int BIN_SHIFT = 4;  // right keypoints are bucketed by 16-row bins of their band
constexpr int LEFT_PER_CTA = SW * 8;

__global__ void __launch_bounds__(SW * 32)
k_stereo_match(const __grid_constant__ Geom g, const __grid_constant__ StereoArgs a, int right_cap, int list_cap, int nbins)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RightKey *rk = reinterpret_cast<RightKey *>(smem_raw);                                   // [right_cap]
    uint16_t *blist = reinterpret_cast<uint16_t *>(smem_raw + sizeof(RightKey) * (size_t)right_cap); // [list_cap]
    int *bin_start = reinterpret_cast<int *>(blist + list_cap); // [nbins + 1]; list_cap is even
    int *bin_fill = bin_start + nbins + 1;                                                   // [nbins]
    __shared__ SadSmem sads[SW];
    __shared__ int s_use_bins;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pair = blockIdx.y;
    const int nL = a.n_l[pair], nR = min(a.n_r[pair], right_cap);
    // more right keypoints than were staged, or than the 16-bit iR of the packed argmin can name: an error the host
    // reports (PSLOT_ERR_CAPACITY), never a silent truncation
    if ((a.n_r[pair] > right_cap || nR > 0xFFFF) && threadIdx.x == 0 && a.err_flag) atomicOr(a.err_flag, PSLOT_FLAG_STEREO_NR);
    const size_t row0 = (size_t)pair * a.K;
    const pslot_keypoint *kR = a.kps_r + row0;
    const pslot_keypoint *kL = a.kps_l + row0;

    const int iL0 = blockIdx.x * LEFT_PER_CTA;
    if (iL0 >= nL) return;

    // ---- stage the right keys and bucket them by row bin -------------------------------------------------
    // The reference's row table vRowIndices[r] lists the right keypoints whose band contains row r (:2152-2169); the
    // candidates of a left keypoint at row r are a subset of the bin r >> 4, and the band test below restores the
    // exact set.  The visiting order inside a bin is irrelevant: the argmin is taken over (distance, iR) pairs.
    for (int i = tid; i <= nbins; i += SW * 32) bin_start[i] = 0;
    __syncthreads();
    for (int i = tid; i < nR; i += SW * 32) {
        const pslot_keypoint k = kR[i];
        const int oct = min(max(k.octave, 0), g.nlevels - 1);
        if (oct != k.octave && a.err_flag) atomicOr(a.err_flag, PSLOT_FLAG_STEREO_OCTAVE);
        const float r = __fmul_rn(2.0f, g.lv[oct].scale); // (:2162)
        RightKey e;
        e.maxr = (short)(int)ceilf(__fadd_rn(k.y, r));
        e.minr = (short)(int)floorf(__fsub_rn(k.y, r));
        e.octave = oct;
        e.x = k.x;
        rk[i] = e;
        const int b0 = min(max((int)e.minr >> BIN_SHIFT, 0), nbins - 1), b1 = min(max((int)e.maxr >> BIN_SHIFT, 0), nbins - 1);
        for (int b = b0; b <= b1; ++b) atomicAdd(&bin_start[b + 1], 1);
    }
    __syncthreads();
    if (warp == 0) { // inclusive scan of the bin counts (nbins <= a few hundred)
        int carry = 0;
        for (int base = 0; base <= nbins; base += 32) {
            const int i = base + lane;
            int v = i <= nbins ? bin_start[i] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            if (i <= nbins) bin_start[i] = v + carry;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) s_use_bins = carry <= list_cap; // very tall bands or no room for the lists: scan every right key
    }
    __syncthreads();
    const bool use_bins = s_use_bins != 0;
    if (use_bins) {
        for (int i = tid; i < nbins; i += SW * 32) bin_fill[i] = 0;
        __syncthreads();
        for (int i = tid; i < nR; i += SW * 32) {
            const RightKey e = rk[i];
            const int b0 = min(max((int)e.minr >> BIN_SHIFT, 0), nbins - 1), b1 = min(max((int)e.maxr >> BIN_SHIFT, 0), nbins - 1);
            for (int b = b0; b <= b1; ++b) blist[bin_start[b] + atomicAdd(&bin_fill[b], 1)] = (uint16_t)i;
        }
        __syncthreads();
    }

    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    unsigned n_row = 0, n_cmp = 0; // work counters of this lane (only reported when a.counters is set)
    for (int q = 0; q < LEFT_PER_CTA / SW; ++q) {
        const int iL = iL0 + q * SW + warp;
        if (iL >= nL) continue; // warp-uniform
        const pslot_keypoint kp = kL[iL];
        const int levelL = kp.octave;
        const float vL = kp.y, uL = kp.x;
        const int row = (int)vL;
        const float minU = __fsub_rn(uL, a.max_d), maxU = uL; // minD = 0 (:2173,2192-2193)

        // ---- Hamming search -----------------------------------------------------------------------
        const uint4 *dl = reinterpret_cast<const uint4 *>(a.desc_l + (row0 + iL) * 32);
        const uint4 l0 = __ldg(dl), l1 = __ldg(dl + 1);
        unsigned best = ((unsigned)TH_HIGH << 16) | 0xFFFFu;
        if (maxU >= 0) {
            int jb = 0, je = nR;
            if (use_bins) {
                const int b = min(max(row >> BIN_SHIFT, 0), nbins - 1);
                jb = bin_start[b]; je = bin_start[b + 1];
            }
            for (int j = jb + lane; j < je; j += 32) {
                const int iR = use_bins ? (int)blist[j] : j;
                const RightKey e = rk[iR];
                if (row < e.minr || row > e.maxr) continue;
                ++n_row;
                if (e.octave < levelL - 1 || e.octave > levelL + 1) continue;
                if (!(e.x >= minU && e.x <= maxU)) continue;
                const uint4 *dr = reinterpret_cast<const uint4 *>(a.desc_r + (row0 + iR) * 32);
                const uint4 r0 = __ldg(dr), r1 = __ldg(dr + 1);
                ++n_cmp;
                const int dist = __popc(l0.x ^ r0.x) + __popc(l0.y ^ r0.y) + __popc(l0.z ^ r0.z) + __popc(l0.w ^ r0.w) +
                                 __popc(l1.x ^ r1.x) + __popc(l1.y ^ r1.y) + __popc(l1.z ^ r1.z) + __popc(l1.w ^ r1.w);
                // first strict minimum in ascending iR (:2219) == minimum of (dist, iR), whatever the visiting order
                best = min(best, ((unsigned)dist << 16) | (unsigned)iR);
            }
        }
        best = __reduce_min_sync(0xffffffffu, best);
        const int bestDist = (int)(best >> 16);
        const int bestIdxR = bestDist < TH_HIGH ? (int)(best & 0xFFFFu) : -1;

        float out_u = -1.0f, out_d = -1.0f;
        int out_sad = -1;
        if (bestDist < thOrbDist) {
            // ---- 11x11 SAD sweep at the keypoint's pyramid level (:2231-2268) -----------------------
            const LevelGeom &L = g.lv[levelL];
            const float uR0 = rk[bestIdxR].x;
            const float sf = L.inv_scale;
            const int su = (int)roundf(__fmul_rn(kp.x, sf));
            const int sv = (int)roundf(__fmul_rn(kp.y, sf));
            const int sr = (int)roundf(__fmul_rn(uR0, sf));
            const int w = 5, Ls = 5;
            // iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1 (:2249-2252)
            const bool ok = !(sr + Ls - w < 0 || sr + Ls + w + 1 >= L.w) &&
                            // in-bounds guard for the loads (always true for extractor keypoints)
                            su - w >= 0 && su + w < L.w && sv - w >= 0 && sv + w < L.h && sr - Ls - w >= 0;
            if (ok) {
                SadSmem &S = sads[warp];
                const uint8_t *pl = a.pyr_l.lv[levelL] + (size_t)pair * a.pyr_l.slot_stride[levelL];
                const uint8_t *pr = a.pyr_r.lv[levelL] + (size_t)pair * a.pyr_r.slot_stride[levelL];
                for (int e = lane; e < 121; e += 32) {
                    const int yy = e / 11, xx = e - yy * 11;
                    S.left[yy * 12 + xx] = __ldg(pl + (size_t)(sv - w + yy) * L.pitch + (su - w + xx));
                }
                for (int e = lane; e < 231; e += 32) {
                    const int yy = e / 21, xx = e - yy * 21;
                    S.right[yy * 24 + xx] = __ldg(pr + (size_t)(sv - w + yy) * L.pitch + (sr - Ls - w + xx));
                }
                if (lane < 11) S.sad[lane] = 0;
                __syncwarp();
                const int cl = S.left[5 * 12 + 5];
                for (int t = lane; t < 121; t += 32) {
                    const int inc = t / 11, yy = t - inc * 11; // inc = incR + 5
                    const int cr = S.right[5 * 24 + inc + 5];
                    int acc = 0;
#pragma unroll
                    for (int xx = 0; xx < 11; ++xx)
                        acc += abs(((int)S.left[yy * 12 + xx] - cl) - ((int)S.right[yy * 24 + inc + xx] - cr));
                    atomicAdd(&S.sad[inc], acc);
                }
                __syncwarp();
                int bestSAD = 0x7fffffff, bestinc = 0;
#pragma unroll
                for (int inc = 0; inc < 11; ++inc) {
                    const int d = S.sad[inc];
                    if (d < bestSAD) { bestSAD = d; bestinc = inc - Ls; }
                }
                if (!(bestinc == -Ls || bestinc == Ls)) {
                    const float d1 = (float)S.sad[Ls + bestinc - 1], d2 = (float)S.sad[Ls + bestinc], d3 = (float)S.sad[Ls + bestinc + 1];
                    // deltaR = (d1-d3)/(2*(d1+d3-2*d2)) (:2278)
                    const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2)));
                    const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), den);
                    if (!(deltaR < -1 || deltaR > 1)) {
                        float bestuR = __fmul_rn(L.scale, __fadd_rn(__fadd_rn((float)sr, (float)bestinc), deltaR));
                        float disparity = __fsub_rn(uL, bestuR);
                        if (disparity >= 0 && disparity < a.max_d) {
                            if (disparity <= 0) {
                                disparity = 0.01f;
                                bestuR = (float)((double)uL - 0.01);
                            }
                            out_d = __fdiv_rn(a.bf, disparity);
                            out_u = bestuR;
                            out_sad = bestSAD;
                        }
                    }
                }
                __syncwarp();
            }
        }
        if (lane == 0) {
            a.u_right[row0 + iL] = out_u;
            a.depth[row0 + iL] = out_d;
            a.best_dist[row0 + iL] = bestDist;
            a.best_idx[row0 + iL] = bestIdxR;
            a.sad[row0 + iL] = out_sad;
        }
    }
    if (a.counters) {
        n_row = __reduce_add_sync(0xffffffffu, n_row);
        n_cmp = __reduce_add_sync(0xffffffffu, n_cmp);
        if (lane == 0) { atomicAdd(a.counters, (unsigned long long)n_row); atomicAdd(a.counters + 1, (unsigned long long)n_cmp); }
    }
}

// median of accepted SADs + rejection (:2302-2315)
__global__ void __launch_bounds__(256)
k_stereo_reject(const __grid_constant__ StereoArgs a)
{
    __shared__ int hist[256];
    __shared__ int s_bin, s_rank, s_count, s_sum;
    const int tid = threadIdx.x;
    const int pair = blockIdx.x;
    const int nL = a.n_l[pair];
    const size_t row0 = (size_t)pair * a.K;
    const int *sad = a.sad + row0;

    hist[tid] = 0;
    if (tid == 0) s_sum = 0;
    __syncthreads();
    int cnt = 0;
    for (int i = tid; i < nL; i += 256) {
        const int s = sad[i];
        if (s >= 0) { atomicAdd(&hist[(s >> 8) & 0xFF], 1); ++cnt; }
    }
    if (cnt) atomicAdd(&s_sum, cnt);
    __syncthreads();
    const int M = s_sum;
    if (M == 0) return; // empty vDistIdx: no-op (DESIGN.md parity rule 4)
    const int k = M / 2; // vDistIdx[size/2] of the ascending sort
    if (tid == 0) {
        int acc = 0, b = 0;
        for (; b < 256; ++b) {
            if (acc + hist[b] > k) break;
            acc += hist[b];
        }
        s_bin = b; s_rank = k - acc;
    }
    __syncthreads();
    const int bin = s_bin, rank = s_rank;
    __syncthreads();
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < nL; i += 256) {
        const int s = sad[i];
        if (s >= 0 && ((s >> 8) & 0xFF) == bin) atomicAdd(&hist[s & 0xFF], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0, b = 0;
        for (; b < 256; ++b) {
            if (acc + hist[b] > rank) break;
            acc += hist[b];
        }
        s_count = (bin << 8) | b;
    }
    __syncthreads();
    const float median = (float)s_count;
    const float thDist = __fmul_rn(1.5f * 1.4f, median);
    for (int i = tid; i < nL; i += 256) {
        const int s = sad[i];
        if (s >= 0 && !((float)s < thDist)) {
            a.u_right[row0 + i] = -1.0f;
            a.depth[row0 + i] = -1.0f;
        }
    }
}

} // namespace

int launch_stereo(const Geom &g, const StereoArgs &a, int npairs, int max_n_r, cudaStream_t st, cudaEvent_t between)
{
    const int nbins = ((g.lv[0].h + (1 << BIN_SHIFT) - 1) >> BIN_SHIFT) + 1;
    const size_t fixed = sizeof(RightKey) * (size_t)max_n_r + sizeof(int) * (2 * (size_t)nbins + 2);
    // a band of 2 * 2 * scale + 2 rows touches at most 3 bins up to scale 7.5; lists are dropped when they do not fit
    int list_cap = (3 * max_n_r + 1) & ~1;
    if (fixed + 2 * (size_t)list_cap > 220 * 1024) list_cap = 0;
    const size_t smem = fixed + 2 * (size_t)list_cap;
    ensure_dyn_smem(k_stereo_match, smem);
    dim3 grid((a.K + LEFT_PER_CTA - 1) / LEFT_PER_CTA, npairs);
    k_stereo_match<<<grid, SW * 32, smem, st>>>(g, a, max_n_r, list_cap, nbins);
    if (between) cudaEventRecord(between, st);
    k_stereo_reject<<<npairs, 256, 0, st>>>(a);
    return 2;
}
