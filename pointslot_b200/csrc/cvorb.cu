// cvorb.cu — OpenCV's own ORB, `cv::ORB::create(nfeatures, 1.2f, nlevels, edgeThreshold)->detectAndCompute(image, mask)`,
// the second extractor PointSLOT runs per frame for its dense object features (reference src/Frame.cc:2623-2627, called
// from Frame::ExtractObjORB :2628-2665).  The algorithm is OpenCV's (modules/features2d/src/orb.cpp:
// ORB_Impl::detectAndCompute, computeKeyPoints, HarrisResponses, ICAngles, computeOrbDescriptors; 4.13 semantics), with
// the arithmetic that the oracle established against cv2.ORB_create (tests/test_oracle_cvorb.py):
//   * pyramid: every level is resize(previous level, INTER_LINEAR_EXACT): 8-bit coefficients, (V + 2^15) >> 16; the mask
//     pyramid is resized the same way and thresholded (THRESH_TOZERO, 254) at every level;
//   * FAST 9-16 with non-maximum suppression over the WHOLE level (not per cell), then the mask and the edgeThreshold
//     border filters, retainBest(2 * quota) on the FAST score, the Harris response of the survivors (7x7 block, k = 0.04,
//     float expression in OpenCV's order) and retainBest(quota) on it - ties are kept by both, as KeyPointsFilter does;
//   * orientation = intensity centroid on the unblurred level (same as ORB-SLAM's IC_Angle);
//   * descriptor = steered BRIEF on GaussianBlur(level, 7x7, sigma 2) evaluated on a sub-matrix, i.e. OpenCV's FLOAT
//     separable filter: row taps left to right, column taps from the centre outwards over the symmetric pairs, every
//     multiply-add fused (__fmaf_rn), lrintf, saturate; steering by glibc cosf / sinf of the float angle.
// Output order: (octave, y, x).  cv::ORB's own order is whatever std::nth_element leaves; the SET and every per-keypoint
// value are the contract.
//
// Kernels (a call carries S images of one size - "slots", e.g. the left and right view of Frame::ExtractObjORB; every kernel
// covers all slots and, except the resize chain, all pyramid levels):
//   k_cv_resize_exact   level l from level l-1 (image and mask)                      HBM stream, tiny
//   k_cv_fast_score     raw FAST score map of every level (0 where not a corner): byte-SWAR rejection test per word,
//                       exact scores of the few survivors on dense warps
//   k_cv_nms_collect    strict 3x3 maxima + mask + border -> candidate list (one atomic per CTA) + score histogram per level
//   k_cv_harris         retainBest(2q) threshold from the histogram; Harris response of the survivors
//   k_cv_select         per level: radix select of the quota-th largest Harris response, ties kept, bitonic sort by (y, x)
//   k_cv_describe       one warp per keypoint: orientation, float blur of the 43x43 patch, 256 tests
#include "pslot_internal.h"
#include "orb_common.cuh"

namespace {

// ---- pyramid --------------------------------------------------------------------------------------------------------------
// thread = 4 adjacent pixels of one output row (one word store; rows are 128-byte aligned and padded, so the word never
// leaves the row); blockIdx.z = 2 * slot + (0: the image, 1: the object mask (same taps; thresholded)); without masks
// blockIdx.z = slot
__global__ void __launch_bounds__(256)
k_cv_resize_exact(const uint8_t *__restrict__ src_img, const uint8_t *__restrict__ src_mask, int sw, int sh, int spitch,
                  uint8_t *__restrict__ dst_img, uint8_t *__restrict__ dst_mask, int dw, int dh, int dpitch,
                  const int2 *__restrict__ xt, const int2 *__restrict__ yt, size_t slot_bytes, int planes)
{
    const int x4 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4, y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x4 >= dw || y >= dh) return;
    const int slot = blockIdx.z / planes;
    const bool is_mask = blockIdx.z - slot * planes != 0;
    const uint8_t *src = (is_mask ? src_mask : src_img) + slot * slot_bytes;
    const int2 ty = yt[y];
    const int ya = ty.x, yb = min(ya + 1, sh - 1), ay = ty.y;
    const uint8_t *r0 = src + (size_t)ya * spitch, *r1 = src + (size_t)yb * spitch;
    uint32_t out = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int2 tx = xt[min(x4 + b, dw - 1)];
        const int xa = tx.x, xb = min(xa + 1, sw - 1), ax = tx.y;
        int p00 = r0[xa], p01 = r0[xb], p10 = r1[xa], p11 = r1[xb];
        if (is_mask) { // cv::ORB looks at a mask as zero / non-zero; levels >= 1 hold 0 / 255 already (the threshold below)
            p00 = p00 ? 255 : 0; p01 = p01 ? 255 : 0; p10 = p10 ? 255 : 0; p11 = p11 ? 255 : 0;
        }
        const int H0 = p00 * (256 - ax) + p01 * ax;
        const int H1 = p10 * (256 - ax) + p11 * ax;
        int v = (H0 * (256 - ay) + H1 * ay + (1 << 15)) >> 16;
        if (is_mask) v = v > 254 ? v : 0; // cv::threshold(mask, mask, 254, 0, THRESH_TOZERO)
        if (x4 + b < dw) out |= (uint32_t)v << (8 * b);
    }
    *reinterpret_cast<uint32_t *>((is_mask ? dst_mask : dst_img) + slot * slot_bytes + (size_t)y * dpitch + x4) = out;
}

// ---- FAST score map ---------------------------------------------------------------------------------------------------------
// R = max(max_k min_{j=k..k+8} (v - p_j), max_k min_{j=k..k+8} (p_j - v)) - 1; corner at threshold t <=> R >= t
// (cv::FAST 9-16, cornerScore<16>).  cv::FAST tests every pixel >= 3 px inside the level, but the
// border filter keeps only [edge, dim - edge) and the maxima test looks one pixel further, so only the band
// [edge - 1, dim - edge] of rows and columns is scored; everything else keeps the 0 the buffer was cleared to.
//
// Tile = 128 px x 8 rows, one thread per aligned 4-pixel word.  Pass 1: byte-SWAR rejection test (a 9-arc holds one pixel
// of every opposite pair, so both compass pairs need a member farther than t from the centre) on the word and its four
// compass words; the word's scores are cleared and the few pixels that pass go to a shared list.  Pass 2: the listed
// pixels, one per thread on dense warps, get the exact score.  blockIdx.x numbers the tiles of all levels, blockIdx.y = slot.
__device__ __forceinline__ int tile_level(const CvGeom &g, int tile)
{
    int l = 0;
    while (l + 1 < g.nlevels && tile >= g.lv[l + 1].tile_base) ++l;
    return l;
}

__device__ __forceinline__ int fast_score_exact(const uint8_t *p, int P)
{
    const int v = p[0];
    int d[25];
    d[0] = v - p[3 * P];      d[1] = v - p[3 * P + 1];  d[2] = v - p[2 * P + 2];   d[3] = v - p[P + 3];
    d[4] = v - p[3];          d[5] = v - p[-P + 3];     d[6] = v - p[-2 * P + 2];  d[7] = v - p[-3 * P + 1];
    d[8] = v - p[-3 * P];     d[9] = v - p[-3 * P - 1]; d[10] = v - p[-2 * P - 2]; d[11] = v - p[-P - 3];
    d[12] = v - p[-3];        d[13] = v - p[P - 3];     d[14] = v - p[2 * P - 2];  d[15] = v - p[3 * P - 1];
#pragma unroll
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    // best = max_k max(min over arc k of (v - p), min over arc k of (p - v)); the second term is -(min_k max over arc k of
    // (v - p)), kept as a min-of-max until the end.  (Written as max(best, max(mn, -mx)) inside the loop, ptxas 12.9 fuses
    // the negated operand into VIMNMX3 and loses the sign: the kernel then scored max |v - p| - found by the parity test.)
    int hi = -256, lo = 256;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
#pragma unroll
        for (int j = 1; j < 9; ++j) { mn = min(mn, d[k + j]); mx = max(mx, d[k + j]); }
        hi = max(hi, mn);
        lo = min(lo, mx);
    }
    return max(hi, -lo) - 1;
}

constexpr int CT_ROWS = 32; // tile = 128 px x 32 rows: a thread owns one word column and the rows warp, warp + 8, ...

__global__ void __launch_bounds__(256)
k_cv_fast_score(const __grid_constant__ CvGeom g, const uint8_t *__restrict__ pyr, uint8_t *__restrict__ score)
{
    __shared__ uint32_t s_list[128 * CT_ROWS];
    __shared__ int s_n;
    const int l = tile_level(g, blockIdx.x);
    const CvLevel &L = g.lv[l];
    const int tile = blockIdx.x - L.tile_base, ty = tile / L.tiles_x, tx = tile - ty * L.tiles_x;
    const int gx = tx * 32 + (threadIdx.x & 31), y0 = ty * CT_ROWS + (threadIdx.x >> 5) + g.edge - 1;
    const int P = L.pitch, PW = P >> 2, t = g.fast_th;
    const int x_lo = g.edge - 1, x_hi = L.w - g.edge, y_hi = L.h - g.edge; // scored columns / last row, inclusive
    pyr += blockIdx.y * g.slot_bytes + L.off;
    score += blockIdx.y * g.slot_bytes + L.off;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    if (4 * gx <= x_hi && 4 * gx + 3 >= x_lo) {
        // all 20 words of the thread's four rows are requested before the first one is used
        uint32_t c[4], lft[4], rgt[4], top[4], bot[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int y = min(y0 + 8 * i, y_hi); // clamped rows repeat the last one; their result is dropped below
            const uint32_t *row = reinterpret_cast<const uint32_t *>(pyr + (size_t)y * P) + gx;
            c[i] = row[0]; lft[i] = gx > 0 ? row[-1] : 0u; rgt[i] = gx + 1 < PW ? row[1] : 0u;
            top[i] = row[-3 * PW]; bot[i] = row[3 * PW];
        }
        // columns of this word outside the scored band
        uint32_t colmask = 0x80808080u;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (4 * gx + b < x_lo || 4 * gx + b > x_hi) colmask &= ~(0x80u << (8 * b));
        // bit 7 of ((d + 127 - t) | d) is set where d > t; a carry from the byte below only adds a false positive
        const uint32_t bias = t <= 127 ? (uint32_t)(127 - t) * 0x01010101u : 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int y = y0 + 8 * i;
            if (y > y_hi) break;
            const uint32_t east = __funnelshift_r(c[i], rgt[i], 24), west = __funnelshift_r(lft[i], c[i], 8); // pixels x + 3 / x - 3
            const uint32_t d0 = __vabsdiffu4(bot[i], c[i]), d8 = __vabsdiffu4(top[i], c[i]);
            const uint32_t d4 = __vabsdiffu4(east, c[i]), d12 = __vabsdiffu4(west, c[i]);
            const uint32_t e0 = (d0 + bias) | d0, e8 = (d8 + bias) | d8, e4 = (d4 + bias) | d4, e12 = (d12 + bias) | d12;
            uint32_t pass = (e0 | e8) & (e4 | e12) & colmask;
            reinterpret_cast<uint32_t *>(score + (size_t)y * P)[gx] = 0u;
            while (pass) {
                const int b = (__ffs((int)pass) - 1) >> 3;
                s_list[atomicAdd(&s_n, 1)] = ((uint32_t)y << 16) | (uint32_t)(4 * gx + b);
                pass &= pass - 1;
            }
        }
    }
    __syncthreads(); // also orders the cleared words before the score bytes written below
    const int n = s_n;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int x = (int)(s_list[i] & 0xFFFFu), yy = (int)(s_list[i] >> 16);
        const int R = fast_score_exact(pyr + (size_t)yy * P + x, P);
        if (R >= t) score[(size_t)yy * P + x] = (uint8_t)R;
    }
}

// strict 3x3 maxima over the whole level, runByPixelsMask, runByImageBorder; survivors go to the level's candidate list
// (unordered: everything downstream is a set operation until the final sort).  Same tiles: a CTA gathers its survivors in
// shared memory and reserves its piece of the level's list with ONE atomic (a global counter per candidate serialised
// 20 000 atomics per image on eight addresses: 407 us per 64 images, 117 us like this).
__global__ void __launch_bounds__(256)
k_cv_nms_collect(const __grid_constant__ CvGeom g, const uint8_t *__restrict__ score, const uint8_t *__restrict__ mask,
                 uint32_t *__restrict__ cand, int *__restrict__ counters, int *__restrict__ hist, int *__restrict__ err)
{
    __shared__ uint32_t s_c[32 * CT_ROWS]; // strict 3x3 maxima: at most one per 2x2 block of the tile
    __shared__ int s_n, s_base;
    const int l = tile_level(g, blockIdx.x), slot = blockIdx.y;
    const CvLevel &L = g.lv[l];
    const int tile = blockIdx.x - L.tile_base, ty = tile / L.tiles_x, tx = tile - ty * L.tiles_x;
    const int gx = tx * 32 + (threadIdx.x & 31);
    const int P = L.pitch;
    score += slot * g.slot_bytes + L.off;
    if (mask) mask += slot * g.slot_bytes + L.off;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    if (4 * gx + 3 >= g.edge && 4 * gx < L.w - g.edge) {
        const int y0 = ty * CT_ROWS + (threadIdx.x >> 5) + g.edge; // candidates live in [edge, h - edge)
        uint32_t wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int y = y0 + 8 * i;
            wv[i] = y < L.h - g.edge ? *reinterpret_cast<const uint32_t *>(score + (size_t)y * P + 4 * gx) : 0u;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int y = y0 + 8 * i;
            for (uint32_t w = wv[i]; w;) {
                const int b = (__ffs((int)w) - 1) >> 3;
                w &= ~(0xFFu << (8 * b));
                const int x = 4 * gx + b;
                if (x < g.edge || x >= L.w - g.edge) continue;
                const uint8_t *s = score + (size_t)y * P + x;
                const int c = s[0];
                if (!(c > s[-1] && c > s[1] && c > s[-P - 1] && c > s[-P] && c > s[-P + 1] && c > s[P - 1] && c > s[P] && c > s[P + 1])) continue;
                if (mask && mask[(size_t)y * P + x] == 0) continue;
                s_c[atomicAdd(&s_n, 1)] = pslot_pack(x, y, c);
            }
        }
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;
    if (threadIdx.x == 0) s_base = atomicAdd(&counters[slot * CV_CNT_TOTAL + CV_CNT_CAND + l], n);
    __syncthreads();
    const int base = s_base;
    for (int i = threadIdx.x; i < n; i += 256) {
        if (base + i >= L.cand_cap) { atomicOr(err, PSLOT_FLAG_CVORB_CAP); continue; }
        const uint32_t c = s_c[i];
        cand[(size_t)slot * g.slot_cand + L.cand_base + base + i] = c;
        atomicAdd(&hist[(slot * PSLOT_MAX_LEVELS + l) * 256 + pslot_pr(c)], 1);
    }
}

// ---- k-th largest over a 256-bin histogram, block-parallel ----------------------------------------------------------------------
// Threads 0..255 each hold one bin count (threads beyond hold 0; the first 8 warps must all call).  Returns in every thread
// the highest bin d whose suffix sum S(d) = sum_{b >= d} cnt[b] reaches k (0 when the whole histogram holds fewer than k),
// and via `above` the number of entries in bins > d.  One suffix scan instead of a serial walk by thread 0 (ncu: 52 % of
// k_cv_harris' stall samples sat on the barrier behind that walk).
__device__ __forceinline__ int kth_bin_from_top(int cnt, int k, int *s_part /* >= 10 ints */, int &above)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // inclusive suffix sum inside the warp (towards higher lanes = higher bins)
    int suf = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_down_sync(0xffffffffu, suf, o);
        if (lane + o < 32) suf += v;
    }
    if (tid < 256 && lane == 0) s_part[warp] = suf; // warp total
    if (tid == 0) { s_part[8] = 0; s_part[9] = 0; }
    __syncthreads();
    int higher = 0; // bins of the warps above this one
    if (tid < 256)
        for (int w = warp + 1; w < 8; ++w) higher += s_part[w];
    const int S = suf + higher;     // S(tid)
    const int S1 = S - cnt;         // S(tid + 1)
    if (tid < 256 && S >= k && S1 < k) { s_part[8] = tid; s_part[9] = S1; }
    __syncthreads();
    const int d = s_part[8];
    above = s_part[9];
    __syncthreads(); // s_part may be reused by the caller's next pass
    return d;
}

// ---- retainBest(2 * quota) on the FAST score, then HarrisResponses -----------------------------------------------------------
// HarrisResponses of one candidate (orb.cpp: 7x7 block of Sobel-3 gradients, k = 0.04): the 9x9 byte patch is fetched as
// three aligned words per row (27 loads instead of 392 byte loads) and the Sobel kernels are applied separably, row by row:
// d = p[x+1] - p[x-1], s = p[x-1] + 2 p[x] + p[x+1];  Ix = d[y-1] + 2 d[y] + d[y+1],  Iy = s[y+1] - s[y-1].  Integer sums,
// so the order of the additions is free.
__device__ __forceinline__ float harris_response(const uint8_t *__restrict__ img, int P, int x0, int y0, float scale_sq_sq)
{
    const int xs = x0 - 4, o8 = 8 * (xs & 3);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(img + (size_t)(y0 - 4) * P + (xs & ~3));
    const int PW = P >> 2;
    int d[3][7], sm[3][7];
    int a = 0, b = 0, cc = 0;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const uint32_t w0 = w[j * PW], w1 = w[j * PW + 1], w2 = w[j * PW + 2];
        const uint32_t q0 = __funnelshift_r(w0, w1, o8), q1 = __funnelshift_r(w1, w2, o8), q2 = w2 >> o8;
        int p[9];
#pragma unroll
        for (int k = 0; k < 4; ++k) { p[k] = (int)((q0 >> (8 * k)) & 0xFFu); p[4 + k] = (int)((q1 >> (8 * k)) & 0xFFu); }
        p[8] = (int)(q2 & 0xFFu);
#pragma unroll
        for (int x = 0; x < 7; ++x) {
            d[j % 3][x] = p[x + 2] - p[x];
            sm[j % 3][x] = p[x] + 2 * p[x + 1] + p[x + 2];
        }
        if (j >= 2) { // block row j - 1: rows j - 2, j - 1, j are in the ring
#pragma unroll
            for (int x = 0; x < 7; ++x) {
                const int Ix = d[(j - 2) % 3][x] + 2 * d[(j - 1) % 3][x] + d[j % 3][x];
                const int Iy = sm[j % 3][x] - sm[(j - 2) % 3][x];
                a += Ix * Ix; b += Iy * Iy; cc += Ix * Iy;
            }
        }
    }
    const float harris_k = 0.04f;
    const float fa = (float)a, fb = (float)b, fc = (float)cc;
    const float tr = __fadd_rn(fa, fb);
    return __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(harris_k, tr), tr)), scale_sq_sq);
}

// Per (level, slot): the retainBest(2 * quota) threshold from the level's score histogram, then rounds of 1024 candidates:
// the ones at or above the threshold (about one in seven) are compacted into a shared list so that the response runs on
// dense warps, and a round reserves its piece of the kept list with one atomic.
__global__ void __launch_bounds__(256)
k_cv_harris(const __grid_constant__ CvGeom g, const uint8_t *__restrict__ pyr, const uint32_t *__restrict__ cand,
            int *__restrict__ counters, const int *__restrict__ hist, uint32_t *__restrict__ kept, float *__restrict__ kept_resp)
{
    __shared__ int s_part[10];
    __shared__ uint32_t s_keep[1024];
    __shared__ int s_nk, s_out;
    const int l = blockIdx.y, slot = blockIdx.z;
    const CvLevel &L = g.lv[l];
    pyr += slot * g.slot_bytes + L.off;
    cand += (size_t)slot * g.slot_cand + L.cand_base;
    kept += (size_t)slot * g.slot_cand + L.cand_base; kept_resp += (size_t)slot * g.slot_cand + L.cand_base;
    counters += slot * CV_CNT_TOTAL;
    hist += slot * PSLOT_MAX_LEVELS * 256;
    const int n = min(counters[CV_CNT_CAND + l], L.cand_cap);
    // KeyPointsFilter::retainBest(keypoints, 2 * quota): keep everything >= the response of the (2 * quota)-th best
    const int keep = 2 * L.quota;
    int above;
    const int bin = kth_bin_from_top(hist[l * 256 + threadIdx.x], keep, s_part, above);
    const int thr = keep == 0 ? 256 : n > keep ? bin : 0;
    const float scale1 = 1.f / ((1 << 2) * 7 * 255.f);
    const float scale_sq_sq = __fmul_rn(__fmul_rn(__fmul_rn(scale1, scale1), scale1), scale1);
    for (int base = blockIdx.x * 1024; base < n; base += gridDim.x * 1024) { // block-uniform
        if (threadIdx.x == 0) s_nk = 0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = base + j * 256 + threadIdx.x;
            if (i < n) {
                const uint32_t c = cand[i];
                if (pslot_pr(c) >= thr) s_keep[atomicAdd(&s_nk, 1)] = c;
            }
        }
        __syncthreads();
        const int nk = s_nk;
        if (threadIdx.x == 0 && nk) s_out = atomicAdd(&counters[CV_CNT_KEPT + l], nk); // <= cand count: the kept arrays share the capacity
        __syncthreads();
        for (int q = threadIdx.x; q < nk; q += 256) {
            const uint32_t c = s_keep[q];
            kept[s_out + q] = c;
            kept_resp[s_out + q] = harris_response(pyr, L.pitch, pslot_px(c), pslot_py(c), scale_sq_sq);
        }
        __syncthreads(); // s_nk / s_out / s_keep are reused by the next round
    }
}

// ---- retainBest(quota) on the Harris response + (y, x) order -------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_key(float f) // monotone: larger float <=> larger key
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(1024)
k_cv_select(const __grid_constant__ CvGeom g, const uint32_t *__restrict__ kept, const float *__restrict__ kept_resp,
            int *__restrict__ counters, uint32_t *__restrict__ final_kp, float *__restrict__ final_resp, int *__restrict__ err)
{
    __shared__ unsigned long long s_e[CV_SORT_CAP];
    __shared__ int s_hist[256];
    __shared__ uint32_t s_prefix;
    __shared__ int s_k, s_n, s_part[10];
    const int l = blockIdx.x, slot = blockIdx.y, tid = threadIdx.x;
    const CvLevel &L = g.lv[l];
    counters += slot * CV_CNT_TOTAL;
    final_kp += (size_t)slot * g.slot_kp; final_resp += (size_t)slot * g.slot_kp;
    const int n1 = counters[CV_CNT_KEPT + l];
    const uint32_t *K = kept + (size_t)slot * g.slot_cand + L.cand_base;
    const float *R = kept_resp + (size_t)slot * g.slot_cand + L.cand_base;
    uint32_t thr_key = 0; // keep everything
    if (L.quota == 0) thr_key = 0xFFFFFFFFu;
    if (n1 > L.quota && L.quota > 0) {
        // the quota-th largest response by radix select over the monotone keys, one byte per pass
        if (tid == 0) { s_prefix = 0; s_k = L.quota; }
        uint32_t mask = 0;
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (tid < 256) s_hist[tid] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            for (int i = tid; i < n1; i += 1024) {
                const uint32_t key = float_key(R[i]);
                if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255], 1);
            }
            __syncthreads();
            int above;
            const int kk = s_k;
            const int d = kth_bin_from_top(tid < 256 ? s_hist[tid] : 0, kk, s_part, above); // kk <= entries matching the prefix
            if (tid == 0) {
                s_k = kk - above;
                s_prefix = prefix | ((uint32_t)d << shift);
            }
            mask |= 255u << shift;
            __syncthreads();
        }
        thr_key = s_prefix;
    }
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int i = tid; i < n1; i += 1024) {
        if (float_key(R[i]) >= thr_key && thr_key != 0xFFFFFFFFu) { // ties with the threshold response are kept
            const int j = atomicAdd(&s_n, 1);
            const uint32_t c = K[i];
            if (j < CV_SORT_CAP) s_e[j] = ((unsigned long long)((pslot_py(c) << 12) | pslot_px(c)) << 32) | (unsigned)i;
        }
    }
    __syncthreads();
    int n = s_n;
    if (n > CV_SORT_CAP) { // thousands of exactly equal responses: not a case the fixed on-chip sort covers
        if (tid == 0) atomicOr(err, PSLOT_FLAG_CVORB_CAP);
        n = CV_SORT_CAP;
    }
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = n + tid; i < m; i += 1024) s_e[i] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < m; i += 1024) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long a = s_e[i], b = s_e[p];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { s_e[i] = b; s_e[p] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < n; i += 1024) {
        const int src = (int)(s_e[i] & 0xFFFFFFFFu);
        final_kp[L.kp_base + i] = K[src];
        final_resp[L.kp_base + i] = R[src];
    }
    if (tid == 0) counters[CV_CNT_FINAL + l] = n;
}

// ---- orientation + float blur + steered BRIEF, one warp per keypoint -----------------------------------------------------------
constexpr int CW = 4;            // warps (keypoints) per CTA
constexpr int CPR = 21;          // raw patch radius: 15 (pattern reach after rotation: 18) + 3 (blur)
constexpr int CPD = 2 * CPR + 1; // 43
constexpr int CRW = 17, CRS = CRW * 4; // four 16-byte chunks cover the 43 bytes of a row at any alignment
constexpr int CBR = 18, CBD = 2 * CBR + 1; // blurred support 37 x 37
constexpr int CCS = 45;          // float column stride of the horizontal pass, [blurred column][patch row]

struct CvWarpSmem {
    uint32_t raw[CPD * CRW];
    float hp[CBD * CCS];
};

__global__ void __launch_bounds__(CW * 32)
k_cv_describe(const __grid_constant__ CvGeom g, const uint8_t *__restrict__ pyr, const uint32_t *__restrict__ final_kp,
              const float *__restrict__ final_resp, const int *__restrict__ counters, pslot_keypoint *__restrict__ out_kps,
              uint8_t *__restrict__ out_desc, int *__restrict__ out_count, int *__restrict__ err)
{
    __shared__ __align__(16) CvWarpSmem sm[CW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = blockIdx.y, slot = blockIdx.z;
    const CvLevel &L = g.lv[l];
    counters += slot * CV_CNT_TOTAL;
    pyr += slot * g.slot_bytes;
    final_kp += (size_t)slot * g.slot_kp; final_resp += (size_t)slot * g.slot_kp;
    out_kps += (size_t)slot * g.slot_kp; out_desc += (size_t)slot * g.slot_kp * 32;
    const int c_lane = lane < g.nlevels ? counters[CV_CNT_FINAL + lane] : 0;
    const int out_base = __reduce_add_sync(0xffffffffu, lane < l ? c_lane : 0);
    const int total = __reduce_add_sync(0xffffffffu, c_lane);
    if (blockIdx.x == 0 && l == 0 && threadIdx.x == 0) out_count[slot] = total;
    const int n_level = __shfl_sync(0xffffffffu, c_lane, l);
    const float k0 = g.gk[0], k1 = g.gk[1], k2 = g.gk[2], k3 = g.gk[3]; // k[4..6] mirror k[2..0]
    for (int i = blockIdx.x * CW + warp; i < n_level; i += gridDim.x * CW) { // warp-uniform

    const uint32_t kp = final_kp[L.kp_base + i];
    const int kx = pslot_px(kp), ky = pslot_py(kp);
    const uint8_t *img = pyr + L.off;
    CvWarpSmem &S = sm[warp];
    uint8_t *rawb = reinterpret_cast<uint8_t *>(S.raw);

    // the 43x43 patch around the keypoint; patch column c lives at byte c + shift of its shared row
    const int x0 = kx - CPR, y0 = ky - CPR;
    int shift;
    if (x0 >= 0 && y0 >= 0 && x0 + CPD <= L.w && y0 + CPD <= L.h) {
        // rows are 128-byte aligned in global memory: the four aligned 16-byte chunks that cover the 43 bytes
        shift = x0 & 15;
        const int xa = x0 - shift;
        const uint8_t *src = img + (size_t)y0 * L.pitch + xa;
        constexpr int NLD = (CPD * 4 + 31) / 32;
        uint4 v[NLD];
#pragma unroll
        for (int k = 0; k < NLD; ++k) { // all loads of a lane are issued before the first store
            const int e = lane + 32 * k;
            const int r = min(e >> 2, CPD - 1), c = e & 3;
            const int cc = xa + 16 * c < L.pitch ? c : 0; // a chunk beyond the row's pitch is never needed
            v[k] = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)r * L.pitch) + cc);
        }
#pragma unroll
        for (int k = 0; k < NLD; ++k) {
            const int e = lane + 32 * k;
            const int r = e >> 2, c = e & 3;
            if (r < CPD) {
                uint32_t *d = S.raw + r * CRW + 4 * c;
                d[0] = v[k].x; d[1] = v[k].y; d[2] = v[k].z; d[3] = v[k].w;
            }
        }
    } else { // within 21 px of the image edge (edgeThreshold 19 or 20): BORDER_REFLECT_101, byte by byte (the blur runs on the whole level)
        shift = 0;
        for (int e = lane; e < CPD * CPD; e += 32) {
            const int r = e / CPD, c = e - r * CPD;
            rawb[r * CRS + c] = __ldg(img + (size_t)reflect101(y0 + r, L.h) * L.pitch + reflect101(x0 + c, L.w));
        }
    }
    __syncwarp();

    // ICAngles: lane r owns disc row v = r - 15
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - 15;
        const int d = c_umax[v < 0 ? -v : v];
        const uint8_t *row = rawb + (CPR + v) * CRS + CPR + shift;
        int sum = 0;
        for (int u = -d; u <= d; ++u) {
            const int val = row[u];
            m10 += u * val;
            sum += val;
        }
        m01 = v * sum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // horizontal float pass of cv::sepFilter2D: taps left to right, fused multiply-adds.  A task is half a patch row (19 or
    // 18 blurred columns); its bytes are converted once and slide through a 7-register window (one load, one conversion and
    // seven multiply-adds per output instead of seven of each).
    for (int task = lane; task < 2 * CPD; task += 32) {
        const int half = task >= CPD ? 1 : 0, r = task - half * CPD;
        const int c0 = half * 19, nc = 19 - half;
        const uint8_t *p = rawb + r * CRS + shift + c0;
        float *o = S.hp + c0 * CCS + r;
        float f0 = (float)p[0], f1 = (float)p[1], f2 = (float)p[2], f3 = (float)p[3], f4 = (float)p[4], f5 = (float)p[5];
#pragma unroll
        for (int c = 0; c < 19; ++c) {
            if (c < nc) {
                const float f6 = (float)p[c + 6];
                float acc = __fmul_rn(k0, f0);
                acc = __fmaf_rn(k1, f1, acc);
                acc = __fmaf_rn(k2, f2, acc);
                acc = __fmaf_rn(k3, f3, acc);
                acc = __fmaf_rn(k2, f4, acc);
                acc = __fmaf_rn(k1, f5, acc);
                acc = __fmaf_rn(k0, f6, acc);
                o[c * CCS] = acc;
                f0 = f1; f1 = f2; f2 = f3; f3 = f4; f4 = f5; f5 = f6;
            }
        }
    }
    __syncwarp();

    // keypoint record (orb.cpp computeKeyPoints: pt * scale, size = patchSize * scale, response = Harris)
    const float scale = L.scale;
    const float fx = __fmul_rn((float)kx, scale), fy = __fmul_rn((float)ky, scale);
    // computeOrbDescriptors centres the pattern on cvRound(pt / scale-as-multiplication), which is the candidate pixel again
    const float inv = __fdiv_rn(1.f, scale);
    const int cx = __float2int_rn(__fmul_rn(fx, inv)), cy = __float2int_rn(__fmul_rn(fy, inv));
    if ((cx != kx || cy != ky) && lane == 0) atomicOr(err, PSLOT_FLAG_CVORB_CAP); // cannot happen below 4096 px

    float a, b;
    steer_coeffs(angle, a, b);
    auto sample = [&](float px, float py) -> int {
        const int r = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
        const int c = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
        // column taps from the centre outwards over the symmetric pairs
        const float *t = S.hp + (c + CBR) * CCS + (r + CPR);
        float acc = __fmul_rn(k3, t[0]);
        acc = __fmaf_rn(k2, __fadd_rn(t[-1], t[1]), acc);
        acc = __fmaf_rn(k1, __fadd_rn(t[-2], t[2]), acc);
        acc = __fmaf_rn(k0, __fadd_rn(t[-3], t[3]), acc);
        return min(max(__float2int_rn(acc), 0), 255);
    };
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const uint32_t p = __ldg(&g_pat.w[t * 32 + lane]);
        const int t0 = sample((float)(int8_t)(p & 0xFF), (float)(int8_t)((p >> 8) & 0xFF));
        const int t1 = sample((float)(int8_t)((p >> 16) & 0xFF), (float)(int8_t)(p >> 24));
        val |= (t0 < t1) << t;
    }
    const size_t o = (size_t)out_base + i;
    out_desc[o * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        pslot_keypoint k;
        k.x = fx; k.y = fy;
        k.size = __fmul_rn(31.f, scale);
        k.angle = angle;
        k.response = final_resp[L.kp_base + i];
        k.octave = l;
        k.class_id = -1;
        out_kps[o] = k;
    }
    __syncwarp(); // the next keypoint's patch overwrites this warp's shared memory
    }
}

} // namespace

int launch_cvorb(const CvGeom &g, const CvBuffers &b, int slots, bool has_mask, cudaStream_t st)
{
    int n = 0;
    const int planes = has_mask ? 2 : 1;
    for (int l = 1; l < g.nlevels; ++l) { // each level from the previous one (orb.cpp: resize(prevImg, currImg, INTER_LINEAR_EXACT))
        const CvLevel &S = g.lv[l - 1], &D = g.lv[l];
        dim3 grd((D.w + 127) / 128, (D.h + 7) / 8, planes * slots);
        k_cv_resize_exact<<<grd, 256, 0, st>>>(b.pyr + S.off, b.mask + S.off, S.w, S.h, S.pitch, b.pyr + D.off, b.mask + D.off,
                                               D.w, D.h, D.pitch, b.taps + D.tab_x, b.taps + D.tab_y, (size_t)g.slot_bytes, planes);
        ++n;
    }
    cudaMemsetAsync(b.counters, 0, sizeof(int) * CV_CNT_TOTAL * slots, st);
    cudaMemsetAsync(b.hist, 0, sizeof(int) * 256 * PSLOT_MAX_LEVELS * slots, st);
    k_cv_fast_score<<<dim3(g.n_tiles, slots), 256, 0, st>>>(g, b.pyr, b.score);
    k_cv_nms_collect<<<dim3(g.n_tiles, slots), 256, 0, st>>>(g, b.score, has_mask ? b.mask : nullptr, b.cand, b.counters, b.hist,
                                                                b.err);
    k_cv_harris<<<dim3(slots > 8 ? 8 : 32, g.nlevels, slots), 256, 0, st>>>(g, b.pyr, b.cand, b.counters, b.hist, b.kept, b.kept_resp);
    k_cv_select<<<dim3(g.nlevels, slots), 1024, 0, st>>>(g, b.kept, b.kept_resp, b.counters, b.final_kp, b.final_resp, b.err);
    k_cv_describe<<<dim3(g.describe_blocks, g.nlevels, slots), CW * 32, 0, st>>>(g, b.pyr, b.final_kp, b.final_resp, b.counters,
                                                                               b.out_kps, b.out_desc, b.out_count, b.err);
    return n + 5;
}
