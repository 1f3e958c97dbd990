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
// Kernels (one image per call; every kernel covers all pyramid levels, blockIdx.y/z = level, except the resize chain):
//   k_cv_resize_exact   level l from level l-1 (image and mask)                      HBM stream, tiny
//   k_cv_fast_score     raw FAST score map of every level (0 where not a corner)     integer ALU
//   k_cv_nms_collect    strict 3x3 maxima + mask + border -> candidate list + score histogram per level
//   k_cv_harris         retainBest(2q) threshold from the histogram; Harris response of the survivors
//   k_cv_select         per level: radix select of the quota-th largest Harris response, ties kept, bitonic sort by (y, x)
//   k_cv_describe       one warp per keypoint: orientation, float blur of the 43x43 patch, 256 tests
#include "pslot_internal.h"
#include "orb_common.cuh"

namespace {

// ---- pyramid --------------------------------------------------------------------------------------------------------------
// blockIdx.z = 0: the image, 1: the object mask (same taps; thresholded)
__global__ void k_cv_resize_exact(const uint8_t *__restrict__ src_img, const uint8_t *__restrict__ src_mask, int sw, int sh,
                                  int spitch, uint8_t *__restrict__ dst_img, uint8_t *__restrict__ dst_mask, int dw, int dh,
                                  int dpitch, const int2 *__restrict__ xt, const int2 *__restrict__ yt)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const bool is_mask = blockIdx.z != 0;
    const uint8_t *src = is_mask ? src_mask : src_img;
    const int2 tx = xt[x], ty = yt[y];
    const int xa = tx.x, xb = min(xa + 1, sw - 1), ax = tx.y;
    const int ya = ty.x, yb = min(ya + 1, sh - 1), ay = ty.y;
    const uint8_t *r0 = src + (size_t)ya * spitch, *r1 = src + (size_t)yb * spitch;
    const int H0 = (int)r0[xa] * (256 - ax) + (int)r0[xb] * ax;
    const int H1 = (int)r1[xa] * (256 - ax) + (int)r1[xb] * ax;
    int v = (H0 * (256 - ay) + H1 * ay + (1 << 15)) >> 16;
    if (is_mask) v = v > 254 ? v : 0; // cv::threshold(mask, mask, 254, 0, THRESH_TOZERO)
    (is_mask ? dst_mask : dst_img)[(size_t)y * dpitch + x] = (uint8_t)v;
}

// ---- FAST score map ---------------------------------------------------------------------------------------------------------
// R = max(max_k min_{j=k..k+8} (v - p_j), max_k min_{j=k..k+8} (p_j - v)) - 1; corner at threshold t <=> R >= t
// (cv::FAST 9-16, cornerScore<16>).  Only pixels >= 3 px inside the level are tested.
__global__ void __launch_bounds__(256)
k_cv_fast_score(const __grid_constant__ CvGeom g, const uint8_t *__restrict__ pyr, uint8_t *__restrict__ score)
{
    const int l = blockIdx.z;
    const CvLevel &L = g.lv[l];
    const int x = blockIdx.x * 32 + (threadIdx.x & 31) + 3, y = blockIdx.y * 8 + (threadIdx.x >> 5) + 3;
    if (x >= L.w - 3 || y >= L.h - 3) return;
    const uint8_t *p = pyr + L.off + (size_t)y * L.pitch + x;
    const int P = L.pitch, v = p[0], t = g.fast_th;
    uint8_t out = 0;
    // a 9-arc holds one pixel of every opposite pair: both compass pairs must have a member farther than t from v
    const int d0 = abs((int)p[3 * P] - v), d8 = abs((int)p[-3 * P] - v), d4 = abs((int)p[3] - v), d12 = abs((int)p[-3] - v);
    if ((d0 > t || d8 > t) && (d4 > t || d12 > t)) {
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
        const int R = max(hi, -lo) - 1;
        if (R >= t) out = (uint8_t)R;
    }
    score[L.off + (size_t)y * L.pitch + x] = out;
}

// strict 3x3 maxima (pixels outside the tested range hold 0 in the score map), runByPixelsMask, runByImageBorder;
// survivors go to the level's candidate list (unordered: everything downstream is a set operation until the final sort)
__global__ void __launch_bounds__(256)
k_cv_nms_collect(const __grid_constant__ CvGeom g, const uint8_t *__restrict__ score, const uint8_t *__restrict__ mask,
                 uint32_t *__restrict__ cand, int *__restrict__ counters, int *__restrict__ hist, int *__restrict__ err)
{
    const int l = blockIdx.z;
    const CvLevel &L = g.lv[l];
    const int x = blockIdx.x * 32 + (threadIdx.x & 31) + 3, y = blockIdx.y * 8 + (threadIdx.x >> 5) + 3;
    if (x >= L.w - 3 || y >= L.h - 3) return;
    const int P = L.pitch;
    const uint8_t *s = score + L.off + (size_t)y * P + x;
    const int c = s[0];
    if (c == 0) return;
    if (!(c > s[-1] && c > s[1] && c > s[-P - 1] && c > s[-P] && c > s[-P + 1] && c > s[P - 1] && c > s[P] && c > s[P + 1])) return;
    if (mask && mask[L.off + (size_t)y * P + x] == 0) return;
    if (x < g.edge || x >= L.w - g.edge || y < g.edge || y >= L.h - g.edge) return;
    const int i = atomicAdd(&counters[CV_CNT_CAND + l], 1);
    if (i >= L.cand_cap) { atomicOr(err, PSLOT_FLAG_CVORB_CAP); return; }
    cand[L.cand_base + i] = pslot_pack(x, y, c);
    atomicAdd(&hist[l * 256 + c], 1);
}

// ---- retainBest(2 * quota) on the FAST score, then HarrisResponses -----------------------------------------------------------
__global__ void __launch_bounds__(256)
k_cv_harris(const __grid_constant__ CvGeom g, const uint8_t *__restrict__ pyr, const uint32_t *__restrict__ cand,
            int *__restrict__ counters, const int *__restrict__ hist, uint32_t *__restrict__ kept, float *__restrict__ kept_resp)
{
    __shared__ int s_thr;
    const int l = blockIdx.y;
    const CvLevel &L = g.lv[l];
    const int n = min(counters[CV_CNT_CAND + l], L.cand_cap);
    if (threadIdx.x == 0) {
        // KeyPointsFilter::retainBest(keypoints, 2 * quota): keep everything >= the response of the (2 * quota)-th best
        const int keep = 2 * L.quota;
        int thr = 0;
        if (n > keep) {
            int acc = 0;
            for (int sc = 255; sc >= 0; --sc) {
                acc += hist[l * 256 + sc];
                if (acc >= keep) { thr = sc; break; }
            }
        }
        s_thr = keep == 0 ? 256 : thr;
    }
    __syncthreads();
    const int thr = s_thr;
    const float harris_k = 0.04f;
    const float scale1 = 1.f / ((1 << 2) * 7 * 255.f);
    const float scale_sq_sq = __fmul_rn(__fmul_rn(__fmul_rn(scale1, scale1), scale1), scale1);
    const int P = L.pitch;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t c = cand[L.cand_base + i];
        if (pslot_pr(c) < thr) continue;
        const int x0 = pslot_px(c), y0 = pslot_py(c);
        const uint8_t *ctr = pyr + L.off + (size_t)y0 * P + x0;
        int a = 0, b = 0, cc = 0;
        for (int dy = -3; dy <= 3; ++dy) {
            const uint8_t *q = ctr + dy * P;
#pragma unroll
            for (int dx = -3; dx <= 3; ++dx) {
                const uint8_t *r = q + dx;
                const int Ix = ((int)r[1] - r[-1]) * 2 + ((int)r[-P + 1] - r[-P - 1]) + ((int)r[P + 1] - r[P - 1]);
                const int Iy = ((int)r[P] - r[-P]) * 2 + ((int)r[P - 1] - r[-P - 1]) + ((int)r[P + 1] - r[-P + 1]);
                a += Ix * Ix; b += Iy * Iy; cc += Ix * Iy;
            }
        }
        const float fa = (float)a, fb = (float)b, fc = (float)cc;
        const float tr = __fadd_rn(fa, fb);
        const float resp = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(harris_k, tr), tr)),
                                     scale_sq_sq);
        const int j = atomicAdd(&counters[CV_CNT_KEPT + l], 1); // <= cand count <= cand_cap: the kept arrays share the capacity
        kept[L.cand_base + j] = c;
        kept_resp[L.cand_base + j] = resp;
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
    __shared__ int s_k, s_n;
    const int l = blockIdx.x, tid = threadIdx.x;
    const CvLevel &L = g.lv[l];
    const int n1 = counters[CV_CNT_KEPT + l];
    const uint32_t *K = kept + L.cand_base;
    const float *R = kept_resp + L.cand_base;
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
            if (tid == 0) {
                int acc = 0, k = s_k, d = 255;
                for (; d > 0; --d) {
                    if (acc + s_hist[d] >= k) break;
                    acc += s_hist[d];
                }
                s_k = k - acc;
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
constexpr int CRW = 13, CRS = CRW * 4;
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
    const int l = blockIdx.y;
    const CvLevel &L = g.lv[l];
    const int c_lane = lane < g.nlevels ? counters[CV_CNT_FINAL + lane] : 0;
    const int out_base = __reduce_add_sync(0xffffffffu, lane < l ? c_lane : 0);
    const int total = __reduce_add_sync(0xffffffffu, c_lane);
    if (blockIdx.x == 0 && l == 0 && threadIdx.x == 0) *out_count = total;
    const int i = blockIdx.x * CW + warp;
    if (i >= __shfl_sync(0xffffffffu, c_lane, l)) return;

    const uint32_t kp = final_kp[L.kp_base + i];
    const int kx = pslot_px(kp), ky = pslot_py(kp);
    const uint8_t *img = pyr + L.off;
    CvWarpSmem &S = sm[warp];
    uint8_t *rawb = reinterpret_cast<uint8_t *>(S.raw);

    // the 43x43 patch around the keypoint, BORDER_REFLECT_101 at the image edge (the blur runs on the whole level)
    const int x0 = kx - CPR, y0 = ky - CPR;
    for (int e = lane; e < CPD * CPD; e += 32) {
        const int r = e / CPD, c = e - r * CPD;
        rawb[r * CRS + c] = __ldg(img + (size_t)reflect101(y0 + r, L.h) * L.pitch + reflect101(x0 + c, L.w));
    }
    __syncwarp();

    // ICAngles: lane r owns disc row v = r - 15
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - 15;
        const int d = c_umax[v < 0 ? -v : v];
        const uint8_t *row = rawb + (CPR + v) * CRS + CPR;
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

    // horizontal float pass of cv::sepFilter2D: taps left to right, fused multiply-adds
    const float k0 = g.gk[0], k1 = g.gk[1], k2 = g.gk[2], k3 = g.gk[3]; // k[4..6] mirror k[2..0]
    for (int r = lane; r < CPD; r += 32) {
        const uint8_t *p = rawb + r * CRS;
        for (int c = 0; c < CBD; ++c) {
            float acc = __fmul_rn(k0, (float)p[c]);
            acc = __fmaf_rn(k1, (float)p[c + 1], acc);
            acc = __fmaf_rn(k2, (float)p[c + 2], acc);
            acc = __fmaf_rn(k3, (float)p[c + 3], acc);
            acc = __fmaf_rn(k2, (float)p[c + 4], acc);
            acc = __fmaf_rn(k1, (float)p[c + 5], acc);
            acc = __fmaf_rn(k0, (float)p[c + 6], acc);
            S.hp[c * CCS + r] = acc;
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
}

} // namespace

int launch_cvorb(const CvGeom &g, const CvBuffers &b, bool has_mask, cudaStream_t st)
{
    int n = 0;
    for (int l = 1; l < g.nlevels; ++l) { // each level from the previous one (orb.cpp: resize(prevImg, currImg, INTER_LINEAR_EXACT))
        const CvLevel &S = g.lv[l - 1], &D = g.lv[l];
        dim3 blk(32, 8), grd((D.w + 31) / 32, (D.h + 7) / 8, has_mask ? 2 : 1);
        k_cv_resize_exact<<<grd, blk, 0, st>>>(b.pyr + S.off, b.mask + S.off, S.w, S.h, S.pitch, b.pyr + D.off, b.mask + D.off,
                                               D.w, D.h, D.pitch, b.taps + D.tab_x, b.taps + D.tab_y);
        ++n;
    }
    cudaMemsetAsync(b.counters, 0, sizeof(int) * CV_CNT_TOTAL, st);
    cudaMemsetAsync(b.hist, 0, sizeof(int) * 256 * PSLOT_MAX_LEVELS, st);
    const CvLevel &L0 = g.lv[0];
    dim3 grd((L0.w - 6 + 31) / 32, (L0.h - 6 + 7) / 8, g.nlevels);
    k_cv_fast_score<<<grd, 256, 0, st>>>(g, b.pyr, b.score);
    k_cv_nms_collect<<<grd, 256, 0, st>>>(g, b.score, has_mask ? b.mask : nullptr, b.cand, b.counters, b.hist, b.err);
    k_cv_harris<<<dim3(32, g.nlevels), 256, 0, st>>>(g, b.pyr, b.cand, b.counters, b.hist, b.kept, b.kept_resp);
    k_cv_select<<<g.nlevels, 1024, 0, st>>>(g, b.kept, b.kept_resp, b.counters, b.final_kp, b.final_resp, b.err);
    k_cv_describe<<<dim3((CV_SORT_CAP + CW - 1) / CW, g.nlevels), CW * 32, 0, st>>>(g, b.pyr, b.final_kp, b.final_resp, b.counters,
                                                                                  b.out_kps, b.out_desc, b.out_count, b.err);
    return n + 5;
}
