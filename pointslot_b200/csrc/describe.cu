// describe.cu — K4+K5+K6 fused: IC_Angle orientation (reference src/ORBextractor.cc:77-104, 472-479),
// GaussianBlur 7x7 sigma 2 (:1085-1086) and the steered-BRIEF descriptor (:107-147, 1034-1041), plus
// the output assembly (:834-847, 1095-1103).  One warp per keypoint.
//
// The reference blurs every whole level and then gathers 512 samples per keypoint.  Here each warp
// stages the 43x43 raw patch around its keypoint in shared memory once (37x37 blurred support,
// radius 18 = max |rotated pattern coordinate|, plus the 3-px blur halo) with aligned 32-bit loads,
// and derives from it
//   * the intensity-centroid moments over the radius-15 disc (integer, exact) -> fastAtan2 angle,
//   * the horizontal half of the separable fixed-point blur [18 34 48 56 48 34 18]/256 for all 43 rows,
//     two rows per lane packed 16x2 in one 32-bit register (each half stays <= 65280, so plain 32-bit
//     multiply-adds never carry across the halves),
//   * the vertical half only at the 512 sampled positions, fused into the pattern tests with the single
//     final rounding (v + 2^15) >> 16; lane i produces descriptor byte i.
// BORDER_REFLECT_101 is applied at the true image edge while staging.  No blurred level is ever written
// to or read from global memory.
//
// Float islands follow DESIGN.md parity rule 3: single-rounded __f*_rn ops (no FMA), cosf/sinf by glibc's own
// algorithm (glibc_sincosf below), cvRound = round-half-even.
//
// Bound: shared memory / integer ALU; algorithmic global bytes per keypoint = 1849 (patch) + 4 (lkp)
// + 28 + 32 (outputs).
#include "pslot_internal.h"
#include "orb_common.cuh"

namespace {

constexpr int DW = 4;           // warps (keypoints) per CTA
constexpr int PR = 21;          // raw patch radius
constexpr int PD = 2 * PR + 1;  // 43
constexpr int RW = 13;          // raw row stride in words (52 bytes): odd -> conflict-free row-parallel access
constexpr int RS = RW * 4;
constexpr int NW = 12;          // aligned words fetched per patch row (covers 43 bytes at any alignment)
constexpr int BR = 18;          // blurred support radius
constexpr int BD = 2 * BR + 1;  // 37
constexpr int CS = 44;          // h-pass column stride (u16): the h-pass is stored column-major, [blurred column][patch row]

struct WarpSmem {
    uint32_t raw[PD * RW];  // patch bytes, row stride 52; column c of the patch is byte c + shift
    uint32_t hp[BD * CS / 2]; // horizontal pass as u16 [blurred column][patch row]: the 7 vertical taps of a sample are contiguous
};

__global__ void k_debug_steer(const float *__restrict__ angle_deg, int n, float *__restrict__ a, float *__restrict__ b)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) steer_coeffs(angle_deg[i], a[i], b[i]);
}

__global__ void __launch_bounds__(DW * 32, 9) // 56 registers: 9 CTAs per SM measured best (8: 565 us, 9: 552 us, 10: 568 us)
k_describe(const __grid_constant__ Geom g, const __grid_constant__ PyrPtrs pyr,
           const uint32_t *__restrict__ lkp, const int *__restrict__ lkp_count,
           pslot_keypoint *__restrict__ out_kps, uint8_t *__restrict__ out_desc, int *__restrict__ out_count)
{
    __shared__ __align__(16) WarpSmem sm[DW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = blockIdx.y;

    // which keypoint: global index over the per-level capacity segments
    int idx = blockIdx.x * DW + warp;
    if (idx >= g.lkp_per_slot) return;
    // level of this keypoint slot and the output offset of its level, lane-parallel over the (<= 16) levels: lane k
    // looks at level k, a ballot / two reductions replace the dependent scans
    const bool at_or_below = lane < g.nlevels && idx >= g.lv[min(lane, g.nlevels - 1)].kp_base; // kp_base ascending, [0] = 0
    const int l = __popc(__ballot_sync(0xffffffffu, at_or_below)) - 1;
    const LevelGeom &L = g.lv[l];
    const int i = idx - L.kp_base;
    const int *cnt = lkp_count + (size_t)slot * g.nlevels;
    const int c_lane = lane < g.nlevels ? cnt[lane] : 0;
    const int out_base = __reduce_add_sync(0xffffffffu, lane < l ? c_lane : 0);
    const int total = __reduce_add_sync(0xffffffffu, c_lane);
    if (idx == 0 && lane == 0) out_count[slot] = total;
    if (i >= __shfl_sync(0xffffffffu, c_lane, l)) return;

    const uint32_t kp = lkp[(size_t)slot * g.lkp_per_slot + idx];
    const int kx = pslot_px(kp), ky = pslot_py(kp), kr = pslot_pr(kp);
    const uint8_t *img = pyr.lv[l] + (size_t)slot * pyr.slot_stride[l];
    WarpSmem &S = sm[warp];
    uint8_t *rawb = reinterpret_cast<uint8_t *>(S.raw);

    // ---- stage the 43x43 raw patch ------------------------------------------------------------------
    const int x0 = kx - PR, y0 = ky - PR;
    int shift; // patch column c lives at byte c + shift of its shared row
    if (x0 >= 0 && y0 >= 0 && x0 + PD <= L.w && y0 + PD <= L.h) {
        // rows are 4-byte aligned in global memory (pitch % 128 == 0): fetch 12 aligned words per row
        shift = x0 & 3;
        const uint8_t *src = img + (size_t)y0 * L.pitch + (x0 - shift);
        // all 17 loads of a lane are issued before the first store (memory-level parallelism)
        constexpr int NLD = (PD * NW + 31) / 32;
        uint32_t v[NLD];
#pragma unroll
        for (int k = 0; k < NLD; ++k) {
            const int e = min(lane + 32 * k, PD * NW - 1);
            const int r = e / NW, wc = e - r * NW;
            v[k] = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)r * L.pitch) + wc);
        }
#pragma unroll
        for (int k = 0; k < NLD; ++k) {
            const int e = lane + 32 * k;
            const int r = e / NW, wc = e - r * NW;
            if (e < PD * NW) S.raw[r * RW + wc] = v[k];
        }
    } else { // within 21 px of the image edge: BORDER_REFLECT_101, byte by byte
        shift = 0;
        for (int e = lane; e < PD * PD; e += 32) {
            const int r = e / PD, c = e - r * PD;
            rawb[r * RS + c] = __ldg(img + (size_t)reflect101(y0 + r, L.h) * L.pitch + reflect101(x0 + c, L.w));
        }
    }
    __syncwarp();

    // ---- IC_Angle: lane r owns disc row v = r - 15 -------------------------------------------------
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - PSLOT_HALF_PATCH;
        const int d = c_umax[v < 0 ? -v : v];
        const uint8_t *row = rawb + (PR + v) * RS + PR + shift;
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

    // ---- horizontal blur pass: lane j owns patch rows 2j and 2j + 1 ------------------------------------
    // Each row is read as 12 words and re-aligned by the warp-uniform byte offset, so that patch column c is byte c;
    // blurred column c = sum_i k[i] * byte(c + i) is then 2-3 IDP.4A (u8 x u8 dot products on the FMA pipe) against
    // the 7 coefficients placed at byte offset c & 3 of a 12-byte window - no byte extraction, no shifts per output.
    // Sums stay <= 255 * 256 = 65280, so the two rows of a lane pack into one word = one (column, row pair) entry of
    // the column-major h-pass array: one conflict-free 32-bit store per column.
    if (lane < 22) {
        const int ra = 2 * lane, rb = min(2 * lane + 1, PD - 1);
        const uint32_t *wa = S.raw + ra * RW, *wb = S.raw + rb * RW;
        const int sh = 8 * shift;
        uint32_t A[11], B[11];
        {
            uint32_t ta[NW], tb[NW];
#pragma unroll
            for (int k = 0; k < NW; ++k) { ta[k] = wa[k]; tb[k] = wb[k]; }
#pragma unroll
            for (int k = 0; k < 11; ++k) { A[k] = __funnelshift_r(ta[k], ta[k + 1], sh); B[k] = __funnelshift_r(tb[k], tb[k + 1], sh); }
        }
        uint32_t *hw = S.hp + lane;
#pragma unroll
        for (int c = 0; c < BD; ++c) {
            constexpr uint32_t K0[4] = {0x38302212u, 0x30221200u, 0x22120000u, 0x12000000u};
            constexpr uint32_t K1[4] = {0x00122230u, 0x12223038u, 0x22303830u, 0x30383022u};
            constexpr uint32_t K2[4] = {0u, 0u, 0x00000012u, 0x00001222u};
            const int k = c >> 2, s = c & 3;
            uint32_t ha = __dp4a(A[k], K0[s], 0u), hb = __dp4a(B[k], K0[s], 0u);
            ha = __dp4a(A[k + 1], K1[s], ha); hb = __dp4a(B[k + 1], K1[s], hb);
            if (s >= 2) { ha = __dp4a(A[k + 2], K2[s], ha); hb = __dp4a(B[k + 2], K2[s], hb); }
            hw[c * (CS / 2)] = hb * 65536u + ha;
        }
    }
    __syncwarp();

    // ---- steered BRIEF with the vertical blur pass at the sampled positions (:107-147) ---------------
    float a, b;
    steer_coeffs(angle, a, b);
    const uint32_t *hp = S.hp;
    auto sample = [&](float px, float py) -> int {
        // cvRound (round half to even) by the 1.5 * 2^23 trick: the sum's low mantissa bits are the integer, on the FMA
        // pipe instead of the quarter-rate conversion unit; the bias folds into the constant below
        const float MAGIC = 12582912.f;
        const int rb = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)), MAGIC));
        const int cb = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)), MAGIC));
        // taps = patch rows (r + 18) .. (r + 24) of blurred column c + 18: 7 contiguous u16 inside 4 words
        const int base = cb * CS + rb + (BR * CS + BR - 0x4B400000 * (CS + 1));
        const uint32_t *w = hp + (base >> 1);
        const int sh = (base & 1) * 16;
        // IDP.2A: two 16-bit taps x two 8-bit coefficients per instruction; an odd first row shifts the coefficient pairs
        // by one tap (upper halves of the constants) instead of shifting the data
        uint32_t v = __dp2a_lo(w[0], 0x12002212u >> sh, 32768u);  // even: (18, 34)   odd: (0, 18)
        v = __dp2a_lo(w[1], 0x30223830u >> sh, v);                // even: (48, 56)   odd: (34, 48)
        v = __dp2a_lo(w[2], 0x30382230u >> sh, v);                // even: (48, 34)   odd: (56, 48)
        v = __dp2a_lo(w[3], 0x12220012u >> sh, v);                // even: (18, 0)    odd: (34, 18)
        return (int)(v >> 16);
    };
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const float4 p = __ldg(&g_patf.v[t * 32 + lane]);
        const int t0 = sample(p.x, p.y);
        const int t1 = sample(p.z, p.w);
        val |= (t0 < t1) << t;
    }
    const size_t o = (size_t)slot * g.kp_cap_total + out_base + i;
    out_desc[o * 32 + lane] = (uint8_t)val;

    // ---- keypoint record (:834-847, 1095-1103) ------------------------------------------------------
    if (lane == 0) {
        pslot_keypoint k;
        k.x = (float)kx; k.y = (float)ky;
        if (l != 0) { k.x = __fmul_rn(k.x, L.scale); k.y = __fmul_rn(k.y, L.scale); }
        k.size = (float)L.patch_size;
        k.angle = angle;
        k.response = (float)kr;
        k.octave = l;
        k.class_id = -1;
        out_kps[o] = k;
    }
}

} // namespace

void launch_debug_steer(const float *angle_deg, int n, float *a, float *b, cudaStream_t st)
{
    k_debug_steer<<<(n + 255) / 256, 256, 0, st>>>(angle_deg, n, a, b);
}

int launch_describe(const Geom &g, const PyrPtrs &pyr, const uint32_t *lkp, const int *lkp_count,
                    pslot_keypoint *out_kps, uint8_t *out_desc, int *out_count, int nslots, cudaStream_t st)
{
    dim3 grid((g.lkp_per_slot + DW - 1) / DW, nslots);
    k_describe<<<grid, DW * 32, 0, st>>>(g, pyr, lkp, lkp_count, out_kps, out_desc, out_count);
    return 1;
}
