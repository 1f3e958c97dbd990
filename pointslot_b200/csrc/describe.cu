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

#ifndef DESC_WARPS
#define DESC_WARPS 4
#endif
#ifndef DESC_MIN_BLOCKS
#define DESC_MIN_BLOCKS (36 / DESC_WARPS)
#endif
constexpr int DW = DESC_WARPS;  // warps (keypoints) per CTA
constexpr int PR = 21;          // raw patch radius
constexpr int PD = 2 * PR + 1;  // 43
constexpr int RW = 17;          // raw row stride in words (68 bytes = 4 aligned 16-byte chunks + 1): odd -> conflict-free row-parallel access
constexpr int RS = RW * 4;
constexpr int NW = 12;          // words of a patch row the blur reads (43 bytes at any byte offset inside the first word)
constexpr int BR = 18;          // blurred support radius
constexpr int BD = 2 * BR + 1;  // 37
constexpr int CS = 44;          // h-pass column stride (u16): the h-pass is stored column-major, [blurred column][patch row]

struct WarpSmem {
    // horizontal pass as u16 [blurred column][patch row]: the 7 vertical taps of a sample are contiguous (+2 words: the dense
    // vertical pass reads one word past the last column); first, so that every column starts 8-byte aligned
    uint32_t hp[BD * CS / 2 + 2];
    uint32_t raw[PD * RW + 1]; // patch bytes, row stride 68; column c of the patch is byte c + shift (+1: size % 16 == 0)
};
static_assert(sizeof(WarpSmem) % 16 == 0 && (BD * CS / 2 + 2) % 2 == 0, "8-byte accesses into hp / raw");

__global__ void k_debug_steer(const float *__restrict__ angle_deg, int n, float *__restrict__ a, float *__restrict__ b)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) steer_coeffs(angle_deg[i], a[i], b[i]);
}

__global__ void __launch_bounds__(DW * 32, DESC_MIN_BLOCKS) // 56 registers: 9 CTAs per SM measured best (8: 565 us, 9: 552 us, 10: 568 us)
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
        // rows are 128-byte aligned in global memory: fetch the four aligned 16-byte chunks that cover the 43 bytes at any
        // alignment (6 wide loads per lane instead of 17 word loads, index arithmetic by shifts only)
        shift = x0 & 15;
        const int xa = x0 - shift;
        const uint8_t *src = img + (size_t)y0 * L.pitch + xa;
        constexpr int NLD = (PD * 4 + 31) / 32;
        uint4 v[NLD];
#pragma unroll
        for (int k = 0; k < NLD; ++k) { // all loads of a lane are issued before the first store (memory-level parallelism)
            const int e = lane + 32 * k;
            const int r = min(e >> 2, PD - 1), c = e & 3;
            // the last chunk may start beyond the row's pitch when the patch ends near the right edge: never needed then
            const int cc = xa + 16 * c < L.pitch ? c : 0;
            v[k] = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)r * L.pitch) + cc);
        }
#pragma unroll
        for (int k = 0; k < NLD; ++k) {
            const int e = lane + 32 * k;
            const int r = e >> 2, c = e & 3;
            if (r < PD) {
                uint32_t *d = S.raw + r * RW + 4 * c;
                d[0] = v[k].x; d[1] = v[k].y; d[2] = v[k].z; d[3] = v[k].w;
            }
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
#ifdef DESC_BYTE_MOMENTS
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
#else
    if (lane < 31) {
        // The row's 32 bytes u = -16 .. 15 as eight re-aligned words; the disc row |u| <= d is a run of ones in a 32-bit
        // mask (one bit per byte), spread to byte masks nibble by nibble with a carry-free multiply, so that the two row
        // sums are IDP.4A dot products with the masked weights (u resp. 1) - no byte loads, no data-dependent loop.
        const int v = lane - PSLOT_HALF_PATCH;
        const int d = c_umax[v < 0 ? -v : v];
        const int off = PR - 16 + shift; // byte offset of u = -16 inside the shared row
        const uint32_t *w = S.raw + (PR + v) * RW + (off >> 2);
        const int sh = 8 * (off & 3);
        const uint32_t bits = ((2u << (2 * d)) - 1u) << (16 - d);
        uint32_t t[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) t[k] = w[k];
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t x = __funnelshift_r(t[k], t[k + 1], sh);
            const uint32_t ones = (((bits >> (4 * k)) & 0xFu) * 0x00204081u) & 0x01010101u;
            // u of the word's four bytes: 4k - 16 + {0, 1, 2, 3}, as signed bytes
            const uint32_t ub = (uint32_t)((4 * k - 16) & 0xFF) * 0x01010101u + 0x03020100u;
            sum = __dp4a(x, ones, sum);
            int m;
            asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(m) : "r"(x), "r"((ones * 0xFFu) & ub), "r"(m10));
            m10 = m;
        }
        m01 = v * (int)sum;
    }
#endif
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
        const uint32_t *wa = S.raw + ra * RW + (shift >> 2), *wb = S.raw + rb * RW + (shift >> 2);
        const int sh = 8 * (shift & 3);
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
    const float MAGIC = 12582912.f; // 1.5 * 2^23: (x + MAGIC)'s low mantissa bits are cvRound(x) (round half to even), on the
                                    // FMA pipe instead of the quarter-rate conversion unit; the bias folds into a constant
#ifndef DESC_GATHER_V
    // ---- vertical blur pass, dense: the 37 x 37 blurred support as bytes, column-major (stride BS), over the raw patch
    // (no longer needed).  A unit = 8 consecutive rows of one column: 14 taps rows = 7 words of the column-major h-pass,
    // 4 IDP.2A per output with compile-time even / odd coefficient pairs; 185 units over the 32 lanes.  The 512 samples
    // are then single byte loads: the random 4-word gathers of the fused variant cost 4.4 shared-memory wavefronts each
    // (ncu), which made the shared-memory pipe the kernel's bound.
    constexpr int BS = 40;
    {
        uint8_t *blur = rawb;
        for (int u = lane; u < BD * 5; u += 32) {
            const int c = (int)(((unsigned)u * 13108u) >> 16), t = u - 5 * c; // u / 5 for u < 185
            const uint2 *src = reinterpret_cast<const uint2 *>(S.hp + c * (CS / 2) + 4 * t);
            const uint2 q0 = src[0], q1 = src[1], q2 = src[2], q3 = src[3];
            const uint32_t w[8] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y};
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = i >> 1;
                uint32_t v;
                if ((i & 1) == 0) { // taps start on a word boundary: (18,34) (48,56) (48,34) (18,-)
                    v = __dp2a_lo(w[k], 0x2212u, 32768u); v = __dp2a_lo(w[k + 1], 0x3830u, v);
                    v = __dp2a_lo(w[k + 2], 0x2230u, v);  v = __dp2a_lo(w[k + 3], 0x0012u, v);
                } else {            // one tap later: (-,18) (34,48) (56,48) (34,18)
                    v = __dp2a_lo(w[k], 0x1200u, 32768u); v = __dp2a_lo(w[k + 1], 0x3022u, v);
                    v = __dp2a_lo(w[k + 2], 0x3038u, v);  v = __dp2a_lo(w[k + 3], 0x1222u, v);
                }
                o[i] = v; // < 2^24: the blurred value is byte 2
            }
            uint2 out;
            out.x = __byte_perm(__byte_perm(o[0], o[1], 0x0062), __byte_perm(o[2], o[3], 0x0062), 0x5410);
            out.y = __byte_perm(__byte_perm(o[4], o[5], 0x0062), __byte_perm(o[6], o[7], 0x0062), 0x5410);
            *reinterpret_cast<uint2 *>(blur + c * BS + 8 * t) = out;
        }
    }
    __syncwarp();
    auto sample = [&](float px, float py) -> int {
        const int rb = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)), MAGIC));
        const int cb = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)), MAGIC));
        return rawb[cb * BS + rb + (int)((uint32_t)(BR * BS + BR) - 0x4B400000u * (uint32_t)(BS + 1))];
    };
#else
    const uint32_t *hp = S.hp;
    auto sample = [&](float px, float py) -> int {
        const int rb = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)), MAGIC));
        const int cb = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)), MAGIC));
        // taps = patch rows (r + 18) .. (r + 24) of blurred column c + 18: 7 contiguous u16 inside 4 words
        const int base = cb * CS + rb + (int)((uint32_t)(BR * CS + BR) - 0x4B400000u * (uint32_t)(CS + 1));
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
#endif
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
