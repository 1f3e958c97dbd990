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

    // ---- horizontal blur pass: lane owns rows (lane, lane + 22), packed 16x2 --------------------------
    if (lane < 22) {
        const int ra = lane, rb = min(lane + 22, PD - 1);
        const uint8_t *pa = rawb + ra * RS + shift, *pb = rawb + rb * RS + shift;
        uint16_t *ha = reinterpret_cast<uint16_t *>(S.hp) + ra, *hb = reinterpret_cast<uint16_t *>(S.hp) + rb;
        uint32_t q[PD];
#pragma unroll
        for (int c = 0; c < PD; ++c) q[c] = (uint32_t)pa[c] | ((uint32_t)pb[c] << 16);
#pragma unroll
        for (int c = 0; c < BD; ++c) {
            const uint32_t h = 18u * (q[c] + q[c + 6]) + 34u * (q[c + 1] + q[c + 5]) + 48u * (q[c + 2] + q[c + 4]) + 56u * q[c + 3];
            ha[c * CS] = (uint16_t)(h & 0xFFFFu); // lanes write consecutive rows of one column: conflict-free
            hb[c * CS] = (uint16_t)(h >> 16);
        }
    }
    __syncwarp();

    // ---- steered BRIEF with the vertical blur pass at the sampled positions (:107-147) ---------------
    float a, b;
    steer_coeffs(angle, a, b);
    auto sample = [&](float px, float py) -> int {
        const int r = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
        const int c = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
        // taps = patch rows (r + 18) .. (r + 24) of blurred column c + 18: 7 contiguous u16, fetched as 4 words
        const int base = (c + BR) * CS + (r + BR);
        const uint32_t *w = S.hp + (base >> 1);
        const int sh = (base & 1) * 16;
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
        const uint32_t t01 = __funnelshift_r(w0, w1, sh), t23 = __funnelshift_r(w1, w2, sh);
        const uint32_t t45 = __funnelshift_r(w2, w3, sh), t6x = __funnelshift_r(w3, w3, sh);
        const int v = 18 * (int)((t01 & 0xFFFFu) + (t6x & 0xFFFFu)) + 34 * (int)((t01 >> 16) + (t45 >> 16)) +
                      48 * (int)((t23 & 0xFFFFu) + (t45 & 0xFFFFu)) + 56 * (int)(t23 >> 16);
        return (v + 32768) >> 16;
    };
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const uint32_t p = __ldg(&g_pat.w[t * 32 + lane]);
        const int t0 = sample((float)(int8_t)(p & 0xFF), (float)(int8_t)((p >> 8) & 0xFF));
        const int t1 = sample((float)(int8_t)((p >> 16) & 0xFF), (float)(int8_t)(p >> 24));
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
