// pyramid.cu — K1: ComputePyramid (reference src/ORBextractor.cc:1107-1132).
//
// Level l is cv::resize(level l-1, INTER_LINEAR) for CV_8UC1, i.e. OpenCV's 11-bit fixed-point
// bilinear kernel; the taps (source indices + coefficients) are computed on the host exactly as
// OpenCV does and the device evaluates the integer formula
//     dst = ( ((c0y*(H0>>4))>>16) + ((c1y*(H1>>4))>>16) + 2 ) >> 2,   H = s0*c0x + s1*c1x.
// Integer-only, hence bit-exact.  Levels are chained (each from the previous one, :1120), so a batch
// needs nlevels-1 dependent launches; each launch covers every slot of the batch.
//
// Each CTA produces a 128 x 64 output tile.  The source rectangle it needs (about 156 x 79 bytes at scale 1.2) arrives
// by ONE TMA box load (cp.async.bulk.tensor.3d: no load / store instructions, no staging registers, out-of-image bytes
// zero-filled), the 64 row taps are unpacked once into shared memory while the box is in flight, and every thread then
// produces 2 adjacent columns x 16 consecutive rows: one broadcast LDS.128 for the row tap, two word loads per
// source row, a PRMT + IDP.2A per row sum, two multiply-highs per pixel.  A warp covers 64 consecutive output columns of
// one row, so every gather is one conflict-free wavefront whatever the box width.  The finished tile goes back with one
// TMA box store, which clips at the level's right / bottom edge.  (History: byte gathers, 4 LDS.U8 per pixel, pinned the
// shared-memory pipe at 90 %; staging the source with 16-byte LDG + STS and writing the tile with 16-byte stores cost a
// fifth of the instructions and two more CTA barriers - ncu: barrier 15 % + long_scoreboard 14 % of the stalls.)
//
// Bound: instruction issue above the HBM/L2 stream; algorithmic bytes per launch = S * (P_{l-1} + P_l).
#include "pslot_internal.h"
#include "tma.cuh"

namespace {

constexpr int TW = PYR_TILE_W; // output tile width
constexpr int TH = PYR_TILE_H; // output tile height
constexpr int NT = 256;        // threads per CTA: 64 column pairs x 4 row groups
constexpr int SRC_W = 208;     // largest source box row in bytes (13 x 16): 128 * 1.45 + 2 + 15 alignment
constexpr int SRC_H = (PYR_TILE_H * 29 + 19) / 20 + 6; // largest source box height: tile height * 1.45 + 2 and slack

__global__ void __launch_bounds__(NT, PYR_TILE_H == 32 ? 8 : PYR_TILE_H == 64 ? 6 : PYR_TILE_H == 96 ? 4 : 3)
k_resize_level(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap dst_map, int box_w, int box_h,
               int dw, int dh, const ResizeTap *__restrict__ xt, const ResizeTap *__restrict__ yt)
{
    __shared__ __align__(128) uint8_t s_src[SRC_H * SRC_W];
    __shared__ __align__(128) uint8_t s_out[TH * TW];
    __shared__ int4 s_yt[TH]; // row tap: byte offsets of the two source rows in s_src, c0, c1
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int dx0 = blockIdx.x * TW, dy0 = blockIdx.y * TH;
    const int dx1 = min(dx0 + TW, dw) - 1, dy1 = min(dy0 + TH, dh) - 1; // last output column / row of the tile
    const int slot = blockIdx.z;

    // source rectangle: columns from xa (rounded down to 16 bytes), rows from ya; pyramid_src_box sized the box so that it
    // covers the taps of every tile of this level
    const int xa = (int)xt[dx0].s0 & ~15;
    const int ya = (int)yt[dy0].s0;
    if (tid == 0) {
        tma::mbar_init(&bar, 1);
        tma::fence_barrier_init();
        tma::mbar_arrive_expect_tx(&bar, (uint32_t)(box_w * box_h));
        tma::load_3d(s_src, &src_map, &bar, xa, ya, slot);
    }
    if (tid < TH) {
        const ResizeTap t = yt[min(dy0 + tid, dy1)];
        // row coefficients pre-shifted by 16 so that (c * (h >> 4)) >> 16 is one multiply-high (c <= 2048, h >> 4 < 2^15)
        s_yt[tid] = make_int4(((int)t.s0 - ya) * box_w, ((int)t.s1 - ya) * box_w, (int)t.c0 << 16, (int)t.c1 << 16);
    }
    // Each thread produces a PAIR of adjacent columns for TH / 4 rows.  The source bytes of a pair span at most 7
    // bytes (ratio <= 1.45), i.e. two aligned words of the staged row: two word loads per source row replace four
    // byte gathers, one PRMT per pixel picks its two tap bytes (selectors are row-invariant), and one IDP.2A
    // (16-bit coefficients x two bytes) yields the row sum H.  The two column taps are adjacent bytes
    // (s1 == s0 + 1) except at the clamped right border (s1 == s0), where the pair (s0, s0 + 1) with coefficients
    // (c0 + c1, 0) gives the same sum.
    const int cp = tid & 63, rg = tid >> 6; // column pair, row group
    const ResizeTap ta = xt[min(dx0 + 2 * cp, dx1)], tb = xt[min(dx0 + 2 * cp + 1, dx1)];
    const int oa = (int)ta.s0 - xa, ob = (int)tb.s0 - xa;
    const int base = oa & ~3;
    const uint32_t ja = (uint32_t)(oa - base), jb = (uint32_t)(ob - base); // 0..3 and ja..ja+2
    const uint32_t sela = ja | ((ja + 1) << 4), selb = jb | ((jb + 1) << 4);
    const uint32_t ca = ta.s1 != ta.s0 ? ((uint32_t)ta.c0 | ((uint32_t)ta.c1 << 16)) : (uint32_t)(ta.c0 + ta.c1);
    const uint32_t cb = tb.s1 != tb.s0 ? ((uint32_t)tb.c0 | ((uint32_t)tb.c1 << 16)) : (uint32_t)(tb.c0 + tb.c1);
    const uint8_t *p0 = s_src + base;
    __syncthreads();          // s_yt and the barrier initialisation are visible
    tma::mbar_wait(&bar, 0);  // the source box has landed
    // A thread's 8 output rows are consecutive, and at scale 1.2 five out of six of them start on the source row the
    // previous one ended on: that row's sums are kept in registers (the test is uniform across the warp, which works on
    // one output row at a time), 1.2 instead of 2 source rows per output row.
    uint32_t pa = 0, pb = 0; // (H >> 4) of the previous output row's second source row, columns a and b
    int prev_off = -1;
#pragma unroll
    for (int k = 0; k < TH / 4; ++k) {
        const int r = rg * (TH / 4) + k;
        const int4 t = s_yt[r];
        uint32_t a0, b0;
        if (t.x == prev_off) {
            a0 = pa; b0 = pb;
        } else {
            const uint32_t *w0 = reinterpret_cast<const uint32_t *>(p0 + t.x);
            const uint32_t u0 = w0[0], u1 = w0[1];
            // H = c0 * src[s0] + c1 * src[s1]
            a0 = __dp2a_lo(ca, __byte_perm(u0, u1, sela), 0u) >> 4;
            b0 = __dp2a_lo(cb, __byte_perm(u0, u1, selb), 0u) >> 4;
        }
        const uint32_t *w1 = reinterpret_cast<const uint32_t *>(p0 + t.y);
        const uint32_t l0 = w1[0], l1 = w1[1];
        const uint32_t a1 = __dp2a_lo(ca, __byte_perm(l0, l1, sela), 0u) >> 4;
        const uint32_t b1 = __dp2a_lo(cb, __byte_perm(l0, l1, selb), 0u) >> 4;
        // in [0, 255] by construction (coefficients sum to 2048 per axis): no saturation needed
        const uint32_t va = (__umulhi((uint32_t)t.z, a0) + __umulhi((uint32_t)t.w, a1) + 2u) >> 2;
        const uint32_t vb = (__umulhi((uint32_t)t.z, b0) + __umulhi((uint32_t)t.w, b1) + 2u) >> 2;
        *reinterpret_cast<uint16_t *>(s_out + r * TW + 2 * cp) = (uint16_t)(va | (vb << 8));
        pa = a1; pb = b1; prev_off = t.y;
    }
    tma::fence_proxy_async(); // this thread's tile bytes are visible to the TMA unit ...
    __syncthreads();          // ... and so are everybody else's
    if (tid == 0) {
        tma::store_3d(&dst_map, s_out, dx0, dy0, slot); // clipped at the level's right / bottom edge
        tma::store_commit();
        tma::store_wait_read();
    }
}

} // namespace

void pyramid_src_box(const ResizeTap *xt, int dw, const ResizeTap *yt, int dh, int *box_w, int *box_h)
{
    int bw = 16, bh = 1;
    for (int dx0 = 0; dx0 < dw; dx0 += TW) {
        const int dx1 = (dx0 + TW < dw ? dx0 + TW : dw) - 1;
        const int xa = (int)xt[dx0].s0 & ~15, xb = (int)xt[dx1].s1;
        // the compute phase reads two aligned words from the first tap of a column pair: up to byte (s0 & ~3) + 7
        const int need = (((int)xt[dx1].s0 - xa) & ~3) + 8;
        const int w = ((xb - xa + 1 > need ? xb - xa + 1 : need) + 15) & ~15;
        bw = w > bw ? w : bw;
    }
    for (int dy0 = 0; dy0 < dh; dy0 += TH) {
        const int dy1 = (dy0 + TH < dh ? dy0 + TH : dh) - 1;
        const int h = (int)yt[dy1].s1 - (int)yt[dy0].s0 + 1;
        bh = h > bh ? h : bh;
    }
    *box_w = bw;
    *box_h = bh;
}

bool pyramid_tiles_fit(const ResizeTap *xt, int dw, const ResizeTap *yt, int dh)
{
    int bw, bh;
    pyramid_src_box(xt, dw, yt, dh, &bw, &bh);
    return bw <= SRC_W && bh <= SRC_H;
}

int launch_pyramid(const Geom &g, const PyrPtrs &pyr, const PyrMaps &maps, const ResizeTap *taps, int nslots, int *err_flag,
                   cudaStream_t st)
{
    (void)pyr; (void)err_flag; // the tile fit is guaranteed on the host (pyramid_tiles_fit): no device-side guard left
    int launches = 0;
    for (int l = 1; l < g.nlevels; ++l) {
        const LevelGeom &D = g.lv[l];
        dim3 grid((D.w + TW - 1) / TW, (D.h + TH - 1) / TH, nslots);
        k_resize_level<<<grid, NT, 0, st>>>(maps.src[l - 1], maps.dst[l], maps.box_w[l - 1], maps.box_h[l - 1], D.w, D.h,
                                            taps + D.tab_x, taps + D.tab_y);
        ++launches;
    }
    return launches;
}
