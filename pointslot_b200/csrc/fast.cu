// fast.cu — K2: the per-cell FAST loop of ComputeKeyPointsOctTree (reference src/ORBextractor.cc:765-829).
//
// One CTA per 30-px cell.  The reference calls cv::FAST(cell, iniThFAST, nms=true) and, if that
// returns nothing, cv::FAST(cell, minThFAST, true) (:809-816).  Restated:
//   raw score R(p) = max over the 16 arcs of 9 ring pixels of min(v - p_k) resp. min(p_k - v), minus 1
//                    (OpenCV cornerScore<16>); p is a corner at threshold t  <=>  R >= t;
//   score map S_t = R where R >= t else 0, on the pixels >= 3 px inside the cell sub-image;
//   survivor <=> S_t(p) > S_t(q) for all 8 neighbours q (outside the tested range counts as 0);
//   the cell yields the iniTh survivors if there is at least one, else the minTh survivors;
//   order: raster inside the cell.
//
// Scoring is dense and branch-free.  With M1 = min_k max(p_k..p_k+8) and M2 = max_k min(p_k..p_k+8)
// over the raw ring pixels, R = max(v - M1, M2 - v) - 1.  Each thread scores 4 adjacent pixels as two
// 16x2-packed pairs: the ring bytes are gathered from 21 aligned shared-memory words with PRMT into
// the HIGH byte of each 16-bit half (the low byte is don't-care: it cannot change which high byte wins
// a min/max), and the 9-wide sliding min/max runs on the native 3-input VIMNMX3.U16x2 (80 per pair).
// The survivors are compacted in raster order with warp scans.
//
// Integer-only, hence bit-exact.  Bound: integer ALU (~60 instructions per pixel, VIMNMX3 dominated);
// algorithmic global bytes per launch = S * (sum_l P_l + 8 * n_candidates).
#include "pslot_internal.h"
#include "tma.cuh"

namespace {

constexpr int FT = 256;      // threads per CTA
constexpr int TILE_BYTES = 96 * 66; // largest TMA box: wCell, hCell <= 59 -> 96 x 65 bytes (+1 score row)
constexpr int MAX_ITERS = 4; // ceil(59 rows * 16 groups / 256)
// A TMA box must start on a 16-byte boundary of the global row, so the box starts at the aligned x below
// (cell start - 1); the residual 0..15 byte offset is uniform per CTA: a whole-word part folded into the
// shared-memory word index and a 0..3 byte part removed with one funnel shift per word.

// window bytes: A = offsets -4..-1, B = 0..3, C = 4..7 relative to the group's first pixel.
// Returns a 16x2 value whose high bytes are window[O0] (low half) and window[O0 + 1] (high half).
template <int O0>
__device__ __forceinline__ uint32_t pick(uint32_t A, uint32_t B, uint32_t C)
{
    static_assert(O0 >= -4 && O0 + 1 <= 7, "offset out of window");
    if constexpr (O0 + 1 <= 3) {
        constexpr int i0 = O0 + 4, i1 = O0 + 5;
        return __byte_perm(A, B, i0 | (i0 << 4) | (i1 << 8) | (i1 << 12));
    } else {
        static_assert(O0 + 1 <= 3 || O0 >= 0, "pair straddles A and C");
        constexpr int i0 = O0, i1 = O0 + 1;
        return __byte_perm(B, C, i0 | (i0 << 4) | (i1 << 8) | (i1 << 12));
    }
}

// ring position k -> (dx, dy) (OpenCV FAST 9-16 pixel order), rows indexed dy + 3
// k:  0      1      2      3      4      5       6       7       8       9        10       11       12      13      14      15
//   (0,3)  (1,3)  (2,2)  (3,1)  (3,0)  (3,-1)  (2,-2)  (1,-3)  (0,-3)  (-1,-3)  (-2,-2)  (-3,-1)  (-3,0)  (-3,1)  (-2,2)  (-1,3)
template <int J>
__device__ __forceinline__ void ring_pairs(const uint32_t (&w)[7][3], uint32_t (&P)[16])
{
    constexpr int o = 2 * J;
#define PK(k, dx, dy) P[k] = pick<o + (dx)>(w[(dy) + 3][0], w[(dy) + 3][1], w[(dy) + 3][2])
    PK(0, 0, 3);   PK(1, 1, 3);   PK(2, 2, 2);    PK(3, 3, 1);
    PK(4, 3, 0);   PK(5, 3, -1);  PK(6, 2, -2);   PK(7, 1, -3);
    PK(8, 0, -3);  PK(9, -1, -3); PK(10, -2, -2); PK(11, -3, -1);
    PK(12, -3, 0); PK(13, -3, 1); PK(14, -2, 2);  PK(15, -1, 3);
#undef PK
}

// M1 = min over the 16 arcs of max(9 contiguous), M2 = max over arcs of min(9 contiguous); 16x2 lanes.
__device__ __forceinline__ void arc_minmax(const uint32_t (&P)[16], uint32_t &M1, uint32_t &M2)
{
    uint32_t a[16], b[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        a[k] = __vimax3_u16x2(P[k], P[(k + 1) & 15], P[(k + 2) & 15]);
        b[k] = __vimin3_u16x2(P[k], P[(k + 1) & 15], P[(k + 2) & 15]);
    }
    uint32_t m1 = 0xFFFFFFFFu, m2 = 0u;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const uint32_t x0 = __vimax3_u16x2(a[k], a[(k + 3) & 15], a[(k + 6) & 15]);
        const uint32_t x1 = __vimax3_u16x2(a[k + 1], a[(k + 4) & 15], a[(k + 7) & 15]);
        m1 = __vimin3_u16x2(m1, x0, x1);
        const uint32_t y0 = __vimin3_u16x2(b[k], b[(k + 3) & 15], b[(k + 6) & 15]);
        const uint32_t y1 = __vimin3_u16x2(b[k + 1], b[(k + 4) & 15], b[(k + 7) & 15]);
        m2 = __vimax3_u16x2(m2, y0, y1);
    }
    M1 = m1;
    M2 = m2;
}

__device__ __forceinline__ int raw_score(int v, int m1, int m2) { return max(v - m1, m2 - v) - 1; }

__global__ void __launch_bounds__(FT)
k_fast_cells(const __grid_constant__ Geom g, const __grid_constant__ TmapSet maps,
             int *__restrict__ cell_count, uint32_t *__restrict__ cell_cand)
{
    __shared__ __align__(128) uint8_t tile[TILE_BYTES];
    __shared__ __align__(16) uint8_t score[TILE_BYTES];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int warp_tot[FT / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = blockIdx.y;
    int cell = blockIdx.x;
    int l = 0;
    while (l + 1 < g.nlevels && cell >= g.lv[l + 1].cell_base) ++l;
    const LevelGeom &L = g.lv[l];
    cell -= L.cell_base;
    const int ci = cell / L.n_cols, cj = cell - ci * L.n_cols;
    int *count_out = cell_count + (size_t)slot * g.cells_per_slot + L.cell_base + cell;

    // cell sub-image [ini, max) (:791-806); integers, so the reference's float compares are exact
    const int iniY = PSLOT_MINB + ci * L.h_cell, iniX = PSLOT_MINB + cj * L.w_cell;
    const int maxY = min(iniY + L.h_cell + 6, L.max_by), maxX = min(iniX + L.w_cell + 6, L.max_bx);
    const int tw = maxX - iniX, th = maxY - iniY;
    const int iw = tw - 6, ih = th - 6; // pixels cv::FAST actually tests
    if (iniY >= L.max_by - 3 || iniX >= L.max_bx - 3 || iw <= 0 || ih <= 0) {
        if (tid == 0) *count_out = 0;
        return;
    }

    // ---- stage the cell: one TMA box load (x, y, slot) -> dense shared tile, row stride = box width ----
    const int TS = L.box_w, TSW = TS >> 2;
    const int xa = (iniX - 1) & ~15;          // 16-byte aligned box origin
    const int sh0 = (iniX - 1) - xa;          // image column iniX-1+b lives at shared byte sh0 + b
    const int wsh = sh0 >> 2, bsh = (sh0 & 3) * 8;
    if (tid == 0) {
        tma::mbar_init(&bar, 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_arrive_expect_tx(&bar, (uint32_t)(L.box_w * L.box_h));
        tma::load_3d(tile, &maps.lv[l], &bar, xa, iniY, slot);
    }
    {
        uint32_t *s4 = reinterpret_cast<uint32_t *>(score);
        const int nwords = (th + 1) * TSW;
        for (int i = tid; i < nwords; i += FT) s4[i] = 0;
    }
    tma::mbar_wait(&bar, 0);
    __syncthreads();

    // tasks: (interior row y, group of 4 pixels gx); G groups per row, G a power of two
    const int gshift = iw <= 32 ? 3 : 4;
    const int G = 1 << gshift;
    const int ntasks = ih << gshift;
    const int tmin = g.min_th, tini = g.ini_th;

    // ---- pass 1: dense exact scores ---------------------------------------------------------------
    for (int t = tid; t < ntasks; t += FT) {
        const int y = t >> gshift, gx = t & (G - 1);
        if (4 * gx >= iw) continue;
        // window of the group: bytes -4..7 around its first pixel = shared bytes sh0 + 4*gx + (0..11) of each ring row
        const uint32_t *base = reinterpret_cast<const uint32_t *>(tile) + (y + 3) * TSW + gx + wsh;
        uint32_t w[7][3];
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const uint32_t *p = base + (r - 3) * TSW;
            const uint32_t w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3];
            w[r][0] = __funnelshift_r(w0, w1, bsh);
            w[r][1] = __funnelshift_r(w1, w2, bsh);
            w[r][2] = __funnelshift_r(w2, w3, bsh);
        }
        uint32_t P[16], m1a, m2a, m1b, m2b;
        ring_pairs<0>(w, P);
        arc_minmax(P, m1a, m2a);
        ring_pairs<1>(w, P);
        arc_minmax(P, m1b, m2b);
        const uint32_t c = w[3][1]; // the 4 centre pixels
        int s0 = raw_score((int)(c & 0xFF), (int)((m1a >> 8) & 0xFF), (int)((m2a >> 8) & 0xFF));
        int s1 = raw_score((int)((c >> 8) & 0xFF), (int)(m1a >> 24), (int)(m2a >> 24));
        int s2 = raw_score((int)((c >> 16) & 0xFF), (int)((m1b >> 8) & 0xFF), (int)((m2b >> 8) & 0xFF));
        int s3 = raw_score((int)(c >> 24), (int)(m1b >> 24), (int)(m2b >> 24));
        const int x0 = 4 * gx;
        s0 = (s0 >= tmin) ? s0 : 0;
        s1 = (s1 >= tmin && x0 + 1 < iw) ? s1 : 0;
        s2 = (s2 >= tmin && x0 + 2 < iw) ? s2 : 0;
        s3 = (s3 >= tmin && x0 + 3 < iw) ? s3 : 0;
        reinterpret_cast<uint32_t *>(score)[(y + 3) * TSW + gx + 1] =
            (uint32_t)s0 | ((uint32_t)s1 << 8) | ((uint32_t)s2 << 16) | ((uint32_t)s3 << 24);
    }
    __syncthreads();

    // ---- pass 2: 3x3 strict-maximum test at both thresholds ------------------------------------------
    // flag bits per task: [3:0] survivor at iniTh, [7:4] survivor at minTh
    uint32_t flags[MAX_ITERS];
    bool any_ini = false;
#pragma unroll
    for (int it = 0; it < MAX_ITERS; ++it) {
        flags[it] = 0;
        const int t = tid + it * FT;
        if (t < ntasks) {
            const int y = t >> gshift, gx = t & (G - 1);
            const int off = (y + 3) * TS + 4 * gx + 4;
            const uint32_t s4 = *reinterpret_cast<const uint32_t *>(score + off);
            if (s4) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int s = (int)((s4 >> (8 * i)) & 0xFF);
                    if (s) {
                        const uint8_t *p = score + off + i;
                        int nmax = max(max((int)p[-TS - 1], (int)p[-TS]), max((int)p[-TS + 1], (int)p[-1]));
                        nmax = max(max(nmax, (int)p[1]), max(max((int)p[TS - 1], (int)p[TS]), (int)p[TS + 1]));
                        // at threshold t the neighbour map is (score >= t ? score : 0)
                        const bool keep_min = s > nmax;
                        const bool keep_ini = (s >= tini) && (s > (nmax >= tini ? nmax : 0));
                        flags[it] |= (keep_ini ? 1u : 0u) << i;
                        flags[it] |= (keep_min ? 16u : 0u) << i;
                        any_ini |= keep_ini;
                    }
                }
            }
        }
    }
    const int use_ini = __syncthreads_or(any_ini);
    const int fshift = use_ini ? 0 : 4;

    // ---- pass 3: raster-ordered compaction ---------------------------------------------------------
    uint32_t *out = cell_cand + (size_t)slot * g.cand_per_slot + L.cand_base + (size_t)cell * L.cand_stride;
    const int ox = iniX + 3 - PSLOT_MINB, oy = iniY + 3 - PSLOT_MINB; // window-relative (:820-825)
    int total = 0;
#pragma unroll
    for (int it = 0; it < MAX_ITERS; ++it) {
        if (it * FT >= ntasks) break; // block-uniform
        const uint32_t m = (flags[it] >> fshift) & 0xFu;
        const int c = __popc(m);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int before = 0, chunk = 0;
#pragma unroll
        for (int w = 0; w < FT / 32; ++w) {
            const int v = warp_tot[w];
            before += (w < warp) ? v : 0;
            chunk += v;
        }
        if (m) {
            const int t = tid + it * FT;
            const int y = t >> gshift, gx = t & (G - 1);
            const int off = (y + 3) * TS + 4 * gx + 4;
            int pos = total + before + incl - c;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (m & (1u << i)) {
                    if (pos < L.cand_stride) out[pos] = pslot_pack(4 * gx + i + ox, y + oy, score[off + i]);
                    ++pos;
                }
        }
        total += chunk;
        __syncthreads();
    }
    if (tid == 0) *count_out = min(total, L.cand_stride);
}

} // namespace

int launch_fast(const Geom &g, const TmapSet &maps, int *cell_count, uint32_t *cell_cand, int nslots, cudaStream_t st)
{
    dim3 grid(g.cells_per_slot, nslots);
    k_fast_cells<<<grid, FT, 0, st>>>(g, maps, cell_count, cell_cand);
    return 1;
}
