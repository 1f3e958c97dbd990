// fast.cu — K2: the per-cell FAST loop of ComputeKeyPointsOctTree (reference src/ORBextractor.cc:765-829).
//
// One WARP per 30-px cell, one cell per CTA (no block-level barrier anywhere).  The reference calls
// cv::FAST(cell, iniThFAST, nms=true) and, if that returns nothing, cv::FAST(cell, minThFAST, true) (:809-816).
// Restated:
//   raw score R(p) = max over the 16 arcs of 9 ring pixels of min(v - p_k) resp. min(p_k - v), minus 1
//                    (OpenCV cornerScore<16>); p is a corner at threshold t  <=>  R >= t;
//   score map S_t = R where R >= t else 0, on the pixels >= 3 px inside the cell sub-image;
//   survivor <=> S_t(p) > S_t(q) for all 8 neighbours q (outside the tested range counts as 0);
//   the cell yields the iniTh survivors if there is at least one, else the minTh survivors;
//   order: raster inside the cell.
//
// Stages per cell:
//   0. one TMA box load (cp.async.bulk.tensor.3d, coordinates x, y, slot) brings the cell tile into shared memory;
//   A. threshold iniThFAST.  A byte-SWAR rejection test (compass ring pixels 0/8 and 4/12 against the centre, 4
//      pixels per VABSDIFF4) lists the 4-pixel groups that can hold a corner - a few percent of them on natural
//      images - and only those get the exact score: with M1 = min_k max(p_k..p_k+8) and M2 = max_k min(p_k..p_k+8)
//      over the raw ring pixels, R = max(v - M1, M2 - v) - 1.  A lane scores its 4 pixels as two 16x2-packed pairs
//      (ring bytes extracted on the FMA pipe, IDP.4A + IMAD) on the native 3-input VIMNMX3.U16x2 (80 per pair;
//      measured on B200: 2- and 3-input integer min/max both issue every 2 cycles per SM sub-partition, so the
//      3-input form - two comparisons per slot - is the cheapest network).  The strict 3x3 maximum test runs packed
//      (16x2) on the same list and the survivors are compacted in raster order by a warp scan, straight to global;
//   B. only if A found nothing (:812-816): threshold minThFAST, at which the rejection test passes most pixels, so
//      every group is scored (dense, branch-free), the groups holding some R >= minTh are listed by ballot, and the
//      same maxima test + compaction follows.
//
// Integer-only, hence bit-exact.  Bound: the integer ALU pipe (ncu: math_pipe_throttle + not_selected stalls; DRAM
// traffic equals the algorithmic bytes); algorithmic global bytes per launch = S * (sum_l P_l + 8 * n_candidates)
// (SURVEY.md 8d).
#include <algorithm>

#include "pslot_internal.h"
#include "tma.cuh"

namespace {

// One warp per CTA: a cell that needs the minThFAST pass takes ~3.6 x as long as one that does not, and a CTA keeps its
// shared memory and warp slots until its slowest warp is done - with 4 cells per CTA two thirds of the CTAs held a slow
// cell and ncu showed 30 % achieved occupancy against 44 % theoretical (827 -> 757 us for 250 images with 1 warp).
#ifndef FAST_WARPS
#define FAST_WARPS 1
#endif
constexpr int FW = FAST_WARPS; // warps (= cells) per CTA
#ifndef FAST_MIN_BLOCKS
#define FAST_MIN_BLOCKS (28 / FAST_WARPS)
#endif
// A TMA box must start on a 16-byte boundary of the global row, so the box starts at the aligned x below
// (cell start - 1); the residual 0..15 byte offset is uniform per warp: a whole-word part folded into the
// shared-memory word index and a 0..3 byte part S that is removed from the ring rows by funnel shifts.  (Four
// compile-time specialisations on S save those shifts, but with warps of one SM spread over four copies of the
// scoring loop the instruction cache thrashes: ncu showed 46 % no_instructions stalls.  One copy it is.)

// n[0..2]: twelve bytes of one ring row, already shifted so that the group's first pixel is byte 4.
// Returns a clean 16x2 value holding the pixels at offsets O0 and O0 + 1 from the group's first pixel.
template <int O0>
__device__ __forceinline__ uint32_t pick(const uint32_t (&n)[3])
{
    constexpr int i0 = 4 + O0, i1 = i0 + 1; // byte indices into the 12-byte row segment
    static_assert(i0 >= 0 && i1 <= 11, "offset out of the staged segment");
    // Byte extraction on the FMA pipe: IDP.4A with a one-hot byte mask returns that byte zero-extended, and one
    // IMAD packs the pair as p0 | p1 << 16.  The integer ALU pipe - the kernel's bottleneck - is left to the
    // min/max network; PRMT would cost one ALU slot per operand.
    const uint32_t p0 = __dp4a(n[i0 >> 2], 1u << (8 * (i0 & 3)), 0u);
    const uint32_t p1 = __dp4a(n[i1 >> 2], 1u << (8 * (i1 & 3)), 0u);
    return p1 * 65536u + p0;
}

// ring position k -> (dx, dy) (OpenCV FAST 9-16 pixel order), rows indexed dy + 3
// k:  0      1      2      3      4      5       6       7       8       9        10       11       12      13      14      15
//   (0,3)  (1,3)  (2,2)  (3,1)  (3,0)  (3,-1)  (2,-2)  (1,-3)  (0,-3)  (-1,-3)  (-2,-2)  (-3,-1)  (-3,0)  (-3,1)  (-2,2)  (-1,3)
template <int J>
__device__ __forceinline__ void ring_pairs(const uint32_t (&n)[7][3], uint32_t (&P)[16])
{
    constexpr int o = 2 * J;
#define PK(k, dx, dy) P[k] = pick<o + (dx)>(n[(dy) + 3])
    PK(0, 0, 3);   PK(1, 1, 3);   PK(2, 2, 2);    PK(3, 3, 1);
    PK(4, 3, 0);   PK(5, 3, -1);  PK(6, 2, -2);   PK(7, 1, -3);
    PK(8, 0, -3);  PK(9, -1, -3); PK(10, -2, -2); PK(11, -3, -1);
    PK(12, -3, 0); PK(13, -3, 1); PK(14, -2, 2);  PK(15, -1, 3);
#undef PK
}

// M1 = min over the 16 arcs of max(9 contiguous ring pixels), M2 = max over arcs of min(...); 16x2 lanes.
// Every integer ALU instruction issues at one per two cycles per SM sub-partition on B200 (measured with ncu: the
// kernel is pinned on that pipe), so the 3-input VIMNMX3 - two comparisons per issue slot - is the cheapest
// form: 16 triples + 16 arcs (max3 of three triples) + 8 to reduce = 40 per side.  (A van Herk prefix/suffix
// network needs 59 two-input operations and measured no faster.)
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

// Rejection test at threshold t for every 4-pixel group of the cell (phase A).  A 9-arc of the 16-pixel ring
// contains at least one pixel of every opposite pair {k, k + 8}; so a pixel whose ring pixels 0 / 8 (or 4 / 12:
// the compass points) BOTH lie within t of the centre cannot be a corner at threshold t.  Byte-SWAR on 4 pixels at
// a time: VABSDIFF4 against the centre quad, then bit 7 of ((d + 127 - t) | d) is set where d > t (a carry from
// the byte below can only add a false positive, which merely costs an exact score).  The groups with a possible
// corner are listed in raster order; on KITTI-like images at iniThFAST = 20 that is ~6 % of the groups.
__device__ __forceinline__ int quick_pass(const uint32_t *tile32, int TSW, int wsh, int S, int lane, int ntasks,
                                          int gshift, int iw, uint32_t bias, uint16_t *glist)
{
    // A lane tests TWO adjacent groups (8 pixels) of one row: the three rows it needs are fetched as 5 + 3 + 3 words
    // and re-aligned by the warp-uniform byte offset S with funnel shifts, after which the centre quads are whole
    // words and the left / right ring quads one more funnel shift each - 11 loads per pair where two separate groups
    // took 20.
    const int pshift = gshift - 1, P1 = (1 << pshift) - 1; // pairs per row
    const int npairs = ntasks >> 1;
    const int gmax = (iw - 1) >> 2; // last group of a row that holds tested pixels
    const int sh = 8 * S;
    const unsigned lt = (1u << lane) - 1u;
    int nlist = 0;
    constexpr int QP = 2; // pairs in flight per lane (4: slower, the clamped tail costs more than the overlap gains)
    for (int base = 0; base < npairs; base += 32 * QP) {
        bool hit[QP][2];
        int tq[QP];
        // independent pairs per lane and iteration, so that the loads of one hide the dependent chain of the other
#pragma unroll
        for (int w = 0; w < QP; ++w) {
            const int p = base + 32 * w + lane;
            const int pc = min(p, npairs - 1);
            const int y = pc >> pshift, gx = (pc & P1) << 1;
            // rows y - 3, y, y + 3 around the pair's first pixel (clamped task, masked result: no divergence)
            const uint32_t *mid = tile32 + (y + 3) * TSW + gx + wsh;
            const uint32_t *top = mid - 3 * TSW, *bot = mid + 3 * TSW;
            const uint32_t m0 = mid[0], m1 = mid[1], m2 = mid[2], m3 = mid[3], m4 = mid[4];
            const uint32_t t1 = top[1], t2 = top[2], t3 = top[3], b1 = bot[1], b2 = bot[2], b3 = bot[3];
            // normalised words: n1 = pixels 0..3, n2 = pixels 4..7 of the pair, n0 / n3 the four pixels before / after
            const uint32_t n0 = __funnelshift_r(m0, m1, sh), n1 = __funnelshift_r(m1, m2, sh),
                           n2 = __funnelshift_r(m2, m3, sh), n3 = __funnelshift_r(m3, m4, sh);
            const uint32_t T0 = __funnelshift_r(t1, t2, sh), T1 = __funnelshift_r(t2, t3, sh);
            const uint32_t B0 = __funnelshift_r(b1, b2, sh), B1 = __funnelshift_r(b2, b3, sh);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const uint32_t v = u ? n2 : n1;
                const uint32_t west = u ? __funnelshift_r(n1, n2, 8) : __funnelshift_r(n0, n1, 8);   // pixels -3..0 of the group
                const uint32_t east = u ? __funnelshift_r(n2, n3, 24) : __funnelshift_r(n1, n2, 24); // pixels 3..6
                const uint32_t d4 = __vabsdiffu4(east, v), d12 = __vabsdiffu4(west, v);
                const uint32_t d0 = __vabsdiffu4(u ? B1 : B0, v), d8 = __vabsdiffu4(u ? T1 : T0, v);
                const uint32_t e0 = (d0 + bias) | d0, e8 = (d8 + bias) | d8, e4 = (d4 + bias) | d4, e12 = (d12 + bias) | d12;
                uint32_t pass = (e0 | e8) & (e4 | e12) & 0x80808080u;
                const int rem = iw - 4 * (gx + u); // tested pixels left in this row from the group's first one
                if (rem < 4) pass &= rem > 0 ? (1u << (8 * rem)) - 1u : 0u;
                hit[w][u] = pass != 0 && p < npairs && gx + u <= gmax;
            }
            tq[w] = (y << gshift) | gx;
        }
        // raster order: lane L's two groups are consecutive, before lane L + 1's; each block of 32 pairs before the next
#pragma unroll
        for (int w = 0; w < QP; ++w) {
            const unsigned q0 = __ballot_sync(0xffffffffu, hit[w][0]), q1 = __ballot_sync(0xffffffffu, hit[w][1]);
            const int before = nlist + __popc(q0 & lt) + __popc(q1 & lt);
            if (hit[w][0]) glist[before] = (uint16_t)tq[w];
            if (hit[w][1]) glist[before + (hit[w][0] ? 1 : 0)] = (uint16_t)(tq[w] + 1);
            nlist += __popc(q0) + __popc(q1);
        }
    }
    return nlist;
}

// Exact scores of the `count` listed 4-pixel groups (phase A), or of every group of the cell when list == nullptr
// (phase B, count = ntasks).
__device__ __forceinline__ void score_pass(const uint32_t *tile32, uint32_t *score32, int TSW, int wsh, int S, int lane,
                                           const uint16_t *list, int count, int gshift, int iw)
{
    const int G1 = (1 << gshift) - 1;
    const int sh = 8 * S;
    for (int j = lane; j < count; j += 32) {
        const int t = list ? (int)list[j] : j;
        const int y = t >> gshift, gx = t & G1;
        if (4 * gx >= iw) continue;
        // 12-byte segment of each ring row holding bytes -4..7 around the group's first pixel
        const uint32_t *base = tile32 + (y + 3) * TSW + gx + wsh;
        uint32_t n[7][3];
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const uint32_t *p = base + (r - 3) * TSW;
            const uint32_t w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3];
            n[r][0] = __funnelshift_r(w0, w1, sh); n[r][1] = __funnelshift_r(w1, w2, sh); n[r][2] = __funnelshift_r(w2, w3, sh);
        }
        uint32_t P[16];
        uint32_t m1a, m2a, m1b, m2b;
        ring_pairs<0>(n, P);
        arc_minmax(P, m1a, m2a);
        ring_pairs<1>(n, P);
        arc_minmax(P, m1b, m2b);
        // the 4 centre pixels as clean (zero-extended) 16x2 pairs (the extrema already are)
        const uint32_t v01 = pick<0>(n[3]), v23 = pick<2>(n[3]);
        const uint32_t a01 = m1a, b01 = m2a, a23 = m1b, b23 = m2b;
        // R = max(v - M1, M2 - v) - 1 clamped at 0; the threshold is applied by the maxima test
        const uint32_t one = 0x00010001u;
        const uint32_t r01 = __vimax_s16x2_relu(__vsub2(__vsub2(v01, a01), one), __vsub2(__vsub2(b01, v01), one));
        const uint32_t r23 = __vimax_s16x2_relu(__vsub2(__vsub2(v23, a23), one), __vsub2(__vsub2(b23, v23), one));
        uint32_t s4 = __byte_perm(r01, r23, 0x6420);
        const int rem = iw - 4 * gx; // tested pixels left in this row (>= 1)
        if (rem < 4) s4 &= (1u << (8 * rem)) - 1u;
        // score map: tested pixel x lives at byte x + 4 of its row
        score32[(y + 3) * TSW + gx + 1] = s4;
    }
}

// Phase A with a short list (the usual case: ~6 % of a cell's 256 groups): one lane per PIXEL PAIR instead of per group, so
// that a list of up to 16 groups fills the warp in one pass of half the length (one extraction + one min/max network per
// lane).  The second pair of a group is the first pair of the group shifted by two bytes: the shift joins the alignment
// funnel shift (per-lane amount), the code stays uniform.
__device__ __forceinline__ void score_pass_pairs(const uint32_t *tile32, uint8_t *score8, int TS, int wsh, int S, int lane,
                                                 const uint16_t *list, int count, int gshift, int iw)
{
    const int TSW = TS >> 2, G1 = (1 << gshift) - 1;
    for (int j = lane; j < 2 * count; j += 32) {
        const int t = (int)list[j >> 1], half = j & 1;
        const int y = t >> gshift, gx = t & G1;
        const int px = 4 * gx + 2 * half; // first pixel of the pair inside the row of tested pixels
        if (px >= iw) continue;
        const int boff = S + 2 * half;    // byte offset of the pair's "pixel 0" from the aligned word
        const int sh = 8 * (boff & 3);
        const uint32_t *base = tile32 + (y + 3) * TSW + gx + wsh + (boff >> 2);
        uint32_t n[7][3];
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const uint32_t *p = base + (r - 3) * TSW;
            const uint32_t w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3];
            n[r][0] = __funnelshift_r(w0, w1, sh); n[r][1] = __funnelshift_r(w1, w2, sh); n[r][2] = __funnelshift_r(w2, w3, sh);
        }
        uint32_t P[16], m1, m2;
        ring_pairs<0>(n, P);
        arc_minmax(P, m1, m2);
        const uint32_t v01 = pick<0>(n[3]);
        const uint32_t one = 0x00010001u;
        const uint32_t r01 = __vimax_s16x2_relu(__vsub2(__vsub2(v01, m1), one), __vsub2(__vsub2(m2, v01), one));
        uint32_t s2 = __byte_perm(r01, 0u, 0x4420) & 0xFFFFu; // the two scores as adjacent bytes
        if (px + 1 >= iw) s2 &= 0xFFu;
        // score map: tested pixel x lives at byte x + 4 of its row
        *reinterpret_cast<uint16_t *>(score8 + (y + 3) * TS + 4 + px) = (uint16_t)s2;
    }
}

__global__ void __launch_bounds__(FW * 32, FAST_MIN_BLOCKS)
k_fast_cells(const __grid_constant__ Geom g, const __grid_constant__ TmapSet maps, const uint4 *__restrict__ cells,
             int *__restrict__ cell_count, uint32_t *__restrict__ cell_cand, int warp_smem_bytes, int *__restrict__ work_counter,
             int total_work)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[FW];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per-warp shared memory: [ TMA tile (box_w x box_h) | score map (same box) | group list ]
    uint8_t *tile = smem + (size_t)warp * warp_smem_bytes;
    uint64_t *bar = &bars[warp];
    if (lane == 0) {
        tma::mbar_init(bar, 1);
        tma::fence_barrier_init();
    }
    __syncwarp();
    // Persistent warps: every warp pulls (slot, cell) work items from a global counter until none are left - a cell that
    // needs the minThFAST pass takes ~3.6 x as long as one that does not, so a static assignment leaves warp slots idle
    // behind the slow cells, and a CTA per cell pays the block scheduler's turnaround 120 000 times per launch.  The next
    // item's id and record are fetched while the current cell is processed.
    uint32_t parity = 0;
    int id = 0;
    if (lane == 0) id = atomicAdd(work_counter, 1);
    id = __shfl_sync(0xffffffffu, id, 0);
    uint4 rec = make_uint4(0, 0, 0, 0);
    int slot = 0, cell = 0;
    if (id < total_work) {
        slot = id / g.cells_per_slot; cell = id - slot * g.cells_per_slot;
        rec = __ldg(cells + cell);
    }
    while (id < total_work) {
    // the cell's record (make_fast_cells): sub-image [ini, ini + (tw, th)) (:791-806), level, TMA box, candidate slots
    const int iniX = (int)(rec.x & 0xFFFFu), iniY = (int)(rec.x >> 16);
    const int tw = (int)(rec.y & 0xFFu), th = (int)((rec.y >> 8) & 0xFFu);
    const int l = (int)((rec.y >> 16) & 0xFFu), box_h = (int)(rec.y >> 24);
    const int cand_stride = (int)(rec.w & 0xFFFFu), TS = (int)(rec.w >> 16), TSW = TS >> 2;
    const uint32_t cand_off = rec.z;
    const int iw = tw - 6, ih = th - 6; // pixels cv::FAST actually tests
    const int my_slot = slot, my_cell = cell;
    const int box_bytes = TS * box_h;
    uint8_t *score = tile + ((box_bytes + 127) & ~127);
    uint16_t *glist = reinterpret_cast<uint16_t *>(score + ((box_bytes + 127) & ~127)); // <= 16 groups x ih rows

    // ---- stage the cell: one TMA box load (x, y, slot) -> dense shared tile, row stride = box width; a second box
    // far outside the image clears the score map (the TMA unit zero-fills out-of-bounds elements: no store instructions)
    const int xa = (iniX - 1) & ~15;          // 16-byte aligned box origin
    const int sh0 = (iniX - 1) - xa;          // image column iniX-1+b lives at shared byte sh0 + b
    const int wsh = sh0 >> 2;
    if (lane == 0 && tw != 0) {
        tma::fence_proxy_async(); // the previous cell's reads of this shared memory are ordered before the TMA writes
        tma::mbar_arrive_expect_tx(bar, 2u * (uint32_t)box_bytes);
        tma::load_3d(tile, &maps.lv[l], bar, xa, iniY, my_slot);
        tma::load_3d(score, &maps.lv[l], bar, xa, 1 << 14, my_slot);
    }
    // next work item, in flight while this cell is processed
    int next_id = 0;
    if (lane == 0) next_id = atomicAdd(work_counter, 1);
    next_id = __shfl_sync(0xffffffffu, next_id, 0);
    id = next_id;
    if (id < total_work) {
        slot = id / g.cells_per_slot; cell = id - slot * g.cells_per_slot;
        rec = __ldg(cells + cell);
    }
    if (tw == 0) {
        if (lane == 0) cell_count[(size_t)my_slot * g.cells_per_slot + my_cell] = 0;
        continue;
    }
    tma::mbar_wait(bar, parity);
    parity ^= 1u;

    // tasks: (interior row y, group of 4 pixels gx); G groups per row, G a power of two
    const int gshift = iw <= 32 ? 3 : 4;
    const int G = 1 << gshift;
    const int ntasks = ih << gshift;
    const uint32_t *tile32 = reinterpret_cast<const uint32_t *>(tile);
    uint32_t *score32 = reinterpret_cast<uint32_t *>(score);
    uint32_t *out = cell_cand + (size_t)my_slot * g.cand_per_slot + cand_off;
    const int ox = iniX + 3 - PSLOT_MINB, oy = iniY + 3 - PSLOT_MINB; // window-relative (:820-825)

    // Phase A = cv::FAST(cell, iniThFAST); phase B = cv::FAST(cell, minThFAST), only if A found nothing (:809-816).
    int n = 0;
    for (int phase = 0; phase < 2; ++phase) {
        const int thr = phase == 0 ? g.ini_th : g.min_th;
        const uint16_t *list = nullptr;
        int count = ntasks, nlist = 0;
        if (phase == 0) {
            // A.1: cheap rejection test on every group; exact scores only where a corner is possible.  Unlisted
            // groups keep score 0, which is what the maxima test needs: their R is below the threshold.
            const uint32_t bias = thr <= 127 ? (uint32_t)(127 - thr) * 0x01010101u : 0u;
            nlist = quick_pass(tile32, TSW, wsh, sh0 & 3, lane, ntasks, gshift, iw, bias, glist);
            __syncwarp();
            if (nlist == 0) continue;
            list = glist;
            count = nlist;
        }
        // exact scores: listed groups in phase A, every group in phase B (whose writes must not overtake the reads of
        // phase A's maxima test)
        __syncwarp();
        if (list && (count <= 16 || (count > 32 && count <= 48)))
            score_pass_pairs(tile32, score, TS, wsh, sh0 & 3, lane, list, count, gshift, iw);
        else
            score_pass(tile32, score32, TSW, wsh, sh0 & 3, lane, list, count, gshift, iw);
        __syncwarp();
        if (phase == 1) {
            // B.1: list (raster order) the groups that hold at least one score >= minTh, so that the neighbourhood
            // test below runs on dense warps
            const uint32_t ge_bias = thr <= 128 ? (uint32_t)(128 - thr) * 0x01010101u : 0u;
            for (int base = 0; base < ntasks; base += 32) {
                const int t = base + lane;
                bool hit = false;
                if (t < ntasks) {
                    const uint32_t s4 = score32[((t >> gshift) + 3) * TSW + (t & (G - 1)) + 1];
                    hit = (((s4 + ge_bias) | s4) & 0x80808080u) != 0; // SWAR "some byte >= thr" (false positives are harmless)
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (hit) glist[nlist + __popc(m & ((1u << lane) - 1u))] = (uint16_t)t;
                nlist += __popc(m);
            }
            __syncwarp();
        }

        // strict 3x3 maxima of the listed groups at this threshold + raster-ordered compaction into the cell's list
        const uint32_t thr1 = (uint32_t)(thr - 1) * 0x00010001u; // (t - 1) in both 16-bit halves
        int total = 0;
        for (int base = 0; base < nlist; base += 32) {
            const int j = base + lane;
            const int t = j < nlist ? (int)glist[j] : 0;
            const int y = t >> gshift, gx = t & (G - 1);
            uint32_t f = 0, s4 = 0;
            if (j < nlist) {
                const uint32_t *sw = score32 + (y + 3) * TSW + gx + 1;
                s4 = sw[0];
                // At threshold t the neighbour map is (R >= t ? R : 0) and the pixel needs R >= t, so
                // "strictly above all 8 neighbours at threshold t" == (R >= t) && (R > max of the 8 raw R).
                const uint32_t uL = sw[-TSW - 1], uM = sw[-TSW], uR = sw[-TSW + 1];
                const uint32_t cL = sw[-1], cR = sw[1];
                const uint32_t dL = sw[TSW - 1], dM = sw[TSW], dR = sw[TSW + 1];
                // words holding, for each of the 4 pixels, its neighbour one column to the left / right
                const uint32_t ul = __funnelshift_r(uL, uM, 24), ur = __funnelshift_r(uM, uR, 8);
                const uint32_t cl = __funnelshift_r(cL, s4, 24), cr = __funnelshift_r(s4, cR, 8);
                const uint32_t dl = __funnelshift_r(dL, dM, 24), dr = __funnelshift_r(dM, dR, 8);
#define EV(x) __byte_perm((x), 0, 0x4240) /* bytes 0,2 -> 16x2 */
#define OD(x) __byte_perm((x), 0, 0x4341) /* bytes 1,3 -> 16x2 */
                uint32_t ne = __vimax3_u16x2(EV(ul), EV(uM), EV(ur));
                ne = __vimax3_u16x2(ne, EV(cl), EV(cr));
                ne = __vimax3_u16x2(ne, EV(dl), EV(dM));
                ne = __vmaxu2(ne, EV(dr));
                uint32_t no = __vimax3_u16x2(OD(ul), OD(uM), OD(ur));
                no = __vimax3_u16x2(no, OD(cl), OD(cr));
                no = __vimax3_u16x2(no, OD(dl), OD(dM));
                no = __vmaxu2(no, OD(dr));
                const uint32_t se = EV(s4), so = OD(s4);
#undef EV
#undef OD
                // sign bit of each half: (nmax - R) < 0  <=>  R > nmax ;  (t - 1 - R) < 0  <=>  R >= t
                const uint32_t ke = __vsub2(ne, se) & __vsub2(thr1, se), ko = __vsub2(no, so) & __vsub2(thr1, so);
                // halves: e = pixels (0, 2), o = pixels (1, 3)  ->  4-bit flags
                const uint32_t c = ((ke & 0x80008000u) >> 1) | (ko & 0x80008000u);
                f = ((c >> 14) & 3u) | ((c >> 28) & 0xCu);
            }
            if (__ballot_sync(0xffffffffu, f != 0) == 0) continue; // nothing survives in these 32 groups (warp-uniform)
            // raster position of this lane's survivors: prefix sum of the per-lane counts (0..4) from three ballots
            // over the bits of the count - no dependent shuffle chain
            const int c = __popc(f);
            const unsigned b0 = __ballot_sync(0xffffffffu, c & 1), b1 = __ballot_sync(0xffffffffu, c & 2),
                           b2 = __ballot_sync(0xffffffffu, c & 4);
            const unsigned lt = (1u << lane) - 1u;
            int pos = total + __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
            const uint32_t vbase = pslot_pack(4 * gx + ox, y + oy, 0);
            for (uint32_t ff = f; ff; ff &= ff - 1) { // survivors of this group in x order
                const int i = __ffs((int)ff) - 1;
                if (pos < cand_stride) out[pos] = vbase + ((uint32_t)i << 8) + ((s4 >> (8 * i)) & 0xFFu);
                ++pos;
            }
            total += __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
        }
        n = total;
        if (n > 0) break;
    }
    if (lane == 0) cell_count[(size_t)my_slot * g.cells_per_slot + my_cell] = min(n, cand_stride);
    __syncwarp(); // every lane is done with the tile, the score map and the list before the next cell's boxes land
    } // work loop
}

} // namespace

// host side of the per-cell records: the cell grid of every level (ORBextractor.cc:784-806), integers throughout
void make_fast_cells(const Geom &g, FastCell *out)
{
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom &L = g.lv[l];
        for (int ci = 0; ci < L.n_rows; ++ci)
            for (int cj = 0; cj < L.n_cols; ++cj) {
                const int cell = ci * L.n_cols + cj;
                FastCell &c = out[L.cell_base + cell];
                const int iniY = PSLOT_MINB + ci * L.h_cell, iniX = PSLOT_MINB + cj * L.w_cell;
                const int maxY = iniY + L.h_cell + 6 < L.max_by ? iniY + L.h_cell + 6 : L.max_by;
                const int maxX = iniX + L.w_cell + 6 < L.max_bx ? iniX + L.w_cell + 6 : L.max_bx;
                const int tw = maxX - iniX, th = maxY - iniY;
                const bool empty = iniY >= L.max_by - 3 || iniX >= L.max_bx - 3 || tw - 6 <= 0 || th - 6 <= 0; // :795-800
                c.ini_x = (uint16_t)iniX; c.ini_y = (uint16_t)iniY;
                c.tw = empty ? 0 : (uint8_t)tw; c.th = empty ? 0 : (uint8_t)th;
                c.level = (uint8_t)l; c.box_h = (uint8_t)L.box_h;
                c.cand_off = (uint32_t)(L.cand_base + cell * L.cand_stride);
                c.cand_stride = (uint16_t)L.cand_stride;
                c.box_w = (uint16_t)L.box_w;
            }
    }
}

int launch_fast(const Geom &g, const TmapSet &maps, const FastCell *cells, int *cell_count, uint32_t *cell_cand, int nslots,
                int *work_counter, cudaStream_t st)
{
    int per_warp = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        const int tile = (g.lv[l].box_w * g.lv[l].box_h + 127) & ~127;
        const int list = (2 * 16 * (g.lv[l].box_h - 6) + 127) & ~127;
        per_warp = per_warp > 2 * tile + list ? per_warp : 2 * tile + list;
    }
    const size_t smem = (size_t)per_warp * FW;
    ensure_dyn_smem(k_fast_cells, smem);
    static int sms = 0; // same device family everywhere in a process
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int total = nslots * g.cells_per_slot;
    const int resident = sms * FAST_MIN_BLOCKS * FW; // warps the device holds at once
    const int ctas = (std::min(total, resident) + FW - 1) / FW;
    cudaMemsetAsync(work_counter, 0, sizeof(int), st);
    k_fast_cells<<<ctas, FW * 32, smem, st>>>(g, maps, reinterpret_cast<const uint4 *>(cells), cell_count, cell_cand, per_warp,
                                              work_counter, total);
    return 1;
}
