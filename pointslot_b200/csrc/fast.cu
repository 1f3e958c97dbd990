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
// The cell sub-image (<= 65 x 65) is staged in shared memory once; pixels that pass the 4-point
// rejection test are compacted with warp ballots so the 16-tap arc scoring runs on dense warps;
// the survivors are compacted in raster order with ballots + a warp-total scan.
//
// Integer-only, hence bit-exact.  Bound: shared-memory / ALU (each staged byte is read ~20 times);
// algorithmic global bytes per launch = S * (sum_l P_l + 4 * n_candidates + 4 * cells).
#include "pslot_internal.h"

namespace {

constexpr int FT = 256;        // threads per CTA
constexpr int TILE_MAX = 66;   // wCell, hCell <= 59 (+6 overlap)
constexpr int TILE_STRIDE = 68;

__device__ __forceinline__ int fast_raw_score(const uint8_t *t)
{
    constexpr int S = TILE_STRIDE;
    const int v = t[0];
    int d[16];
    d[0] = v - t[3 * S];      d[1] = v - t[3 * S + 1];  d[2] = v - t[2 * S + 2];  d[3] = v - t[S + 3];
    d[4] = v - t[3];          d[5] = v - t[-S + 3];     d[6] = v - t[-2 * S + 2]; d[7] = v - t[-3 * S + 1];
    d[8] = v - t[-3 * S];     d[9] = v - t[-3 * S - 1]; d[10] = v - t[-2 * S - 2]; d[11] = v - t[-S - 3];
    d[12] = v - t[-3];        d[13] = v - t[S - 3];     d[14] = v - t[2 * S - 2]; d[15] = v - t[3 * S - 1];
    int lo2[16], hi2[16], lo4[16], hi4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { lo2[k] = min(d[k], d[(k + 1) & 15]); hi2[k] = max(d[k], d[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; ++k) { lo4[k] = min(lo2[k], lo2[(k + 2) & 15]); hi4[k] = max(hi2[k], hi2[(k + 2) & 15]); }
    int a = -1000, b = 1000;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int lo9 = min(min(lo4[k], lo4[(k + 4) & 15]), d[(k + 8) & 15]); // min over d[k..k+8]
        const int hi9 = max(max(hi4[k], hi4[(k + 4) & 15]), d[(k + 8) & 15]);
        a = max(a, lo9);
        b = min(b, hi9);
    }
    return max(a, -b) - 1;
}

__global__ void __launch_bounds__(FT)
k_fast_cells(const __grid_constant__ Geom g, const __grid_constant__ PyrPtrs pyr,
             int *__restrict__ cell_count, uint32_t *__restrict__ cell_cand)
{
    __shared__ __align__(16) uint8_t tile[TILE_MAX * TILE_STRIDE];
    __shared__ __align__(16) uint8_t score[TILE_MAX * TILE_STRIDE];
    __shared__ uint16_t queue[TILE_MAX * TILE_MAX];
    __shared__ int q_count;
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

    const uint8_t *img = pyr.lv[l] + (size_t)slot * pyr.slot_stride[l] + (size_t)iniY * L.pitch + iniX;
    for (int idx = tid; idx < tw * th; idx += FT) {
        const int y = idx / tw, x = idx - y * tw;
        tile[y * TILE_STRIDE + x] = __ldg(img + (size_t)y * L.pitch + x);
        score[y * TILE_STRIDE + x] = 0;
    }
    if (tid == 0) q_count = 0;
    __syncthreads();

    // pass 1: 4-point rejection at the lower threshold; survivors go to a dense work queue
    const int npx = iw * ih;
    const int tmin = g.min_th;
    for (int base = 0; base < npx; base += FT) {
        const int idx = base + tid;
        bool pass = false;
        int off = 0;
        if (idx < npx) {
            const int y = idx / iw, x = idx - y * iw;
            off = (y + 3) * TILE_STRIDE + x + 3;
            const uint8_t *t = tile + off;
            const int v = t[0];
            const int d0 = v - t[3 * TILE_STRIDE], d8 = v - t[-3 * TILE_STRIDE], d4 = v - t[3], d12 = v - t[-3];
            // every arc of 9 holds one pixel of each diametral pair
            const bool hi = ((d0 > tmin) | (d8 > tmin)) & ((d4 > tmin) | (d12 > tmin));
            const bool lo = ((d0 < -tmin) | (d8 < -tmin)) & ((d4 < -tmin) | (d12 < -tmin));
            pass = hi | lo;
        }
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        int wbase = 0;
        if (lane == 0 && m) wbase = atomicAdd(&q_count, __popc(m));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (pass) queue[wbase + __popc(m & ((1u << lane) - 1))] = (uint16_t)off;
    }
    __syncthreads();

    // pass 2: exact arc score on the queued pixels
    const int nq = q_count;
    for (int q = tid; q < nq; q += FT) {
        const int off = queue[q];
        const int s = fast_raw_score(tile + off);
        if (s >= tmin) score[off] = (uint8_t)s;
    }
    __syncthreads();

    // pass 3: 3x3 strict-maximum test at both thresholds; flags overwrite the tile
    const int tini = g.ini_th;
    bool any_ini = false;
    for (int idx = tid; idx < npx; idx += FT) {
        const int y = idx / iw, x = idx - y * iw;
        const int off = (y + 3) * TILE_STRIDE + x + 3;
        const int s = score[off];
        uint8_t f = 0;
        if (s) {
            int nmax = 0;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx)
                    if (dx | dy) nmax = max(nmax, (int)score[off + dy * TILE_STRIDE + dx]);
            // at threshold t the neighbour map is (score >= t ? score : 0); with s >= t that changes
            // nothing for the strict comparison unless nmax < t, where the neighbours read 0
            const bool keep_min = s > nmax;
            const bool keep_ini = (s >= tini) && (s > (nmax >= tini ? nmax : 0));
            f = (uint8_t)((keep_ini ? 1 : 0) | (keep_min ? 2 : 0));
            any_ini |= keep_ini;
        }
        tile[off] = f;
    }
    const int use_ini = __syncthreads_or(any_ini);
    const uint8_t want = use_ini ? 1 : 2;

    // pass 4: raster-ordered compaction
    uint32_t *out = cell_cand + (size_t)slot * g.cand_per_slot + L.cand_base + (size_t)cell * L.cand_stride;
    const int ox = iniX + 3 - PSLOT_MINB, oy = iniY + 3 - PSLOT_MINB; // window-relative (:820-825)
    int total = 0;
    for (int base = 0; base < npx; base += FT) {
        const int idx = base + tid;
        bool keep = false;
        int x = 0, y = 0, off = 0;
        if (idx < npx) {
            y = idx / iw; x = idx - y * iw;
            off = (y + 3) * TILE_STRIDE + x + 3;
            keep = (tile[off] & want) != 0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int before = 0, chunk = 0;
#pragma unroll
        for (int w = 0; w < FT / 32; ++w) {
            const int c = warp_tot[w];
            before += (w < warp) ? c : 0;
            chunk += c;
        }
        if (keep) {
            const int pos = total + before + __popc(m & ((1u << lane) - 1));
            if (pos < L.cand_stride) out[pos] = pslot_pack(x + ox, y + oy, score[off]);
        }
        total += chunk;
        __syncthreads();
    }
    if (tid == 0) *count_out = min(total, L.cand_stride);
}

} // namespace

int launch_fast(const Geom &g, const PyrPtrs &pyr, int *cell_count, uint32_t *cell_cand, int nslots, cudaStream_t st)
{
    dim3 grid(g.cells_per_slot, nslots);
    k_fast_cells<<<grid, FT, 0, st>>>(g, pyr, cell_count, cell_cand);
    return 1;
}
