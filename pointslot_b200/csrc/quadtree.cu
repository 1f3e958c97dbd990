// quadtree.cu — K3: ORBextractor::DistributeOctTree (reference src/ORBextractor.cc:539-763) with
// ExtractorNode::DivideNode (:481-537), one CTA per (slot, level).
//
// The reference is a std::list walked and edited sequentially.  Data-parallel restatement
// (SURVEY.md Appendix E), exact by construction:
//   * Every insertion is a push_front, so list order == reverse creation order.  The node table is
//     kept in list order; one "step" (a coarse pass :600-665, or one iteration of the fine phase
//     :676-737) splits an ordered set P of nodes at once: the children of P[j], pushed in the order
//     n1..n4, land at the front in reverse push order, everything not split keeps its relative order
//     behind them.  New positions are prefix sums over (#non-empty children).
//   * A node's keys stay in candidate order (DivideNode is a stable partition), so the keys never
//     move: each key only carries the table index of its node, child occupancy is a shared-memory
//     histogram over 2-bit quadrant codes, and "first key with the maximum response" (:748-757) is an
//     atomicMax over (response, -candidate index).
//   * The fine phase sorts P by (key count, heap address) (:684) and walks it from the back, stopping
//     as soon as the list reaches N nodes (:731).  Net growth per split is (#non-empty children - 1)
//     >= 0, so the stop index is a prefix-sum threshold.  The address tie is replaced by the canonical
//     rule "the later-created node first" (DESIGN.md parity rule 1): sort key (count, creation index).
//
// Latency-bound (a few thousand keys, ~10 dependent steps); reported as time only.
#include "pslot_internal.h"

namespace {

constexpr int QT = 256; // threads per CTA
constexpr uint16_t NONE16 = 0xFFFFu;

struct NodeTab {
    uint32_t *xr;   // ulx | urx << 16
    uint32_t *yr;   // uly | bry << 16
    int *cnt;       // keys in the node
    uint16_t *born; // creation index within the step that made it
    uint8_t *isnew; // created by the most recent step
};

struct Smem {
    NodeTab tab[2];
    int *cnt4;           // [cap][4] child key counts, indexed by old table index
    uint16_t *childslot; // [cap][4] new table index of each child
    uint16_t *newpos;    // [cap]    new table index of an unsplit node
    uint16_t *ord;       // [cap]    nodes to split, in processing order
    uint16_t *pos_in_ord;// [cap]    inverse of ord (NONE16 = not in P)
    int *scan_a;         // [cap]    scratch for prefix sums
    int *scan_b;         // [cap]
    unsigned long long *sortkey; // [cap]
    uint32_t *best;      // [cap]
    uint32_t *keys;      // candidates (shared or global)
    uint16_t *knode;     // table index of each key's node
};

__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Block-wide exclusive prefix sum of a[0..n) in place (n <= a few thousand); returns the total.
// All QT threads must call.  Uses red[QT/32 + 1] ints of shared scratch.
__device__ int block_excl_scan(int *a, int n, int *red)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + QT - 1) / QT;
    const int beg = min(tid * per, n), end = min(beg + per, n);
    int sum = 0;
    for (int i = beg; i < end; ++i) sum += a[i];
    const int incl = warp_incl_scan(sum, lane);
    if (lane == 31) red[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = lane < QT / 32 ? red[lane] : 0;
        const int iv = warp_incl_scan(v, lane);
        if (lane < QT / 32) red[lane] = iv - v;
        if (lane == QT / 32 - 1) red[QT / 32] = iv;
    }
    __syncthreads();
    int run = red[warp] + incl - sum;
    for (int i = beg; i < end; ++i) {
        const int v = a[i];
        a[i] = run;
        run += v;
    }
    const int total = red[QT / 32];
    __syncthreads();
    return total;
}

__device__ __forceinline__ int quadrant(uint32_t key, uint32_t xr, uint32_t yr)
{
    const int ulx = (int)(xr & 0xFFFFu), urx = (int)(xr >> 16);
    const int uly = (int)(yr & 0xFFFFu), bry = (int)(yr >> 16);
    const int mx = ulx + ((urx - ulx + 1) >> 1); // UL.x + ceil((UR.x-UL.x)/2)  (:483)
    const int my = uly + ((bry - uly + 1) >> 1);
    return (pslot_px(key) < mx ? 0 : 1) + (pslot_py(key) < my ? 0 : 2); // n1, n2, n3, n4
}

__global__ void __launch_bounds__(QT)
k_quadtree(const __grid_constant__ Geom g, const int *__restrict__ cell_count,
           const uint32_t *__restrict__ cell_cand, int *__restrict__ cell_prefix,
           uint32_t *__restrict__ g_keys, uint16_t *__restrict__ g_knode,
           uint32_t *__restrict__ lkp, int *__restrict__ lkp_count, int *__restrict__ err_flag,
           int cap, int key_cap, int level_base)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int red[QT / 32 + 2];
    __shared__ int s_total;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = blockIdx.x + level_base, slot = blockIdx.y;
    const LevelGeom &L = g.lv[l];
    const int N = L.quota;

    // ---- carve shared memory -------------------------------------------------------------------
    Smem s;
    {
        unsigned char *p = smem_raw;
        auto take = [&](size_t bytes) { unsigned char *r = p; p += (bytes + 15) & ~size_t(15); return r; };
        s.sortkey = (unsigned long long *)take(sizeof(unsigned long long) * cap);
        for (int t = 0; t < 2; ++t) {
            s.tab[t].xr = (uint32_t *)take(4 * cap);
            s.tab[t].yr = (uint32_t *)take(4 * cap);
            s.tab[t].cnt = (int *)take(4 * cap);
            s.tab[t].born = (uint16_t *)take(2 * cap);
            s.tab[t].isnew = (uint8_t *)take(cap);
        }
        s.cnt4 = (int *)take(16 * cap);
        s.childslot = (uint16_t *)take(8 * cap);
        s.newpos = (uint16_t *)take(2 * cap);
        s.ord = (uint16_t *)take(2 * cap);
        s.pos_in_ord = (uint16_t *)take(2 * cap);
        s.scan_a = (int *)take(4 * cap);
        s.scan_b = (int *)take(4 * cap);
        s.best = (uint32_t *)take(4 * cap);
        s.keys = (uint32_t *)take(4 * (size_t)key_cap);
        s.knode = (uint16_t *)take(2 * (size_t)key_cap);
    }

    // ---- gather this level's candidates in reference order (cells row-major, raster in a cell) -----
    const int ncells = L.n_cols * L.n_rows;
    const int *cc = cell_count + (size_t)slot * g.cells_per_slot + L.cell_base;
    int *cp = cell_prefix + (size_t)slot * g.cells_per_slot + L.cell_base;
    {
        int running = 0;
        for (int base = 0; base < ncells; base += QT) {
            const int i = base + tid;
            const int v = i < ncells ? cc[i] : 0;
            const int incl = warp_incl_scan(v, lane);
            if (lane == 31) red[warp] = incl;
            __syncthreads();
            int wbase = 0, chunk = 0;
#pragma unroll
            for (int w = 0; w < QT / 32; ++w) {
                const int c = red[w];
                wbase += (w < warp) ? c : 0;
                chunk += c;
            }
            if (i < ncells) cp[i] = running + wbase + incl - v;
            running += chunk;
            __syncthreads();
        }
        if (tid == 0) s_total = running;
        __syncthreads();
    }
    const int nkeys = s_total;
    uint32_t *keys = s.keys;
    uint16_t *knode = s.knode;
    if (nkeys > key_cap) { // rare: spill the key arrays to global scratch (L2-resident)
        keys = g_keys + (size_t)slot * g.cand_per_slot + L.cand_base;
        knode = g_knode + (size_t)slot * g.cand_per_slot + L.cand_base;
    }
    {
        const uint32_t *src = cell_cand + (size_t)slot * g.cand_per_slot + L.cand_base;
        // one thread per cell: the (few) survivors of a cell sit in one or two sectors, so independent
        // threads keep hundreds of loads in flight instead of one dependent chain per warp
        for (int c = tid; c < ncells; c += QT) {
            const int n = cc[c], o = cp[c];
            const uint32_t *p = src + (size_t)c * L.cand_stride;
            for (int k = 0; k < n; ++k) keys[o + k] = p[k];
        }
    }

    // ---- roots (:543-585) ------------------------------------------------------------------------
    const int hd = L.max_by - PSLOT_MINB;
    const int nIni = L.n_ini;
    const float hX = L.hx;
    int cur = 0;
    for (int i = tid; i < nIni; i += QT) s.scan_a[i] = 0; // root key counts
    __syncthreads();
    for (int i = tid; i < nkeys; i += QT) {
        int r = (int)__fdiv_rn((float)pslot_px(keys[i]), hX); // vpIniNodes[kp.pt.x/hX] (:569)
        r = min(r, nIni - 1);
        knode[i] = (uint16_t)r;
        atomicAdd(&s.scan_a[r], 1);
    }
    __syncthreads();
    for (int i = tid; i < nIni; i += QT) { s.scan_b[i] = s.scan_a[i] > 0 ? 1 : 0; }
    __syncthreads();
    int nlive = block_excl_scan(s.scan_b, nIni, red); // erase empty roots, keep order
    for (int i = tid; i < nIni; i += QT) {
        const int c = s.scan_a[i];
        if (c > 0) {
            const int p = s.scan_b[i];
            const int ulx = (int)__fmul_rn(hX, (float)i), urx = (int)__fmul_rn(hX, (float)(i + 1)); // (:555-556)
            s.tab[0].xr[p] = (uint32_t)ulx | ((uint32_t)urx << 16);
            s.tab[0].yr[p] = 0u | ((uint32_t)hd << 16);
            s.tab[0].cnt[p] = c;
            s.tab[0].born[p] = 0;
            s.tab[0].isnew[p] = 0;
            s.newpos[i] = (uint16_t)p;
        } else {
            s.newpos[i] = NONE16;
        }
    }
    __syncthreads();
    for (int i = tid; i < nkeys; i += QT) knode[i] = s.newpos[knode[i]];
    __syncthreads();

    // ---- main loop (:594-739) --------------------------------------------------------------------
    bool fine = false;
    bool finish = false;
    int overflow = 0;
    while (!finish) {
        const NodeTab &T = s.tab[cur];
        const NodeTab &U = s.tab[cur ^ 1];
        const int prev_size = nlive;

        // P: nodes to split, in processing order
        for (int i = tid; i < nlive; i += QT) {
            const bool in = fine ? (T.isnew[i] && T.cnt[i] > 1) : (T.cnt[i] > 1);
            s.scan_a[i] = in ? 1 : 0;
            s.pos_in_ord[i] = NONE16;
        }
        __syncthreads();
        const int m = block_excl_scan(s.scan_a, nlive, red);
        if (m == 0) break; // nothing splittable: size == prevSize (:669 / :735)
        for (int i = tid; i < nlive; i += QT) {
            const bool in = fine ? (T.isnew[i] && T.cnt[i] > 1) : (T.cnt[i] > 1);
            if (in) {
                const int j = s.scan_a[i];
                s.ord[j] = (uint16_t)i; // coarse pass: list order
                // fine phase sort key: larger count first, ties -> later-created first
                s.sortkey[j] = ((unsigned long long)(uint32_t)T.cnt[i] << 32) | (unsigned long long)T.born[i] << 16 | (unsigned)i;
            }
        }
        __syncthreads();
        if (fine) {
            // rank sort, descending by (count, born); keys are unique because born is
            for (int j = tid; j < m; j += QT) {
                const unsigned long long kj = s.sortkey[j] >> 16;
                int rank = 0;
                for (int k = 0; k < m; ++k) rank += (s.sortkey[k] >> 16) > kj ? 1 : 0;
                s.scan_b[j] = rank;
            }
            __syncthreads();
            for (int j = tid; j < m; j += QT) s.ord[s.scan_b[j]] = (uint16_t)(s.sortkey[j] & 0xFFFFu);
            __syncthreads();
        }
        for (int j = tid; j < m; j += QT) s.pos_in_ord[s.ord[j]] = (uint16_t)j;
        for (int i = tid; i < nlive * 4; i += QT) s.cnt4[i] = 0;
        __syncthreads();

        // child occupancy of every node in P
        for (int i = tid; i < nkeys; i += QT) {
            const int n = knode[i];
            if (s.pos_in_ord[n] != NONE16) atomicAdd(&s.cnt4[n * 4 + quadrant(keys[i], T.xr[n], T.yr[n])], 1);
        }
        __syncthreads();

        // scan_a[j] = #non-empty children of P[j] -> exclusive prefix (push order of the children)
        for (int j = tid; j < m; j += QT) {
            const int *c = s.cnt4 + 4 * s.ord[j];
            const int ne = (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0);
            s.scan_a[j] = ne;
            s.scan_b[j] = ne; // keep the raw value
        }
        __syncthreads();
        block_excl_scan(s.scan_a, m, red);
        // number of nodes of P actually split: all of them in a coarse pass; in the fine phase stop
        // right after the first j with prev_size + sum_{k<=j}(ne_k - 1) >= N  (:731)
        int mproc = m;
        if (fine) {
            int below = 0;
            for (int base = 0; base < m; base += QT) {
                const int j = base + tid;
                const bool b = j < m && (prev_size + s.scan_a[j] + s.scan_b[j] - (j + 1) < N);
                below += __syncthreads_count(b);
            }
            mproc = min(m, below + 1);
        }
        const int total_new = s.scan_a[mproc - 1] + s.scan_b[mproc - 1];
        __syncthreads();

        // survivors keep their relative order behind the new children
        for (int i = tid; i < nlive; i += QT) {
            const uint16_t j = s.pos_in_ord[i];
            const bool split = j != NONE16 && (int)j < mproc;
            s.best[i] = split ? 0u : 1u; // reuse best[] as the keep flag
        }
        __syncthreads();
        // scan the keep flags in an int array: reuse cnt of the *other* table as scratch
        for (int i = tid; i < nlive; i += QT) U.cnt[i] = (int)s.best[i];
        __syncthreads();
        // NOTE: U.cnt is about to be overwritten with the new table, so the prefix goes to newpos first
        {
            const int nkeep = block_excl_scan(U.cnt, nlive, red);
            if (total_new + nkeep > cap) { overflow = 1; break; }
            for (int i = tid; i < nlive; i += QT)
                s.newpos[i] = s.best[i] ? (uint16_t)(total_new + U.cnt[i]) : NONE16;
            __syncthreads();
            nlive = total_new + nkeep;
        }

        // write the new table
        for (int i = tid; i < prev_size; i += QT) {
            const uint16_t p = s.newpos[i];
            if (p != NONE16) {
                U.xr[p] = T.xr[i]; U.yr[p] = T.yr[i]; U.cnt[p] = T.cnt[i];
                U.born[p] = T.born[i]; U.isnew[p] = 0;
            }
        }
        int n_expand_local = 0;
        for (int j = tid; j < mproc; j += QT) {
            const int n = s.ord[j];
            const uint32_t xr = T.xr[n], yr = T.yr[n];
            const int ulx = (int)(xr & 0xFFFFu), urx = (int)(xr >> 16);
            const int uly = (int)(yr & 0xFFFFu), bry = (int)(yr >> 16);
            const int mx = ulx + ((urx - ulx + 1) >> 1), my = uly + ((bry - uly + 1) >> 1);
            int r = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = s.cnt4[4 * n + q];
                if (c > 0) {
                    const int push = s.scan_a[j] + r; // creation index within this step
                    const int p = total_new - 1 - push;
                    const int x0 = (q & 1) ? mx : ulx, x1 = (q & 1) ? urx : mx;
                    const int y0 = (q & 2) ? my : uly, y1 = (q & 2) ? bry : my;
                    U.xr[p] = (uint32_t)x0 | ((uint32_t)x1 << 16);
                    U.yr[p] = (uint32_t)y0 | ((uint32_t)y1 << 16);
                    U.cnt[p] = c;
                    U.born[p] = (uint16_t)push;
                    U.isnew[p] = 1;
                    s.childslot[4 * n + q] = (uint16_t)p;
                    n_expand_local += c > 1;
                    ++r;
                }
            }
        }
        __syncthreads();
        // re-label the keys
        for (int i = tid; i < nkeys; i += QT) {
            const int n = knode[i];
            const uint16_t p = s.newpos[n];
            knode[i] = (p != NONE16) ? p : s.childslot[4 * n + quadrant(keys[i], T.xr[n], T.yr[n])];
        }
        // nToExpand (:603) = children with more than one key
        int n_expand = 0;
        {
            // block sum of n_expand_local
            int v = n_expand_local;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            __syncthreads();
            if (lane == 0) red[warp] = v;
            __syncthreads();
#pragma unroll
            for (int w = 0; w < QT / 32; ++w) n_expand += red[w];
            __syncthreads();
        }
        cur ^= 1;

        if (nlive >= N || nlive == prev_size) finish = true;             // (:669, :735)
        else if (!fine && nlive + n_expand * 3 > N) fine = true;         // (:673)
    }

    if (overflow) {
        if (tid == 0) atomicExch(err_flag, 1);
        nlive = 0;
    }
    __syncthreads();

    // ---- keep the first maximum-response key of every node, list order (:741-760) -----------------
    for (int i = tid; i < nlive; i += QT) s.best[i] = 0;
    __syncthreads();
    for (int i = tid; i < nkeys; i += QT)
        atomicMax(&s.best[knode[i]], ((uint32_t)pslot_pr(keys[i]) << 24) | (0xFFFFFFu - (uint32_t)i));
    __syncthreads();
    uint32_t *out = lkp + (size_t)slot * g.lkp_per_slot + L.kp_base;
    for (int i = tid; i < nlive; i += QT) {
        const uint32_t k = keys[0xFFFFFFu - (s.best[i] & 0xFFFFFFu)];
        // add the border offset (:837-847): level image coordinates
        out[i] = pslot_pack(pslot_px(k) + PSLOT_MINB, pslot_py(k) + PSLOT_MINB, pslot_pr(k));
    }
    if (tid == 0) lkp_count[(size_t)slot * g.nlevels + l] = nlive;
}

size_t smem_for(int cap, int key_cap)
{
    auto al = [](size_t b) { return (b + 15) & ~size_t(15); };
    size_t t = al(8 * (size_t)cap);
    t += 2 * (al(4 * (size_t)cap) * 3 + al(2 * (size_t)cap) + al((size_t)cap));
    t += al(16 * (size_t)cap) + al(8 * (size_t)cap) + 3 * al(2 * (size_t)cap) + 3 * al(4 * (size_t)cap);
    t += al(4 * (size_t)key_cap) + al(2 * (size_t)key_cap);
    return t;
}

} // namespace

// shared memory for the levels [l0, l1): node tables sized for the largest level of the range, key arrays for up to
// `want_keys` candidates (levels with more spill their key arrays to global scratch), within `budget` bytes per CTA
static size_t range_smem(const Geom &g, int l0, int l1, int want_keys, size_t budget, int *cap_out, int *key_cap)
{
    int cap = 0;
    for (int l = l0; l < l1; ++l) cap = cap > g.lv[l].kp_cap ? cap : g.lv[l].kp_cap;
    const size_t fixed = smem_for(cap, 0);
    int kc = want_keys;
    if (fixed + 6 * (size_t)kc > budget) kc = fixed + 6 * 1024 < budget ? (int)((budget - fixed) / 6) : 1024;
    kc = kc & ~7;
    *cap_out = cap;
    *key_cap = kc;
    return smem_for(cap, kc);
}

static size_t part_smem(const Geom &g, int part, int *l0, int *l1, int *cap, int *key_cap)
{
    const int split = g.nlevels < 2 ? g.nlevels : 2;
    *l0 = part == 0 ? 0 : split;
    *l1 = part == 0 ? split : g.nlevels;
    if (*l1 <= *l0) return 0;
    // the two finest levels hold over half of the candidates: room for 6144 keys beside the node tables, capped so
    // that at least two CTAs fit per SM; the coarse levels get small tables
    return part == 0 ? range_smem(g, *l0, *l1, 6144, 100 * 1024, cap, key_cap)
                     : range_smem(g, *l0, *l1, 2048, 48 * 1024, cap, key_cap);
}

// largest dynamic shared-memory request of the two launches (the host refuses configurations beyond the SM's limit)
size_t quadtree_smem_bytes(const Geom &g, int *key_cap)
{
    int l0, l1, cap, kc = 0, kc1 = 0;
    const size_t a = part_smem(g, 0, &l0, &l1, &cap, &kc), b = part_smem(g, 1, &l0, &l1, &cap, &kc1);
    *key_cap = kc;
    return a > b ? a : b;
}

int launch_quadtree(const Geom &g, const int *cell_count, const uint32_t *cell_cand, int *cell_prefix,
                    uint32_t *qt_keys, uint16_t *qt_nodes, uint32_t *lkp, int *lkp_count, int *err_flag,
                    int nslots, cudaStream_t st, cudaStream_t aux, cudaEvent_t ev_fork, cudaEvent_t ev_join)
{
    // Two launches: the finest levels need big tables (4 CTAs per SM); the coarse levels - most of the CTAs, each a
    // chain of a dozen barrier-separated steps - run with small tables (twice as many resident CTAs overlap their
    // barrier waits; ncu: 44 % of the stalls were barrier waits at 4 CTAs per SM) on a side stream, concurrently with
    // the fine-level launch, whose 2 x nslots long-running CTAs would otherwise leave most of the GPU idle.
    const bool fork = aux && ev_fork && ev_join && g.nlevels > 2;
    if (fork) {
        cudaEventRecord(ev_fork, st);
        cudaStreamWaitEvent(aux, ev_fork, 0);
    }
    int launches = 0;
    for (int part = 1; part >= 0; --part) { // coarse part first: it goes to the side stream
        int l0, l1, cap = 0, key_cap = 0;
        const size_t smem = part_smem(g, part, &l0, &l1, &cap, &key_cap);
        if (l1 <= l0) continue;
        ensure_dyn_smem(k_quadtree, smem);
        dim3 grid(l1 - l0, nslots);
        k_quadtree<<<grid, QT, smem, part == 1 && fork ? aux : st>>>(g, cell_count, cell_cand, cell_prefix, qt_keys, qt_nodes,
                                                                    lkp, lkp_count, err_flag, cap, key_cap, l0);
        ++launches;
    }
    if (fork) {
        cudaEventRecord(ev_join, aux);
        cudaStreamWaitEvent(st, ev_join, 0);
    }
    return launches;
}
