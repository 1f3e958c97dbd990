// engine.cu — host side of libpslot_orb.so: geometry tables, device buffers, launch sequencing and the
// C ABI of include/pslot_orb.h.  No CPU compute path exists here: every result comes from the kernels
// in pyramid.cu / fast.cu / quadtree.cu / describe.cu / stereo.cu.
//
// Host float code below reproduces the reference constructor arithmetic (src/ORBextractor.cc:410-470,
// 1107-1132, 765-787, 543-545); build with -ffp-contract=off so no expression is fused.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <atomic>
#include <new>
#include <map>
#include <vector>

#include "pslot_internal.h"
#include "tma.cuh"

namespace {

thread_local char t_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(PSLOT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                        \
    } while (0)

inline int cv_round(double v) { return (int)lrint(v); } // cvRound: half-to-even

struct Tables {
    float sf[PSLOT_MAX_LEVELS], inv_sf[PSLOT_MAX_LEVELS], sigma2[PSLOT_MAX_LEVELS], inv_sigma2[PSLOT_MAX_LEVELS];
    int per_level[PSLOT_MAX_LEVELS];
};

struct Params {
    int nfeatures;
    float scale_factor;
    int nlevels, ini_th, min_th;
};

// ORBextractor::ORBextractor (ORBextractor.cc:410-446)
void make_tables(const Params &p, Tables &t)
{
    const double scaleFactor = (double)p.scale_factor; // ORBextractor.h:98: double member
    t.sf[0] = 1.0f;
    t.sigma2[0] = 1.0f;
    for (int i = 1; i < p.nlevels; ++i) {
        t.sf[i] = (float)((double)t.sf[i - 1] * scaleFactor);
        t.sigma2[i] = t.sf[i] * t.sf[i];
    }
    for (int i = 0; i < p.nlevels; ++i) {
        t.inv_sf[i] = 1.0f / t.sf[i];
        t.inv_sigma2[i] = 1.0f / t.sigma2[i];
    }
    const float factor = (float)(1.0 / scaleFactor);
    float nDesired = (float)p.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)p.nlevels));
    int sum = 0;
    for (int l = 0; l < p.nlevels - 1; ++l) {
        t.per_level[l] = cv_round((double)nDesired);
        sum += t.per_level[l];
        nDesired *= factor;
    }
    t.per_level[p.nlevels - 1] = std::max(p.nfeatures - sum, 0);
}

// taps of cv::resize INTER_LINEAR for one axis (OpenCV resize.cpp, 8-bit path)
void resize_taps(int n_dst, int n_src, ResizeTap *out, int padded)
{
    const double scale = 1.0 / ((double)n_dst / (double)n_src);
    for (int d = 0; d < n_dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= n_src - 1) { s = n_src - 1; f = 0.f; }
        out[d].s0 = (uint16_t)s;
        out[d].s1 = (uint16_t)std::min(s + 1, n_src - 1);
        out[d].c0 = (int16_t)cv_round((double)((1.f - f) * 2048.f));
        out[d].c1 = (int16_t)cv_round((double)(f * 2048.f));
    }
    for (int d = n_dst; d < padded; ++d) out[d] = out[n_dst - 1];
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Engine {
    Params p{};
    Tables t{};
    int device = 0;
    cudaStream_t st = nullptr;
    bool own_stream = false;
    // side stream for the coarse-level quadtree launch, forked from / joined to `st` inside every submit
    cudaStream_t st_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    int W = 0, H = 0, S = 0; // configured geometry / slot capacity
    Geom g{};
    PyrPtrs pyr{};
    TmapSet maps{};
    PyrMaps pmaps{};
    bool has_pyramid = false;

    uint8_t *d_pyr = nullptr;
    ResizeTap *d_taps = nullptr;
    int *d_cell_count = nullptr, *d_cell_prefix = nullptr;
    FastCell *d_fast_cells = nullptr;
    int *d_fast_work = nullptr; // work counter of the persistent FAST warps
    uint32_t *d_cell_cand = nullptr, *d_qt_keys = nullptr;
    uint16_t *d_qt_nodes = nullptr;
    uint32_t *d_lkp = nullptr;
    int *d_lkp_count = nullptr;
    int *h_err = nullptr, *d_err = nullptr; // error flag in host-mapped pinned memory: no copy needed to read it
    // outputs: one block [counts | kps | desc] so a whole batch comes back in few copies
    unsigned char *d_out = nullptr;
    size_t out_bytes = 0, off_kps = 0, off_desc = 0;
    int *d_out_count = nullptr;
    pslot_keypoint *d_out_kps = nullptr;
    uint8_t *d_out_desc = nullptr;
    // stereo results, [S/2 (>=1)][K]
    float *d_uright = nullptr, *d_depth = nullptr;
    int *d_best_dist = nullptr, *d_best_idx = nullptr, *d_sad = nullptr;

    ~Engine()
    {
        release();
        if (st_aux) cudaStreamDestroy(st_aux);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (own_stream && st) cudaStreamDestroy(st);
    }

    void release()
    {
        cudaFree(d_pyr); cudaFree(d_taps); cudaFree(d_cell_count); cudaFree(d_cell_prefix); cudaFree(d_fast_cells); cudaFree(d_fast_work);
        cudaFree(d_cell_cand); cudaFree(d_qt_keys); cudaFree(d_qt_nodes); cudaFree(d_lkp);
        cudaFree(d_lkp_count); cudaFree(d_out);
        if (h_err) cudaFreeHost(h_err);
        cudaFree(d_uright); cudaFree(d_depth); cudaFree(d_best_dist); cudaFree(d_best_idx); cudaFree(d_sad);
        d_pyr = nullptr; d_taps = nullptr; d_cell_count = d_cell_prefix = nullptr; d_fast_cells = nullptr; d_fast_work = nullptr; d_cell_cand = d_qt_keys = nullptr;
        d_qt_nodes = nullptr; d_lkp = nullptr; d_lkp_count = nullptr; h_err = d_err = nullptr; d_out = nullptr;
        d_uright = d_depth = nullptr; d_best_dist = d_best_idx = d_sad = nullptr;
        W = H = S = 0; has_pyramid = false;
    }

    int init(const Params &pp, int dev, void *stream)
    {
        if (pp.nlevels < 1 || pp.nlevels > PSLOT_MAX_LEVELS) return fail(PSLOT_ERR_ARG, "nlevels must be in [1,%d]", PSLOT_MAX_LEVELS);
        if (pp.nfeatures < 1 || pp.nfeatures > 30000) return fail(PSLOT_ERR_ARG, "nfeatures must be in [1,30000]");
        if (!(pp.scale_factor > 1.0f) || pp.scale_factor > 1.45f)
            return fail(PSLOT_ERR_ARG, "scaleFactor must be in (1, 1.45] (the pyramid kernel stages <= 1.49 source px per output px)");
        if (pp.min_th < 1 || pp.ini_th < pp.min_th || pp.ini_th > 254)
            return fail(PSLOT_ERR_ARG, "need 1 <= minThFAST <= iniThFAST <= 254");
        p = pp;
        device = dev;
        make_tables(p, t);
        int ndev = 0;
        CU(cudaGetDeviceCount(&ndev));
        if (dev < 0 || dev >= ndev) return fail(PSLOT_ERR_CUDA, "CUDA device %d not available (%d visible)", dev, ndev);
        CU(cudaSetDevice(dev));
        if (stream) { st = (cudaStream_t)stream; own_stream = false; }
        else { CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); own_stream = true; }
        CU(cudaStreamCreateWithFlags(&st_aux, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        return PSLOT_OK;
    }

    // Geometry of every level for a w x h image (ORBextractor.cc:1111-1112, 773-787, 543-545).
    int make_geom(int w, int h, Geom &G, std::vector<ResizeTap> &taps) const
    {
        if (w > 4096 || h > 4096) return fail(PSLOT_ERR_SHAPE, "image %dx%d exceeds 4096 px", w, h);
        memset(&G, 0, sizeof(G));
        G.nlevels = p.nlevels; G.ini_th = p.ini_th; G.min_th = p.min_th;
        int cells = 0, cand = 0, lkp = 0, maxcap = 0;
        taps.clear();
        for (int l = 0; l < p.nlevels; ++l) {
            LevelGeom &L = G.lv[l];
            L.w = cv_round((double)((float)w * t.inv_sf[l]));
            L.h = cv_round((double)((float)h * t.inv_sf[l]));
            L.pitch = (int)align_up((size_t)std::max(L.w, 1), 128);
            L.max_bx = L.w - PSLOT_EDGE + 3;
            L.max_by = L.h - PSLOT_EDGE + 3;
            const float width = (float)(L.max_bx - PSLOT_MINB), height = (float)(L.max_by - PSLOT_MINB);
            L.n_cols = (int)(width / 30.f);
            L.n_rows = (int)(height / 30.f);
            if (L.n_cols < 1 || L.n_rows < 1)
                return fail(PSLOT_ERR_SHAPE, "level %d of a %dx%d image is %dx%d: below the 62x62 minimum", l, w, h, L.w, L.h);
            L.w_cell = (int)ceilf(width / (float)L.n_cols);
            L.h_cell = (int)ceilf(height / (float)L.n_rows);
            if (L.w_cell > 59 || L.h_cell > 59) return fail(PSLOT_ERR_SHAPE, "cell %dx%d too large", L.w_cell, L.h_cell);
            L.box_w = (int)align_up((size_t)L.w_cell + 30, 16);
            L.box_h = L.h_cell + 6;
            L.cell_base = cells;
            cells += L.n_cols * L.n_rows;
            L.cand_stride = ((L.w_cell + 1) / 2) * ((L.h_cell + 1) / 2);
            L.cand_base = cand;
            cand += L.n_cols * L.n_rows * L.cand_stride;
            L.quota = t.per_level[l];
            const int wd = L.max_bx - PSLOT_MINB, hd = L.max_by - PSLOT_MINB;
            L.n_ini = (int)roundf((float)wd / (float)hd);
            if (L.n_ini < 1) return fail(PSLOT_ERR_SHAPE, "level %d aspect ratio gives no quadtree root", l);
            L.hx = (float)wd / (float)L.n_ini;
            L.kp_cap = std::max(L.quota + 3, 4 * L.n_ini) + 1;
            L.kp_base = lkp;
            lkp += L.kp_cap;
            maxcap = std::max(maxcap, L.kp_cap);
            L.scale = t.sf[l];
            L.inv_scale = t.inv_sf[l];
            L.patch_size = (int)(31.f * t.sf[l]);
            if (l > 0) {
                const LevelGeom &P = G.lv[l - 1];
                const int px = (int)align_up((size_t)L.w, 4), py = (int)align_up((size_t)L.h, 4);
                L.tab_x = (int)taps.size();
                taps.resize(taps.size() + px);
                resize_taps(L.w, P.w, taps.data() + L.tab_x, px);
                L.tab_y = (int)taps.size();
                taps.resize(taps.size() + py);
                resize_taps(L.h, P.h, taps.data() + L.tab_y, py);
                // per-level ratios after cvRound can exceed the nominal scale factor (48/33 = 1.4545): check the real taps
                if (!pyramid_tiles_fit(taps.data() + L.tab_x, L.w, taps.data() + L.tab_y, L.h))
                    return fail(PSLOT_ERR_SHAPE, "level %d (%dx%d from %dx%d): resize ratio exceeds the pyramid kernel's staging tile", l,
                                L.w, L.h, P.w, P.h);
            }
        }
        G.cells_per_slot = cells; G.cand_per_slot = cand; G.lkp_per_slot = lkp; G.kp_cap_total = lkp;
        G.max_level_cap = maxcap;
        int key_cap = 0;
        if (quadtree_smem_bytes(G, &key_cap) > 200 * 1024)
            return fail(PSLOT_ERR_ARG, "nfeatures too large for the on-chip quadtree tables");
        return PSLOT_OK;
    }

    // One (x, y, slot) u8 tensor map per pyramid level; box = the FAST cell tile.  Out-of-image box
    // elements are zero-filled by the TMA unit.
    int make_tensor_maps(int nslots, const std::vector<ResizeTap> &taps)
    {
        typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        // resolved once through the runtime (no link-time dependency on libcuda); function-local static
        // initialisation is thread-safe, and two extractor contexts may configure concurrently (Frame.cc:161-164)
        static const EncodeFn encode = []() -> EncodeFn {
            void *fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
                q != cudaDriverEntryPointSuccess)
                return nullptr;
            return (EncodeFn)fn;
        }();
        if (!encode) return fail(PSLOT_ERR_CUDA, "cuTensorMapEncodeTiled unavailable in this driver");
        for (int l = 0; l < g.nlevels; ++l) {
            const LevelGeom &L = g.lv[l];
            const cuuint64_t dims[3] = {(cuuint64_t)L.w, (cuuint64_t)L.h, (cuuint64_t)nslots};
            const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)pyr.slot_stride[l]};
            const cuuint32_t box[3] = {(cuuint32_t)L.box_w, (cuuint32_t)L.box_h, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = encode(&maps.lv[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, pyr.lv[l], dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(PSLOT_ERR_CUDA, "cuTensorMapEncodeTiled failed for level %d (CUresult %d)", l, (int)r);
            // the pyramid kernel's views of the same level: destination tile, and source box of the next level's tiles
            const cuuint32_t obox[3] = {PYR_TILE_W, PYR_TILE_H, 1};
            r = encode(&pmaps.dst[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, pyr.lv[l], dims, strides, obox, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS && l + 1 < g.nlevels) {
                const LevelGeom &D = g.lv[l + 1];
                pyramid_src_box(taps.data() + D.tab_x, D.w, taps.data() + D.tab_y, D.h, &pmaps.box_w[l], &pmaps.box_h[l]);
                const cuuint32_t sbox[3] = {(cuuint32_t)pmaps.box_w[l], (cuuint32_t)pmaps.box_h[l], 1};
                r = encode(&pmaps.src[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, pyr.lv[l], dims, strides, sbox, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            }
            if (r != CUDA_SUCCESS) return fail(PSLOT_ERR_CUDA, "cuTensorMapEncodeTiled (pyramid) failed for level %d (CUresult %d)", l, (int)r);
        }
        return PSLOT_OK;
    }

    int configure(int w, int h, int nslots)
    {
        if (w == W && h == H && nslots <= S) return PSLOT_OK;
        CU(cudaSetDevice(device));
        Geom G;
        std::vector<ResizeTap> taps;
        int rc = make_geom(w, h, G, taps);
        if (rc) return rc;
        CU(cudaStreamSynchronize(st));
        release();
        ++generation; // every device pointer a captured graph may hold is gone
        g = G;
        size_t off = 0;
        size_t lvoff[PSLOT_MAX_LEVELS];
        for (int l = 0; l < g.nlevels; ++l) {
            lvoff[l] = off;
            pyr.slot_stride[l] = (unsigned long long)g.lv[l].pitch * g.lv[l].h;
            off += align_up((size_t)pyr.slot_stride[l] * nslots, 256);
        }
        CU(cudaMalloc(&d_pyr, off + 256));
        CU(cudaMemset(d_pyr, 0, off + 256)); // row padding is read (never used) by 16-byte loads: keep it defined
        for (int l = 0; l < g.nlevels; ++l) pyr.lv[l] = d_pyr + lvoff[l];
        rc = make_tensor_maps(nslots, taps);
        if (rc) return rc;
        CU(cudaMalloc(&d_taps, sizeof(ResizeTap) * std::max<size_t>(taps.size(), 4)));
        if (!taps.empty()) CU(cudaMemcpy(d_taps, taps.data(), sizeof(ResizeTap) * taps.size(), cudaMemcpyHostToDevice));
        const size_t ns = (size_t)nslots;
        CU(cudaMalloc(&d_cell_count, 4 * ns * g.cells_per_slot));
        CU(cudaMalloc(&d_cell_prefix, 4 * ns * g.cells_per_slot));
        {
            std::vector<FastCell> fc((size_t)g.cells_per_slot);
            make_fast_cells(g, fc.data());
            CU(cudaMalloc(&d_fast_cells, sizeof(FastCell) * fc.size()));
            CU(cudaMalloc(&d_fast_work, sizeof(int)));
            CU(cudaMemcpy(d_fast_cells, fc.data(), sizeof(FastCell) * fc.size(), cudaMemcpyHostToDevice));
        }
        CU(cudaMalloc(&d_cell_cand, 4 * ns * g.cand_per_slot));
        CU(cudaMalloc(&d_qt_keys, 4 * ns * g.cand_per_slot));
        CU(cudaMalloc(&d_qt_nodes, 2 * ns * g.cand_per_slot));
        CU(cudaMalloc(&d_lkp, 4 * ns * g.lkp_per_slot));
        CU(cudaMalloc(&d_lkp_count, 4 * ns * g.nlevels));
        CU(cudaHostAlloc((void **)&h_err, sizeof(int), cudaHostAllocMapped));
        *h_err = 0;
        CU(cudaHostGetDevicePointer((void **)&d_err, h_err, 0));
        const size_t K = (size_t)g.kp_cap_total;
        off_kps = align_up(4 * ns, 256);
        off_desc = off_kps + align_up(sizeof(pslot_keypoint) * ns * K, 256);
        out_bytes = off_desc + align_up(32 * ns * K, 256);
        CU(cudaMalloc(&d_out, out_bytes));
        CU(cudaMemset(d_out, 0, out_bytes)); // result rows are copied out at full capacity; the unused tail reads as zeros
        d_out_count = (int *)d_out;
        d_out_kps = (pslot_keypoint *)(d_out + off_kps);
        d_out_desc = (uint8_t *)(d_out + off_desc);
        const size_t np = std::max<size_t>(ns / 2, 1);
        CU(cudaMalloc(&d_uright, 4 * np * K)); CU(cudaMalloc(&d_depth, 4 * np * K));
        CU(cudaMemset(d_uright, 0, 4 * np * K)); CU(cudaMemset(d_depth, 0, 4 * np * K));
        CU(cudaMalloc(&d_best_dist, 4 * np * K)); CU(cudaMalloc(&d_best_idx, 4 * np * K)); CU(cudaMalloc(&d_sad, 4 * np * K));
        W = w; H = h; S = nslots;
        return PSLOT_OK;
    }

    // K1..K6 on slots [0, nslots); level 0 must already be in place.  Returns kernels launched (>0) or <0.
    int generation = 0; // bumped by every (re)configuration
    int run_extract(int nslots, cudaEvent_t *ev = nullptr)
    {
        int n = 0;
        if (ev) cudaEventRecord(ev[1], st);
        n += launch_pyramid(g, pyr, pmaps, d_taps, nslots, d_err, st);
        if (ev) cudaEventRecord(ev[2], st);
        n += launch_fast(g, maps, d_fast_cells, d_cell_count, d_cell_cand, nslots, d_fast_work, st);
        if (ev) cudaEventRecord(ev[3], st);
        n += launch_quadtree(g, d_cell_count, d_cell_cand, d_cell_prefix, d_qt_keys, d_qt_nodes, d_lkp, d_lkp_count,
                             d_err, nslots, st, st_aux, ev_fork, ev_join);
        if (ev) cudaEventRecord(ev[4], st);
        n += launch_describe(g, pyr, d_lkp, d_lkp_count, d_out_kps, d_out_desc, d_out_count, nslots, st);
        if (ev) cudaEventRecord(ev[5], st);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(PSLOT_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
        has_pyramid = true;
        return n;
    }
};

int check_err_flag(Engine &E)
{
    CU(cudaStreamSynchronize(E.st));
    const int f = *(volatile int *)E.h_err;
    if (f) {
        *E.h_err = 0;
        if (f & PSLOT_FLAG_QT_OVERFLOW) return fail(PSLOT_ERR_CUDA, "quadtree node table overflow (internal capacity bound violated)");
        if (f & PSLOT_FLAG_PYR_TILE) return fail(PSLOT_ERR_SHAPE, "a pyramid resize tile exceeded the staging buffer (level not written)");
        if (f & PSLOT_FLAG_STEREO_NR) return fail(PSLOT_ERR_CAPACITY, "stereo matcher: more right keypoints than it can stage (results would be truncated)");
        if (f & PSLOT_FLAG_STEREO_OCTAVE) return fail(PSLOT_ERR_ARG, "stereo matcher: a right keypoint's octave is out of range");
        return fail(PSLOT_ERR_CUDA, "device error flag 0x%x", f);
    }
    return PSLOT_OK;
}

} // namespace

// -------------------------------------------------------------------------------------------------
struct pslot_orb {
    Engine E;
    // device scratch for host-provided feature sets (pslot_stereo_match)
    pslot_keypoint *d_kl = nullptr, *d_kr = nullptr;
    uint8_t *d_dl = nullptr, *d_dr = nullptr;
    int *d_n = nullptr;
    float *d_ur = nullptr, *d_dp = nullptr;
    int *d_bd = nullptr, *d_bi = nullptr, *d_sd = nullptr;
    int st_cap = 0;
    int last_nl = 0;
    unsigned char *h_stage = nullptr; // pinned mirror of E.d_out
    size_t h_stage_bytes = 0;
    // single-frame path: kernels + result download replayed as one CUDA graph (13 launches + 1 copy -> 1 launch); built on
    // the second frame of a geometry (the first one sets the kernels' attributes), dropped when the engine re-configures
    cudaGraphExec_t extract_graph = nullptr;
    int extract_graph_gen = -1, extract_calls_gen = -1, extract_calls = 0;
    bool extract_pending = false; // pslot_orb_extract_begin ran, pslot_orb_extract_end has not
    uint64_t extract_token = 0;   // identity of the feature set the last extraction left on the device (0: none)
    float *h_stereo = nullptr;    // pinned landing zone of pslot_stereo_match_resident (uRight | depth)
    size_t h_stereo_cap = 0;
    unsigned char *h_match = nullptr, *d_match = nullptr; // pslot_stereo_match: counts | kpsL | kpsR | descL | descR (+ results on the host side)
    size_t h_match_cap = 0;
    uint8_t *d_cstage = nullptr;      // tightly packed frame (colour, or gray at any pitch-free layout) awaiting conversion / re-pitch
    size_t cstage_bytes = 0;
    // feature grid of the last pslot_assign_grid (kept for pslot_features_in_area)
    pslot_keypoint *d_gk = nullptr;
    int *d_gidx = nullptr, *d_gcs = nullptr, *d_qcount = nullptr, *d_qout = nullptr;
    pslot_area_query *d_q = nullptr;
    int gk_cap = 0, gcs_cap = 0, q_cap = 0;
    size_t qout_cap = 0;
    int grid_n = -1;
    uint64_t grid_token = 0; // identity of the gridded keypoint set (pslot_grid_token); 0 = none
    pslot_grid_spec grid_spec = {};
    std::vector<float> h_gangle; // mvKeysUn[i].angle of the gridded frame (rotation check of the last-frame search)
    // projection search scratch (pslot_search_by_projection)
    void *d_proj = nullptr;
    size_t proj_bytes = 0;
    int last_sweeps = 0;
    // pslot_assign_features scratch: one arena (keypoints | depth | ids | cls | order | class_start | effective) + the mask
    unsigned char *d_part = nullptr;
    size_t part_bytes = 0;
    uint8_t *d_pmask = nullptr;
    size_t pmask_bytes = 0;
    int pmask_w = 0, pmask_h = 0;

    ~pslot_orb()
    {
        if (extract_graph) cudaGraphExecDestroy(extract_graph);
        if (h_stereo) cudaFreeHost(h_stereo);
        if (h_match) cudaFreeHost(h_match);
        cudaFree(d_match);
        cudaFree(d_part); cudaFree(d_pmask);
        free_stereo();
        cudaFree(d_gk); cudaFree(d_gidx); cudaFree(d_gcs); cudaFree(d_qcount); cudaFree(d_qout); cudaFree(d_q);
        cudaFree(d_proj);
        cudaFree(d_cstage);
        if (h_stage) cudaFreeHost(h_stage);
    }
    void free_stereo()
    {
        cudaFree(d_kl); cudaFree(d_kr); cudaFree(d_dl); cudaFree(d_dr); cudaFree(d_n);
        cudaFree(d_ur); cudaFree(d_dp); cudaFree(d_bd); cudaFree(d_bi); cudaFree(d_sd);
        d_kl = d_kr = nullptr; d_dl = d_dr = nullptr; d_n = nullptr; d_ur = d_dp = nullptr; d_bd = d_bi = d_sd = nullptr;
        st_cap = 0;
    }
};

struct pslot_batch {
    Engine E;
    int max_pairs = 0;
    int last_pairs = 0;
    unsigned long long submits = 0; // identity of what the pyramids hold
    int launches = 0;
    bool profiling = false, profiled = false;
    cudaEvent_t ev[PSLOT_NUM_STAGES + 1] = {};
    int *d_gcs = nullptr, *d_gidx = nullptr; // feature grids of the last submit (pslot_batch_assign_grid)
    unsigned char *d_part = nullptr;         // pslot_batch_assign_features: masks | ids | cls | order | class_start | effective
    size_t part_bytes = 0;
    int gcs_cells = 0;
    unsigned long long *d_counters = nullptr; // stereo work counters, accumulated by profiled submits (pslot_batch_stereo_counters)
    uint8_t *d_stage = nullptr; // tightly packed landing zone for host images (one linear H2D per view)
    uint8_t *d_cstage = nullptr; // same for colour frames
    size_t cstage_bytes = 0;
    // optional shared upload stream: several batch contexts that copy concurrently on their own streams make the
    // DMA engine interleave their transfers (measured: 35 GB/s instead of 55 GB/s); funnelling every upload through
    // one stream keeps each transfer at the full PCIe rate.  Events order upload -> ingest kernel -> next upload.
    cudaStream_t copy_st = nullptr;
    cudaEvent_t ev_uploaded = nullptr, ev_stage_free = nullptr;
    bool stage_busy = false;
    ~pslot_batch()
    {
        if (ev_uploaded) cudaEventDestroy(ev_uploaded);
        if (ev_stage_free) cudaEventDestroy(ev_stage_free);
        for (auto &e : ev)
            if (e) cudaEventDestroy(e);
        cudaFree(d_stage);
        cudaFree(d_cstage);
        cudaFree(d_counters);
        cudaFree(d_gcs);
        cudaFree(d_part);
        cudaFree(d_gidx);
    }
};


// ---- cv::ORB (OpenCV's own extractor; reference Frame.cc:2623-2627) ---------------------------------------------------
struct pslot_cvorb {
    int nfeatures = 0, nlevels = 0, edge = 0, fast_th = 0, device = 0;
    double scale_factor = 1.2;
    cudaStream_t st = nullptr, st_up = nullptr; // kernels + downloads / uploads (a batch is uploaded in chunks that run ahead)
    cudaEvent_t ev_up[4] = {nullptr, nullptr, nullptr, nullptr};
    int W = 0, H = 0, slots = 0; // geometry and number of image slots the buffers are sized for
    CvGeom g{};
    CvBuffers b{};
    int *h_err = nullptr;
    size_t level_bytes = 0;
    int cand_total = 0;
    unsigned char *h_out = nullptr; // pinned staging of the outputs
    size_t h_out_bytes = 0;
    uint8_t *d_stage = nullptr;     // tightly packed uploads (images, then masks) before the re-pitch kernel
    size_t stage_view = 0;          // bytes of one of its two halves
    double t_phase[4] = {0, 0, 0, 0}; // host clock of the last call: uploads enqueued, kernels enqueued, wait, copy-out (ms)
    // the kernels + result downloads of an un-chunked call as one CUDA graph per (images, masks, tight input): ~20 launches
    // become one (the per-frame pair of Frame::ExtractObjORB is latency-bound); dropped whenever a buffer they name moves
    // what the slots hold after pslot_cvorb_detect_and_compute_resident (for pslot_cvorb_stereo_match_resident)
    const pslot_batch *res_batch = nullptr;
    unsigned long long res_submits = 0;
    int res_first = 0, res_n = 0;
    unsigned char *d_stereo = nullptr; // uRight | depth | best_dist | best_idx | sad, [pairs][slot_kp] each
    size_t stereo_bytes = 0;
    std::map<int, cudaGraphExec_t> graphs;
    void drop_graphs()
    {
        for (auto &kv : graphs) cudaGraphExecDestroy(kv.second);
        graphs.clear();
    }

    void release()
    {
        drop_graphs();
        cudaFree(d_stage); d_stage = nullptr;
        cudaFree(b.pyr); cudaFree(b.mask); cudaFree(b.score); cudaFree((void *)b.taps); cudaFree(b.cand); cudaFree(b.kept);
        cudaFree(b.final_kp); cudaFree(b.kept_resp); cudaFree(b.final_resp); cudaFree(b.counters); cudaFree(b.hist);
        cudaFree(b.out_kps); cudaFree(b.out_desc); cudaFree(b.out_count);
        b = CvBuffers{};
        W = H = slots = 0;
    }
    ~pslot_cvorb()
    {
        release();
        if (h_err) cudaFreeHost(h_err);
        if (h_out) cudaFreeHost(h_out);
        cudaFree(d_stereo);
        for (cudaEvent_t e : ev_up)
            if (e) cudaEventDestroy(e);
        if (st_up) cudaStreamDestroy(st_up);
        if (st) cudaStreamDestroy(st);
    }
};

// ---- ORB vocabulary (Frame::ComputeBoW; reference Frame.cc:2040-2047, Thirdparty/DBoW2) ------------------------------------
struct pslot_vocab {
    int n_nodes = 0, k = 0, L = 0, scoring = 0, weighting = 0, device = 0;
    std::vector<int32_t> word_id;  // host copies of what pslot_bow_vectors needs
    std::vector<double> weight;
    uint8_t *d_desc = nullptr;
    int *d_child_start = nullptr, *d_child = nullptr;
    // per-call scratch
    uint8_t *d_feat = nullptr;
    int *d_leaf = nullptr, *d_nid = nullptr;
    int cap = 0;
    cudaStream_t st = nullptr;
    ~pslot_vocab()
    {
        cudaFree(d_desc); cudaFree(d_child_start); cudaFree(d_child); cudaFree(d_feat); cudaFree(d_leaf); cudaFree(d_nid);
        if (st) cudaStreamDestroy(st);
    }
};

namespace {

// INTER_LINEAR_EXACT taps of one axis (OpenCV resize.cpp, 8-bit fixed point): source index and fraction * 256
void exact_taps(int n_dst, int n_src, int2 *out)
{
    const double scale = (double)n_src / (double)n_dst;
    for (int d = 0; d < n_dst; ++d) {
        double f = (d + 0.5) * scale - 0.5;
        int s = (int)floor(f);
        double fr = f - s;
        if (s < 0) { s = 0; fr = 0.0; }
        if (s >= n_src - 1) { s = n_src - 1; fr = 0.0; }
        out[d].x = s;
        out[d].y = (int)lrint(fr * 256.0);
    }
}

// the buffers of slots s0.. as a CvBuffers of their own (a chunk of a batch)
CvBuffers cv_slot_view(const CvGeom &g, const CvBuffers &b, int s0)
{
    CvBuffers v = b;
    const size_t s = (size_t)s0;
    v.pyr += s * g.slot_bytes; v.mask += s * g.slot_bytes; v.score += s * g.slot_bytes;
    v.cand += s * g.slot_cand; v.kept += s * g.slot_cand; v.kept_resp += s * g.slot_cand;
    v.final_kp += s * g.slot_kp; v.final_resp += s * g.slot_kp;
    v.out_kps += s * g.slot_kp; v.out_desc += s * g.slot_kp * 32; v.out_count += s;
    v.counters += s * CV_CNT_TOTAL; v.hist += s * 256 * PSLOT_MAX_LEVELS;
    return v;
}

int cvorb_configure(pslot_cvorb &c, int w, int h, int slots)
{
    if (w == c.W && h == c.H && slots <= c.slots) return PSLOT_OK;
    if (w > 4095 || h > 4095) return fail(PSLOT_ERR_SHAPE, "image %dx%d exceeds 4095 px", w, h);
    if (w == c.W && h == c.H) slots = std::max(slots, std::min(2 * c.slots, 1024)); // growing batch: amortise the re-allocation
    CU(cudaStreamSynchronize(c.st_up));
    CU(cudaStreamSynchronize(c.st));
    c.release();
    CvGeom &G = c.g;
    memset(&G, 0, sizeof(G));
    G.edge = c.edge; G.fast_th = c.fast_th;
    { // getGaussianKernel(7, 2, CV_32F): exp(-x^2 / (2 sigma^2)) normalised in double, stored as float
        double gk[7], sum = 0;
        for (int i = 0; i < 7; ++i) { const double x = i - 3; gk[i] = exp(-0.125 * x * x); sum += gk[i]; }
        for (int i = 0; i < 4; ++i) G.gk[i] = (float)(gk[i] / sum);
    }
    // feature quota per level (orb.cpp computeKeyPoints)
    int quota[PSLOT_MAX_LEVELS];
    {
        const float factor = (float)(1.0 / c.scale_factor);
        float nd = c.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)c.nlevels));
        int sum = 0;
        for (int l = 0; l < c.nlevels - 1; ++l) { quota[l] = cv_round((double)nd); sum += quota[l]; nd *= factor; }
        quota[c.nlevels - 1] = std::max(c.nfeatures - sum, 0);
    }
    std::vector<int2> taps;
    size_t off = 0;
    int cand = 0, nl = 0, tiles = 0, max_quota = 1;
    for (int l = 0; l < c.nlevels; ++l) {
        const float scale = (float)pow(c.scale_factor, (double)l);
        // level size = cvRound(dim * (1.f / scale)): one float reciprocal, then a float product - what OpenCV 4.13 (cv2) is observed
        // to compute; a float division differs where dim / scale lands on x.5 (465 / 1.2f -> 387, 465 * (1 / 1.2f) -> 388).
        // Pinned against cv2.ORB_create by tools/cvorb_fuzz.py
        const float inv_scale = 1.0f / scale;
        const int lw = cv_round((double)((float)w * inv_scale)), lh = cv_round((double)((float)h * inv_scale));
        if (lw < 2 * c.edge + 1 || lh < 2 * c.edge + 1) break; // nothing can pass the border filter from here on
        CvLevel &L = G.lv[l];
        L.w = lw; L.h = lh; L.pitch = (int)align_up((size_t)lw, 128);
        L.off = off;
        off += align_up((size_t)L.pitch * lh, 256);
        L.quota = quota[l];
        L.scale = scale;
        L.cand_base = cand;
        L.cand_cap = ((lw + 1) / 2) * ((lh + 1) / 2); // strict 3x3 maxima cannot be denser
        cand += L.cand_cap;
        L.kp_base = l * CV_SORT_CAP;
        // only scores within one pixel of the border filter's interior can matter: rows / columns [edge - 1, dim - edge]
        const int band_rows = lh - 2 * c.edge + 2;
        L.tile_base = tiles;
        L.tiles_x = (lw + 127) / 128;
        tiles += L.tiles_x * ((band_rows + 31) / 32);
        max_quota = std::max(max_quota, quota[l]);
        if (l > 0) {
            L.tab_x = (int)taps.size(); taps.resize(taps.size() + lw);
            exact_taps(lw, G.lv[l - 1].w, taps.data() + L.tab_x);
            L.tab_y = (int)taps.size(); taps.resize(taps.size() + lh);
            exact_taps(lh, G.lv[l - 1].h, taps.data() + L.tab_y);
        }
        nl = l + 1;
    }
    if (nl == 0) return fail(PSLOT_ERR_SHAPE, "image %dx%d is smaller than 2 * edgeThreshold + 1", w, h);
    G.nlevels = nl;
    G.slot_bytes = off;
    G.slot_cand = cand;
    G.slot_kp = CV_SORT_CAP * nl;
    G.n_tiles = tiles;
    G.describe_blocks = std::min(CV_SORT_CAP / 4, (max_quota + 3) / 4);
    c.level_bytes = off;
    c.cand_total = cand;
    CvBuffers &B = c.b;
    const size_t S = (size_t)slots;
    off *= S;
    CU(cudaMalloc(&B.pyr, off)); CU(cudaMalloc(&B.mask, off)); CU(cudaMalloc(&B.score, off));
    CU(cudaMemset(B.pyr, 0, off)); CU(cudaMemset(B.mask, 0, off));
    CU(cudaMemset(B.score, 0, off)); // the untested 3-px frame of every level must read as 0 in the maxima test
    int2 *d_taps = nullptr;
    CU(cudaMalloc(&d_taps, sizeof(int2) * std::max<size_t>(taps.size(), 1)));
    if (!taps.empty()) CU(cudaMemcpy(d_taps, taps.data(), sizeof(int2) * taps.size(), cudaMemcpyHostToDevice));
    B.taps = d_taps;
    CU(cudaMalloc(&B.cand, 4 * S * cand)); CU(cudaMalloc(&B.kept, 4 * S * cand)); CU(cudaMalloc(&B.kept_resp, 4 * S * cand));
    const size_t kcap = S * (size_t)G.slot_kp;
    CU(cudaMalloc(&B.final_kp, 4 * kcap)); CU(cudaMalloc(&B.final_resp, 4 * kcap));
    CU(cudaMalloc(&B.counters, sizeof(int) * CV_CNT_TOTAL * S)); CU(cudaMalloc(&B.hist, sizeof(int) * 256 * PSLOT_MAX_LEVELS * S));
    CU(cudaMalloc(&B.out_kps, sizeof(pslot_keypoint) * kcap)); CU(cudaMalloc(&B.out_desc, 32 * kcap)); CU(cudaMalloc(&B.out_count, sizeof(int) * S));
    CU(cudaHostGetDevicePointer((void **)&B.err, c.h_err, 0));
    c.stage_view = align_up(S * (size_t)w * h + 16, 256);
    CU(cudaMalloc(&c.d_stage, 2 * c.stage_view));
    c.W = w; c.H = h; c.slots = slots;
    return PSLOT_OK;
}

} // namespace

extern "C" {

const char *pslot_last_error(void) { return t_err; }
int pslot_abi_version(void) { return PSLOT_ABI_VERSION; }

int pslot_device_count(void)
{
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    return n;
}

// Page-locked host buffers for the batch API's inputs / outputs (so callers need no CUDA headers of their own).
// mode 0: cudaHostAlloc default; 1: write-combined; 2: anonymous mapping advised to transparent huge pages, touched, then
// registered with cudaHostRegister (fewer IOMMU / page-table entries per transfer on virtualised hosts).
static std::mutex g_host_mu;
static std::map<void *, size_t> g_host_mapped; // mode-2 buffers: base -> mapped length

int pslot_host_alloc(void **ptr, size_t bytes, int mode)
{
    if (!ptr) return fail(PSLOT_ERR_ARG, "ptr is null");
    *ptr = nullptr;
    if (bytes == 0) return PSLOT_OK;
    if (mode == 2) {
        const size_t huge = (size_t)2 << 20, len = align_up(bytes, huge);
        void *p = mmap(nullptr, len + huge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) return fail(PSLOT_ERR_ARG, "mmap of %zu bytes failed", len);
        void *base = (void *)align_up((size_t)p, huge);
        madvise(base, len, MADV_HUGEPAGE);
        memset(base, 0, len); // fault the pages in before they are pinned
        cudaError_t e = cudaHostRegister(base, len, cudaHostRegisterPortable);
        if (e != cudaSuccess) {
            munmap(p, len + huge);
            return fail(PSLOT_ERR_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e));
        }
        std::lock_guard<std::mutex> lock(g_host_mu);
        g_host_mapped[base] = len;
        *ptr = base; // the unaligned head (< 2 MB) stays mapped until process exit: munmap needs the original base
        return PSLOT_OK;
    }
    CU(cudaHostAlloc(ptr, bytes, (mode == 1 ? cudaHostAllocWriteCombined : cudaHostAllocDefault) | cudaHostAllocPortable));
    return PSLOT_OK;
}

int pslot_host_free(void *ptr)
{
    if (!ptr) return PSLOT_OK;
    {
        std::lock_guard<std::mutex> lock(g_host_mu);
        auto it = g_host_mapped.find(ptr);
        if (it != g_host_mapped.end()) {
            cudaHostUnregister(ptr);
            munmap(ptr, it->second);
            g_host_mapped.erase(it);
            return PSLOT_OK;
        }
    }
    CU(cudaFreeHost(ptr));
    return PSLOT_OK;
}

int pslot_hamming256(const void *a, const void *b)
{
    uint64_t x[4], y[4];
    memcpy(x, a, 32);
    memcpy(y, b, 32);
    int d = 0;
    for (int i = 0; i < 4; ++i) d += __builtin_popcountll(x[i] ^ y[i]);
    return d;
}

int pslot_orb_create(pslot_orb **ctx, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int device)
{
    if (!ctx) return fail(PSLOT_ERR_ARG, "ctx is null");
    *ctx = nullptr;
    pslot_orb *c = new (std::nothrow) pslot_orb();
    if (!c) return fail(PSLOT_ERR_ARG, "out of host memory");
    Params p{nfeatures, scale_factor, nlevels, ini_th, min_th};
    int rc = c->E.init(p, device, nullptr);
    if (rc) { delete c; return rc; }
    *ctx = c;
    return PSLOT_OK;
}

void pslot_orb_destroy(pslot_orb *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->E.device);
    if (ctx->E.st) cudaStreamSynchronize(ctx->E.st);
    delete ctx;
}

int pslot_orb_levels(const pslot_orb *ctx) { return ctx ? ctx->E.p.nlevels : fail(PSLOT_ERR_ARG, "ctx is null"); }

int pslot_orb_tables(const pslot_orb *ctx, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2, int32_t *per_level)
{
    if (!ctx) return fail(PSLOT_ERR_ARG, "ctx is null");
    const Tables &t = ctx->E.t;
    for (int i = 0; i < ctx->E.p.nlevels; ++i) {
        if (scale) scale[i] = t.sf[i];
        if (inv_scale) inv_scale[i] = t.inv_sf[i];
        if (sigma2) sigma2[i] = t.sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = t.inv_sigma2[i];
        if (per_level) per_level[i] = t.per_level[i];
    }
    return PSLOT_OK;
}

static int max_keypoints(const Engine &E)
{
    // sum over levels of max(quota+3, 4*nIni)+1; nIni depends on the image aspect, bound it generously
    int k = 0;
    for (int l = 0; l < E.p.nlevels; ++l) k += std::max(E.t.per_level[l] + 3, 64) + 1;
    return k;
}

int pslot_orb_max_keypoints(const pslot_orb *ctx)
{
    if (!ctx) return fail(PSLOT_ERR_ARG, "ctx is null");
    return ctx->E.W ? ctx->E.g.kp_cap_total : max_keypoints(ctx->E);
}

// first half of an extraction: upload + every kernel + the result download are enqueued on the context's stream
static int extract_enqueue(pslot_orb *ctx, const uint8_t *img, int w, int h, ptrdiff_t pitch, int channels, int rgb)
{
    if (!ctx) return fail(PSLOT_ERR_ARG, "ctx is null");
    ctx->extract_pending = false;
    ctx->extract_token = 0;
    if (!img || w <= 0 || h <= 0) return PSLOT_OK; // ORBextractor.cc:1046-1047: no keypoints
    if (channels != 1 && channels != 3 && channels != 4) return fail(PSLOT_ERR_ARG, "channels must be 1, 3 or 4");
    if (pitch < (ptrdiff_t)w * channels) return fail(PSLOT_ERR_ARG, "pitch < width * channels");
    Engine &E = ctx->E;
    CU(cudaSetDevice(E.device));
    int rc = E.configure(w, h, 1);
    if (rc) return rc;
    if (ctx->extract_graph && (ctx->h_stage_bytes < E.out_bytes || ctx->extract_graph_gen != E.generation)) {
        cudaGraphExecDestroy(ctx->extract_graph); // it names buffers that are about to move, or already have
        ctx->extract_graph = nullptr;
    }
    if (ctx->h_stage_bytes < E.out_bytes) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr;
        CU(cudaHostAlloc((void **)&ctx->h_stage, E.out_bytes, cudaHostAllocDefault));
        ctx->h_stage_bytes = E.out_bytes;
    }
    if (channels == 1 && pitch != w) {
        CU(cudaMemcpy2DAsync(E.pyr.lv[0], E.g.lv[0].pitch, img, pitch, w, h, cudaMemcpyHostToDevice, E.st));
    } else {
        // tightly packed gray frame: one linear copy + the re-pitch kernel (a 2-D copy of 1242-byte rows runs far below the
        // link rate); colour frame: upload once, convert on the device straight into level 0 (Tracking.cc:1014-1039)
        const size_t need = (size_t)w * h * channels + 16;
        if (ctx->cstage_bytes < need) {
            CU(cudaStreamSynchronize(E.st));
            cudaFree(ctx->d_cstage);
            ctx->d_cstage = nullptr; ctx->cstage_bytes = 0;
            CU(cudaMalloc(&ctx->d_cstage, need));
            ctx->cstage_bytes = need;
        }
        if (pitch == (ptrdiff_t)w * channels)
            CU(cudaMemcpyAsync(ctx->d_cstage, img, (size_t)w * h * channels, cudaMemcpyHostToDevice, E.st));
        else
            CU(cudaMemcpy2DAsync(ctx->d_cstage, (size_t)w * channels, img, pitch, (size_t)w * channels, h, cudaMemcpyHostToDevice, E.st));
        launch_gray_ingest(ctx->d_cstage, (size_t)w * h * channels, channels, rgb, E.pyr.lv[0], w, h, E.g.lv[0].pitch,
                           E.pyr.slot_stride[0], 1, E.st);
    }
    if (ctx->extract_calls_gen != E.generation) { ctx->extract_calls_gen = E.generation; ctx->extract_calls = 0; }
    if (!ctx->extract_graph && ctx->extract_calls >= 1) {
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(E.st, cudaStreamCaptureModeThreadLocal));
        rc = E.run_extract(1);
        cudaError_t ce = cudaMemcpyAsync(ctx->h_stage, E.d_out, E.out_bytes, cudaMemcpyDeviceToHost, E.st);
        const cudaError_t ee = cudaStreamEndCapture(E.st, &graph);
        if (ce == cudaSuccess) ce = ee;
        if (rc >= 0 && ce == cudaSuccess) ce = cudaGraphInstantiate(&ctx->extract_graph, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc < 0) return rc;
        if (ce != cudaSuccess) { ctx->extract_graph = nullptr; return fail(PSLOT_ERR_CUDA, "extract graph: %s", cudaGetErrorString(ce)); }
        ctx->extract_graph_gen = E.generation;
    }
    if (ctx->extract_graph) {
        CU(cudaGraphLaunch(ctx->extract_graph, E.st));
    } else {
        rc = E.run_extract(1);
        if (rc < 0) return rc;
        CU(cudaMemcpyAsync(ctx->h_stage, E.d_out, E.out_bytes, cudaMemcpyDeviceToHost, E.st));
    }
    ++ctx->extract_calls;
    ctx->extract_pending = true;
    static std::atomic<uint64_t> extract_generation{0}; // process-wide: tokens of different contexts never collide
    ctx->extract_token = ++extract_generation;
    return PSLOT_OK;
}

// second half: wait for the stream, hand the results over
static int extract_finish(pslot_orb *ctx, pslot_keypoint *kps, uint8_t *desc32, int cap, int *n)
{
    if (!ctx || !n) return fail(PSLOT_ERR_ARG, "ctx / n is null");
    *n = 0;
    if (!ctx->extract_pending) return PSLOT_OK; // empty image
    Engine &E = ctx->E;
    CU(cudaSetDevice(E.device));
    int rc = check_err_flag(E); // synchronises the stream
    if (rc) { ctx->extract_pending = false; return rc; }
    const int count = *(const int *)ctx->h_stage;
    *n = count;
    // (the results stay in the staging buffer: the caller may come back with more room)
    if (count > cap) return fail(PSLOT_ERR_CAPACITY, "need room for %d keypoints, caller gave %d", count, cap);
    ctx->extract_pending = false;
    if (count > 0) {
        if (!kps || !desc32) return fail(PSLOT_ERR_ARG, "kps / desc32 is null");
        memcpy(kps, ctx->h_stage + E.off_kps, sizeof(pslot_keypoint) * (size_t)count);
        memcpy(desc32, ctx->h_stage + E.off_desc, 32 * (size_t)count);
    }
    return PSLOT_OK;
}

static int extract_impl(pslot_orb *ctx, const uint8_t *img, int w, int h, ptrdiff_t pitch, int channels, int rgb,
                        pslot_keypoint *kps, uint8_t *desc32, int cap, int *n)
{
    if (!ctx || !n) return fail(PSLOT_ERR_ARG, "ctx / n is null");
    *n = 0;
    if (int rc = extract_enqueue(ctx, img, w, h, pitch, channels, rgb)) return rc;
    return extract_finish(ctx, kps, desc32, cap, n);
}

int pslot_orb_extract(pslot_orb *ctx, const uint8_t *img, int w, int h, ptrdiff_t pitch, pslot_keypoint *kps,
                      uint8_t *desc32, int cap, int *n)
{
    return extract_impl(ctx, img, w, h, pitch, 1, 0, kps, desc32, cap, n);
}

int pslot_orb_extract_color(pslot_orb *ctx, const uint8_t *img, int w, int h, ptrdiff_t pitch, int channels,
                            int rgb_order, pslot_keypoint *kps, uint8_t *desc32, int cap, int *n)
{
    return extract_impl(ctx, img, w, h, pitch, channels, rgb_order, kps, desc32, cap, n);
}

int pslot_orb_extract_begin(pslot_orb *ctx, const uint8_t *img, int w, int h, ptrdiff_t pitch)
{
    return extract_enqueue(ctx, img, w, h, pitch, 1, 0);
}

int pslot_orb_extract_end(pslot_orb *ctx, pslot_keypoint *kps, uint8_t *desc32, int cap, int *n)
{
    return extract_finish(ctx, kps, desc32, cap, n);
}

int pslot_orb_level(pslot_orb *ctx, int level, uint8_t *dst, ptrdiff_t pitch, int *w, int *h)
{
    if (!ctx) return fail(PSLOT_ERR_ARG, "ctx is null");
    Engine &E = ctx->E;
    if (!E.has_pyramid) return fail(PSLOT_ERR_STATE, "no image has been extracted yet");
    if (level < 0 || level >= E.p.nlevels) return fail(PSLOT_ERR_ARG, "level out of range");
    const LevelGeom &L = E.g.lv[level];
    if (w) *w = L.w;
    if (h) *h = L.h;
    if (!dst) return PSLOT_OK;
    if (pitch < L.w) return fail(PSLOT_ERR_ARG, "pitch < level width");
    CU(cudaSetDevice(E.device));
    CU(cudaMemcpy2DAsync(dst, pitch, E.pyr.lv[level], L.pitch, L.w, L.h, cudaMemcpyDeviceToHost, E.st));
    CU(cudaStreamSynchronize(E.st));
    return PSLOT_OK;
}

int pslot_stereo_match(pslot_orb *cl, pslot_orb *cr, const pslot_keypoint *kps_l, int n_l, const uint8_t *desc_l,
                       const pslot_keypoint *kps_r, int n_r, const uint8_t *desc_r, float bf, float max_d,
                       float *u_right, float *depth)
{
    if (!cl || !cr) return fail(PSLOT_ERR_ARG, "ctx is null");
    if (n_l < 0 || n_r < 0) return fail(PSLOT_ERR_ARG, "negative count");
    if (n_l == 0) return PSLOT_OK;
    if (!kps_l || !desc_l || !u_right || !depth || (n_r > 0 && (!kps_r || !desc_r))) return fail(PSLOT_ERR_ARG, "null array");
    Engine &L = cl->E, &R = cr->E;
    if (!L.has_pyramid || !R.has_pyramid) return fail(PSLOT_ERR_STATE, "both extractors must have processed an image");
    if (L.W != R.W || L.H != R.H || L.p.nlevels != R.p.nlevels || L.device != R.device)
        return fail(PSLOT_ERR_STATE, "left / right extractors differ in geometry or device");
    if (n_r > 16000 || n_l > (1 << 24)) return fail(PSLOT_ERR_ARG, "too many keypoints");
    for (int i = 0; i < n_l; ++i) {
        if (kps_l[i].octave < 0 || kps_l[i].octave >= L.p.nlevels) return fail(PSLOT_ERR_ARG, "left keypoint %d: octave out of range", i);
        if (!(fabsf(kps_l[i].x) < 1e6f) || !(fabsf(kps_l[i].y) < 1e6f)) return fail(PSLOT_ERR_ARG, "left keypoint %d: coordinate is not finite", i);
    }
    for (int i = 0; i < n_r; ++i) {
        if (kps_r[i].octave < 0 || kps_r[i].octave >= L.p.nlevels) return fail(PSLOT_ERR_ARG, "right keypoint %d: octave out of range", i);
        if (!(fabsf(kps_r[i].x) < 1e6f) || !(kps_r[i].y > -30000.f && kps_r[i].y < 30000.f))
            return fail(PSLOT_ERR_ARG, "right keypoint %d: coordinate is not finite / out of range", i);
    }
    CU(cudaSetDevice(L.device));
    const int need = std::max(std::max(n_l, n_r), 1);
    if (need > cl->st_cap) {
        cl->free_stereo();
        const size_t c = (size_t)need + 256;
        CU(cudaMalloc(&cl->d_kl, sizeof(pslot_keypoint) * c)); CU(cudaMalloc(&cl->d_kr, sizeof(pslot_keypoint) * c));
        CU(cudaMalloc(&cl->d_dl, 32 * c)); CU(cudaMalloc(&cl->d_dr, 32 * c));
        CU(cudaMalloc(&cl->d_n, 8));
        CU(cudaMalloc(&cl->d_ur, 4 * c)); CU(cudaMalloc(&cl->d_dp, 4 * c));
        CU(cudaMalloc(&cl->d_bd, 4 * c)); CU(cudaMalloc(&cl->d_bi, 4 * c)); CU(cudaMalloc(&cl->d_sd, 4 * c));
        cl->st_cap = (int)c;
    }
    cudaStream_t st = L.st;
    CU(cudaStreamSynchronize(R.st)); // the right pyramid must be complete (extract is synchronous anyway)
    // The four feature arrays and the counts gather in one page-locked block and travel as ONE copy (five copies from
    // pageable memory block the host once each), and the two result arrays come back the same way.
    const size_t cap = (size_t)cl->st_cap;
    const size_t o_kl = 256, o_kr = align_up(o_kl + sizeof(pslot_keypoint) * cap, 256), o_dl = align_up(o_kr + sizeof(pslot_keypoint) * cap, 256),
                 o_dr = align_up(o_dl + 32 * cap, 256), in_bytes = align_up(o_dr + 32 * cap, 256), out_bytes = 8 * cap;
    if (cl->h_match_cap != cap) {
        if (cl->h_match) cudaFreeHost(cl->h_match);
        cudaFree(cl->d_match);
        cl->h_match = nullptr; cl->d_match = nullptr; cl->h_match_cap = 0;
        CU(cudaHostAlloc((void **)&cl->h_match, in_bytes + out_bytes, cudaHostAllocDefault));
        CU(cudaMalloc(&cl->d_match, in_bytes));
        cl->h_match_cap = cap;
    }
    unsigned char *hm = cl->h_match, *dm = cl->d_match;
    const int counts[2] = {n_l, n_r};
    memcpy(hm, counts, 8);
    memcpy(hm + o_kl, kps_l, sizeof(pslot_keypoint) * (size_t)n_l);
    memcpy(hm + o_dl, desc_l, 32 * (size_t)n_l);
    if (n_r) {
        memcpy(hm + o_kr, kps_r, sizeof(pslot_keypoint) * (size_t)n_r);
        memcpy(hm + o_dr, desc_r, 32 * (size_t)n_r);
    }
    // (one copy up to the end of what is used: the right descriptors are the last block)
    CU(cudaMemcpyAsync(dm, hm, n_r ? o_dr + 32 * (size_t)n_r : o_dl + 32 * (size_t)n_l, cudaMemcpyHostToDevice, st));
    StereoArgs a{};
    a.kps_l = reinterpret_cast<const pslot_keypoint *>(dm + o_kl); a.kps_r = reinterpret_cast<const pslot_keypoint *>(dm + o_kr);
    a.desc_l = dm + o_dl; a.desc_r = dm + o_dr;
    a.n_l = reinterpret_cast<const int *>(dm); a.n_r = reinterpret_cast<const int *>(dm) + 1;
    a.K = need; // one pair: the row stride is never used, and the grid (sized from K) covers exactly the left set
    a.pyr_l = L.pyr; a.pyr_r = R.pyr;
    a.bf = bf; a.max_d = max_d;
    a.u_right = cl->d_ur; a.depth = cl->d_dp; a.best_dist = cl->d_bd; a.best_idx = cl->d_bi; a.sad = cl->d_sd;
    a.err_flag = L.d_err;
    launch_stereo(L.g, a, 1, std::max(n_r, 1), st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PSLOT_ERR_CUDA, "stereo launch failed: %s", cudaGetErrorString(e));
    float *ho = reinterpret_cast<float *>(hm + in_bytes);
    CU(cudaMemcpyAsync(ho, cl->d_ur, 4 * (size_t)n_l, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ho + cap, cl->d_dp, 4 * (size_t)n_l, cudaMemcpyDeviceToHost, st));
    if (int rc = check_err_flag(L)) return rc; // synchronises the stream
    memcpy(u_right, ho, 4 * (size_t)n_l);
    memcpy(depth, ho + cap, 4 * (size_t)n_l);
    cl->last_nl = n_l;
    return PSLOT_OK;
}

uint64_t pslot_orb_extract_token(const pslot_orb *ctx) { return ctx ? ctx->extract_token : 0; }

int pslot_stereo_match_resident(pslot_orb *cl, pslot_orb *cr, uint64_t token_l, uint64_t token_r, float bf, float max_d,
                                float *u_right, float *depth, int cap, int *n_l)
{
    if (!cl || !cr || !n_l) return fail(PSLOT_ERR_ARG, "null argument");
    *n_l = 0;
    if (!token_l || !token_r || token_l != cl->extract_token || token_r != cr->extract_token || cl->extract_pending || cr->extract_pending)
        return fail(PSLOT_ERR_STATE, "the feature sets on the device are not the ones the caller means (a later extraction ran, or one is pending)");
    Engine &L = cl->E, &R = cr->E;
    if (!L.has_pyramid || !R.has_pyramid) return fail(PSLOT_ERR_STATE, "both extractors must have processed an image");
    if (L.W != R.W || L.H != R.H || L.p.nlevels != R.p.nlevels || L.device != R.device)
        return fail(PSLOT_ERR_STATE, "left / right extractors differ in geometry or device");
    if (R.g.kp_cap_total > 16000) return fail(PSLOT_ERR_CAPACITY, "too many right keypoint slots for the stereo matcher");
    CU(cudaSetDevice(L.device));
    const int K = L.g.kp_cap_total;
    if (K > cl->st_cap) { // result arrays (the feature arrays are the extractors' own)
        cl->free_stereo();
        const size_t c = (size_t)K + 256;
        CU(cudaMalloc(&cl->d_kl, sizeof(pslot_keypoint) * c)); CU(cudaMalloc(&cl->d_kr, sizeof(pslot_keypoint) * c));
        CU(cudaMalloc(&cl->d_dl, 32 * c)); CU(cudaMalloc(&cl->d_dr, 32 * c));
        CU(cudaMalloc(&cl->d_n, 8));
        CU(cudaMalloc(&cl->d_ur, 4 * c)); CU(cudaMalloc(&cl->d_dp, 4 * c));
        CU(cudaMalloc(&cl->d_bd, 4 * c)); CU(cudaMalloc(&cl->d_bi, 4 * c)); CU(cudaMalloc(&cl->d_sd, 4 * c));
        cl->st_cap = (int)c;
    }
    if ((size_t)K > cl->h_stereo_cap) {
        if (cl->h_stereo) cudaFreeHost(cl->h_stereo);
        cl->h_stereo = nullptr; cl->h_stereo_cap = 0;
        CU(cudaHostAlloc((void **)&cl->h_stereo, sizeof(float) * 2 * (size_t)K + 16, cudaHostAllocDefault));
        cl->h_stereo_cap = (size_t)K;
    }
    cudaStream_t st = L.st;
    StereoArgs a{};
    a.kps_l = L.d_out_kps; a.kps_r = R.d_out_kps; a.desc_l = L.d_out_desc; a.desc_r = R.d_out_desc;
    a.n_l = L.d_out_count; a.n_r = R.d_out_count; // one slot each: the counts of the two extractions
    a.K = K;
    a.pyr_l = L.pyr; a.pyr_r = R.pyr;
    a.bf = bf; a.max_d = max_d;
    a.u_right = cl->d_ur; a.depth = cl->d_dp; a.best_dist = cl->d_bd; a.best_idx = cl->d_bi; a.sad = cl->d_sd;
    a.err_flag = L.d_err;
    launch_stereo(L.g, a, 1, R.g.kp_cap_total, st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PSLOT_ERR_CUDA, "stereo launch failed: %s", cudaGetErrorString(e));
    int *hn = reinterpret_cast<int *>(cl->h_stereo + 2 * (size_t)K);
    CU(cudaMemcpyAsync(hn, L.d_out_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(cl->h_stereo, cl->d_ur, 4 * (size_t)K, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(cl->h_stereo + K, cl->d_dp, 4 * (size_t)K, cudaMemcpyDeviceToHost, st));
    if (int rc = check_err_flag(L)) return rc; // synchronises the stream
    const int n = *hn;
    *n_l = n;
    cl->last_nl = n;
    if (n > cap) return fail(PSLOT_ERR_CAPACITY, "need room for %d results, caller gave %d", n, cap);
    if (n > 0) {
        if (!u_right || !depth) return fail(PSLOT_ERR_ARG, "u_right / depth is null");
        memcpy(u_right, cl->h_stereo, 4 * (size_t)n);
        memcpy(depth, cl->h_stereo + K, 4 * (size_t)n);
    }
    return PSLOT_OK;
}

int pslot_stereo_debug(pslot_orb *cl, int n_l, int32_t *best_dist, int32_t *best_idx, int32_t *best_sad)
{
    if (!cl) return fail(PSLOT_ERR_ARG, "ctx is null");
    if (n_l < 0) return fail(PSLOT_ERR_ARG, "negative count");
    if (n_l > cl->last_nl) return fail(PSLOT_ERR_STATE, "no stereo result of that size");
    CU(cudaSetDevice(cl->E.device));
    if (best_dist) CU(cudaMemcpy(best_dist, cl->d_bd, 4 * (size_t)n_l, cudaMemcpyDeviceToHost));
    if (best_idx) CU(cudaMemcpy(best_idx, cl->d_bi, 4 * (size_t)n_l, cudaMemcpyDeviceToHost));
    if (best_sad) CU(cudaMemcpy(best_sad, cl->d_sd, 4 * (size_t)n_l, cudaMemcpyDeviceToHost));
    return PSLOT_OK;
}

int pslot_orb_debug_candidates(pslot_orb *ctx, int level, int cap, int32_t *x, int32_t *y, int32_t *resp, int *n)
{
    if (!ctx || !n) return fail(PSLOT_ERR_ARG, "null argument");
    Engine &E = ctx->E;
    if (!E.has_pyramid) return fail(PSLOT_ERR_STATE, "no image has been extracted yet");
    if (level < 0 || level >= E.p.nlevels) return fail(PSLOT_ERR_ARG, "level out of range");
    const LevelGeom &L = E.g.lv[level];
    const int ncells = L.n_cols * L.n_rows;
    std::vector<int> cnt(ncells);
    std::vector<uint32_t> cand((size_t)ncells * L.cand_stride);
    CU(cudaSetDevice(E.device));
    CU(cudaMemcpy(cnt.data(), E.d_cell_count + L.cell_base, 4 * (size_t)ncells, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(cand.data(), E.d_cell_cand + L.cand_base, 4 * cand.size(), cudaMemcpyDeviceToHost));
    int k = 0;
    for (int c = 0; c < ncells; ++c)
        for (int i = 0; i < cnt[c]; ++i, ++k)
            if (k < cap) {
                const uint32_t p = cand[(size_t)c * L.cand_stride + i];
                x[k] = pslot_px(p); y[k] = pslot_py(p); resp[k] = pslot_pr(p);
            }
    *n = k;
    return k > cap ? fail(PSLOT_ERR_CAPACITY, "need room for %d candidates", k) : PSLOT_OK;
}

int pslot_orb_debug_level_keypoints(pslot_orb *ctx, int level, int cap, int32_t *x, int32_t *y, int32_t *resp,
                                    float *angle, int *n)
{
    if (!ctx || !n) return fail(PSLOT_ERR_ARG, "null argument");
    Engine &E = ctx->E;
    if (!E.has_pyramid) return fail(PSLOT_ERR_STATE, "no image has been extracted yet");
    if (level < 0 || level >= E.p.nlevels) return fail(PSLOT_ERR_ARG, "level out of range");
    const LevelGeom &L = E.g.lv[level];
    std::vector<int> cnt(E.p.nlevels);
    std::vector<uint32_t> kp(L.kp_cap);
    CU(cudaSetDevice(E.device));
    CU(cudaMemcpy(cnt.data(), E.d_lkp_count, 4 * (size_t)E.p.nlevels, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(kp.data(), E.d_lkp + L.kp_base, 4 * (size_t)L.kp_cap, cudaMemcpyDeviceToHost));
    int base = 0;
    for (int l = 0; l < level; ++l) base += cnt[l];
    const int c = cnt[level];
    std::vector<pslot_keypoint> out(std::max(c, 1));
    if (c) CU(cudaMemcpy(out.data(), E.d_out_kps + base, sizeof(pslot_keypoint) * (size_t)c, cudaMemcpyDeviceToHost));
    *n = c;
    if (c > cap) return fail(PSLOT_ERR_CAPACITY, "need room for %d keypoints", c);
    for (int i = 0; i < c; ++i) {
        x[i] = pslot_px(kp[i]); y[i] = pslot_py(kp[i]); resp[i] = pslot_pr(kp[i]);
        if (angle) angle[i] = out[i].angle;
    }
    return PSLOT_OK;
}

int pslot_debug_steer(int device, const float *angle_deg, int n, float *a, float *b)
{
    if (!angle_deg || !a || !b || n < 0) return fail(PSLOT_ERR_ARG, "bad argument");
    if (n == 0) return PSLOT_OK;
    CU(cudaSetDevice(device));
    float *d = nullptr;
    CU(cudaMalloc(&d, 12 * (size_t)n));
    cudaError_t e = cudaMemcpy(d, angle_deg, 4 * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        launch_debug_steer(d, n, d + n, d + 2 * (size_t)n, 0);
        e = cudaMemcpy(a, d + n, 4 * (size_t)n, cudaMemcpyDeviceToHost);
    }
    if (e == cudaSuccess) e = cudaMemcpy(b, d + 2 * (size_t)n, 4 * (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(PSLOT_ERR_CUDA, "pslot_debug_steer: %s", cudaGetErrorString(e));
    return PSLOT_OK;
}


// ---- cv::ORB ----------------------------------------------------------------------------------------------------------
int pslot_cvorb_create(pslot_cvorb **ctx, int nfeatures, float scale_factor, int nlevels, int edge_threshold,
                       int fast_threshold, int device)
{
    if (!ctx) return fail(PSLOT_ERR_ARG, "ctx is null");
    *ctx = nullptr;
    if (nfeatures < 1 || nfeatures > 30000) return fail(PSLOT_ERR_ARG, "nfeatures must be in [1,30000]");
    if (nlevels < 1 || nlevels > PSLOT_MAX_LEVELS) return fail(PSLOT_ERR_ARG, "nlevels must be in [1,%d]", PSLOT_MAX_LEVELS);
    if (!(scale_factor > 1.0f)) return fail(PSLOT_ERR_ARG, "scaleFactor must be > 1");
    if (edge_threshold < 19) return fail(PSLOT_ERR_ARG, "edgeThreshold must be >= 19 (patch 31: 15 px pattern + 3 px blur inside the 21 px staged)");
    if (fast_threshold < 1 || fast_threshold > 254) return fail(PSLOT_ERR_ARG, "fastThreshold must be in [1,254]");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PSLOT_ERR_CUDA, "CUDA device %d not available (%d visible)", device, ndev);
    pslot_cvorb *c = new (std::nothrow) pslot_cvorb();
    if (!c) return fail(PSLOT_ERR_ARG, "out of host memory");
    c->nfeatures = nfeatures; c->nlevels = nlevels; c->edge = edge_threshold; c->fast_th = fast_threshold; c->device = device;
    c->scale_factor = (double)scale_factor; // ORB::create takes a float, the member is a double
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->st_up, cudaStreamNonBlocking);
    for (cudaEvent_t &ev : c->ev_up)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&c->h_err, sizeof(int), cudaHostAllocMapped);
    if (e != cudaSuccess) { delete c; return fail(PSLOT_ERR_CUDA, "pslot_cvorb_create: %s", cudaGetErrorString(e)); }
    *c->h_err = 0;
    *ctx = c;
    return PSLOT_OK;
}

void pslot_cvorb_destroy(pslot_cvorb *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->st_up) cudaStreamSynchronize(ctx->st_up);
    if (ctx->st) cudaStreamSynchronize(ctx->st);
    delete ctx;
}

// The images come from the host (imgs) or - dev_src != NULL - are already on the device as padded level-0 images with the
// extractor's row pitch (align_up(w, 128), the same as here), dev_stride bytes apart.
static int cvorb_batch_impl(pslot_cvorb *ctx, int n_images, const uint8_t *const *imgs, const uint8_t *dev_src, size_t dev_stride,
                            int w, int h, ptrdiff_t pitch, const uint8_t *const *masks, ptrdiff_t mask_pitch, pslot_keypoint *kps,
                            uint8_t *desc32, int cap, int *counts)
{
    if (!ctx || !counts) return fail(PSLOT_ERR_ARG, "ctx / counts is null");
    if (n_images < 0 || n_images > 1024) return fail(PSLOT_ERR_ARG, "n_images must be in [0,1024]");
    if (cap < 0) return fail(PSLOT_ERR_ARG, "cap < 0");
    for (int s = 0; s < n_images; ++s) counts[s] = 0;
    if (n_images == 0 || w <= 0 || h <= 0) return PSLOT_OK; // empty image: no keypoints (orb.cpp: CV_Assert only on type)
    const bool img_dev = dev_src != nullptr;
    ctx->res_batch = nullptr; // set again by the resident entry point when this call succeeds
    if (!img_dev && !imgs) return fail(PSLOT_ERR_ARG, "imgs is null");
    bool any_mask = false;
    for (int s = 0; s < n_images; ++s) {
        if (!img_dev && !imgs[s]) return fail(PSLOT_ERR_ARG, "image %d is null", s);
        any_mask |= masks && masks[s];
    }
    if ((!img_dev && pitch < w) || (any_mask && mask_pitch < w)) return fail(PSLOT_ERR_ARG, "pitch < width");
    pslot_cvorb &C = *ctx;
    CU(cudaSetDevice(C.device));
    if (w < 2 * C.edge + 1 || h < 2 * C.edge + 1) return PSLOT_OK; // nothing passes runByImageBorder: cv::ORB returns no keypoints
    if (int rc = cvorb_configure(C, w, h, n_images)) return rc;
    const CvLevel &L0 = C.g.lv[0];
    const auto t0 = std::chrono::steady_clock::now();
    // A larger batch travels in up to four chunks: chunk k + 1 is uploaded (second stream) while chunk k is processed.
    // Page-locked images arrive at link speed (0.55 ms per 64 KITTI frames): two chunks hide half of that at little cost in
    // smaller launches; pageable images are staged by the driver at a fifth of the rate: four chunks.
    int nchunks = 1;
    if (img_dev) {
        nchunks = any_mask && n_images >= 32 ? 2 : 1; // only the masks travel
    } else if (n_images >= 16) {
        cudaPointerAttributes pa{};
        const bool pinned = cudaPointerGetAttributes(&pa, imgs[0]) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError(); // an unregistered pointer is not an error here
        nchunks = pinned ? (n_images >= 32 ? 2 : 1) : 4;
    }
    const int chunk = (n_images + nchunks - 1) / nchunks;
    const bool tight = (img_dev || pitch == w) && (!any_mask || mask_pitch == w);
    const size_t img_bytes = (size_t)w * h;
    // one round trip: the counts and the (almost always sufficient) first `guess` records of every image travel together
    // into pinned memory
    const int guess = std::min(C.nfeatures + 64, C.g.slot_kp);
    const size_t S = (size_t)n_images, hdr = align_up(sizeof(int) * S, 16);
    const size_t need = hdr + S * guess * (sizeof(pslot_keypoint) + 32);
    if (C.h_out_bytes < need) {
        C.drop_graphs();
        if (C.h_out) cudaFreeHost(C.h_out);
        C.h_out = nullptr; C.h_out_bytes = 0;
        CU(cudaHostAlloc((void **)&C.h_out, need, cudaHostAllocDefault));
        C.h_out_bytes = need;
    }
    pslot_keypoint *hk = reinterpret_cast<pslot_keypoint *>(C.h_out + hdr);
    uint8_t *hd = C.h_out + hdr + S * guess * sizeof(pslot_keypoint);
    const int *hc = reinterpret_cast<const int *>(C.h_out);
    auto device_work = [&](int c0, int c1) { // re-pitch + the kernels of slots [c0, c1) on the main stream
        if (tight)
            for (int view = img_dev ? 1 : 0; view < (any_mask ? 2 : 1); ++view)
                launch_gray_ingest(C.d_stage + view * C.stage_view + c0 * img_bytes, img_bytes, 1, 0,
                                   (view ? C.b.mask : C.b.pyr) + c0 * (size_t)C.g.slot_bytes + L0.off, w, h, L0.pitch, C.g.slot_bytes,
                                   c1 - c0, C.st);
        launch_cvorb(C.g, cv_slot_view(C.g, C.b, c0), c1 - c0, any_mask, C.st);
    };
    auto downloads = [&]() -> int {
        CU(cudaMemcpyAsync(C.h_out, C.b.out_count, sizeof(int) * S, cudaMemcpyDeviceToHost, C.st));
        CU(cudaMemcpy2DAsync(hk, sizeof(pslot_keypoint) * (size_t)guess, C.b.out_kps, sizeof(pslot_keypoint) * (size_t)C.g.slot_kp,
                             sizeof(pslot_keypoint) * (size_t)guess, S, cudaMemcpyDeviceToHost, C.st));
        CU(cudaMemcpy2DAsync(hd, 32 * (size_t)guess, C.b.out_desc, 32 * (size_t)C.g.slot_kp, 32 * (size_t)guess, S,
                             cudaMemcpyDeviceToHost, C.st));
        return PSLOT_OK;
    };
    double t_up = 0;
    for (int c0 = 0, ci = 0; c0 < n_images; c0 += chunk, ++ci) {
        const int c1 = std::min(c0 + chunk, n_images);
        const auto u0 = std::chrono::steady_clock::now();
        if (img_dev) // level-0 images as they lie in the extractor's pyramid: one device-to-device copy, image = "row"
            CU(cudaMemcpy2DAsync(C.b.pyr + c0 * (size_t)C.g.slot_bytes + L0.off, (size_t)C.g.slot_bytes, dev_src + c0 * dev_stride, dev_stride,
                                 (size_t)L0.pitch * h, (size_t)(c1 - c0), cudaMemcpyDeviceToDevice, C.st_up));
        for (int view = img_dev ? 1 : 0; view < (any_mask ? 2 : 1); ++view) {
            const uint8_t *const *src = view ? masks : imgs;
            if (tight) {
                // tightly packed rows: linear copies at the full PCIe rate into a staging buffer (images that follow each other
                // in host memory travel as one copy); a re-pitch launch then fills the 128-byte aligned level-0 rows (a 2-D
                // copy of 1242-byte rows runs far below the link rate)
                uint8_t *stg = C.d_stage + view * C.stage_view;
                for (int s = c0; s < c1;) {
                    if (!src[s]) { // no mask for this image of a masked batch: keep everything
                        CU(cudaMemsetAsync(stg + s * img_bytes, 255, img_bytes, C.st_up));
                        ++s;
                        continue;
                    }
                    int e = s + 1;
                    while (e < c1 && src[e] == src[e - 1] + img_bytes) ++e;
                    CU(cudaMemcpyAsync(stg + s * img_bytes, src[s], img_bytes * (size_t)(e - s), cudaMemcpyHostToDevice, C.st_up));
                    s = e;
                }
            } else {
                uint8_t *dst = (view ? C.b.mask : C.b.pyr) + L0.off;
                const ptrdiff_t sp = view ? mask_pitch : pitch;
                for (int s = c0; s < c1; ++s) {
                    uint8_t *d = dst + s * (size_t)C.g.slot_bytes;
                    if (src[s]) CU(cudaMemcpy2DAsync(d, L0.pitch, src[s], sp, w, h, cudaMemcpyHostToDevice, C.st_up));
                    else CU(cudaMemset2DAsync(d, L0.pitch, 255, w, h, C.st_up));
                }
            }
        }
        CU(cudaEventRecord(C.ev_up[ci], C.st_up));
        CU(cudaStreamWaitEvent(C.st, C.ev_up[ci], 0));
        t_up += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - u0).count();
        if (nchunks > 1) device_work(c0, c1);
    }
    const auto t1 = t0 + std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::duration<double, std::milli>(t_up));
    if (nchunks > 1) {
        if (int rc = downloads()) return rc;
    } else {
        const int key = n_images * 8 + (img_dev ? 4 : 0) + (any_mask ? 2 : 0) + (tight ? 1 : 0);
        auto it = C.graphs.find(key);
        if (it == C.graphs.end()) {
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            CU(cudaStreamBeginCapture(C.st, cudaStreamCaptureModeThreadLocal));
            device_work(0, n_images);
            const int rc = downloads();
            cudaError_t ce = cudaStreamEndCapture(C.st, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce == cudaSuccess) ce = cudaGraphInstantiate(&exec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return fail(PSLOT_ERR_CUDA, "cv::ORB graph: %s", cudaGetErrorString(ce));
            if (C.graphs.size() >= 16) C.drop_graphs(); // a caller that keeps changing the batch size: do not hoard executables
            it = C.graphs.emplace(key, exec).first;
        }
        CU(cudaGraphLaunch(it->second, C.st));
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PSLOT_ERR_CUDA, "cv::ORB launch failed: %s", cudaGetErrorString(e));
    const auto t2 = std::chrono::steady_clock::now();
    CU(cudaStreamSynchronize(C.st));
    const auto t3 = std::chrono::steady_clock::now();
    if (*(volatile int *)C.h_err) {
        *C.h_err = 0;
        return fail(PSLOT_ERR_CAPACITY, "cv::ORB: a per-level candidate list or the keypoint sort overflowed");
    }
    int worst = 0;
    for (int s = 0; s < n_images; ++s) { counts[s] = hc[s]; worst = std::max(worst, hc[s]); }
    if (worst > cap) return fail(PSLOT_ERR_CAPACITY, "need room for %d keypoints per image, caller gave %d", worst, cap);
    if (worst > 0 && (!kps || !desc32)) return fail(PSLOT_ERR_ARG, "kps / desc32 is null");
    bool refetch = false;
    for (int s = 0; s < n_images; ++s) {
        const int count = hc[s];
        if (count == 0) continue;
        pslot_keypoint *ok = kps + (size_t)s * cap;
        uint8_t *od = desc32 + (size_t)s * cap * 32;
        if (count <= guess) {
            memcpy(ok, hk + (size_t)s * guess, sizeof(pslot_keypoint) * (size_t)count);
            memcpy(od, hd + (size_t)s * guess * 32, 32 * (size_t)count);
        } else { // massive response ties kept by retainBest: fetch the full set
            CU(cudaMemcpyAsync(ok, C.b.out_kps + (size_t)s * C.g.slot_kp, sizeof(pslot_keypoint) * (size_t)count, cudaMemcpyDeviceToHost, C.st));
            CU(cudaMemcpyAsync(od, C.b.out_desc + (size_t)s * C.g.slot_kp * 32, 32 * (size_t)count, cudaMemcpyDeviceToHost, C.st));
            refetch = true;
        }
    }
    if (refetch) CU(cudaStreamSynchronize(C.st));
    const auto t4 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    C.t_phase[0] = ms(t0, t1); C.t_phase[1] = ms(t1, t2); C.t_phase[2] = ms(t2, t3); C.t_phase[3] = ms(t3, t4);
    return PSLOT_OK;
}

int pslot_cvorb_last_timing(const pslot_cvorb *ctx, double out_ms[4])
{
    if (!ctx || !out_ms) return fail(PSLOT_ERR_ARG, "null argument");
    for (int i = 0; i < 4; ++i) out_ms[i] = ctx->t_phase[i];
    return PSLOT_OK;
}

int pslot_cvorb_detect_and_compute_batch(pslot_cvorb *ctx, int n_images, const uint8_t *const *imgs, int w, int h, ptrdiff_t pitch,
                                         const uint8_t *const *masks, ptrdiff_t mask_pitch, pslot_keypoint *kps, uint8_t *desc32,
                                         int cap, int *counts)
{
    return cvorb_batch_impl(ctx, n_images, imgs, nullptr, 0, w, h, pitch, masks, mask_pitch, kps, desc32, cap, counts);
}

int pslot_cvorb_detect_and_compute_resident(pslot_cvorb *ctx, pslot_batch *b, int first_image, int n_images,
                                            const uint8_t *const *masks, ptrdiff_t mask_pitch, pslot_keypoint *kps,
                                            uint8_t *desc32, int cap, int *counts)
{
    if (!ctx || !b) return fail(PSLOT_ERR_ARG, "ctx / b is null");
    if (b->last_pairs <= 0) return fail(PSLOT_ERR_STATE, "no batch has been submitted");
    if (first_image < 0 || n_images < 0 || first_image + n_images > 2 * b->last_pairs)
        return fail(PSLOT_ERR_ARG, "images [%d, %d) are not among the %d images of the last submit", first_image, first_image + n_images,
                    2 * b->last_pairs);
    Engine &E = b->E;
    if (E.device != ctx->device) return fail(PSLOT_ERR_STATE, "the batch lives on another device");
    CU(cudaSetDevice(E.device));
    CU(cudaStreamSynchronize(E.st)); // the pyramids of the last submit are complete (a no-op after pslot_batch_wait)
    const int rc = cvorb_batch_impl(ctx, n_images, nullptr, E.pyr.lv[0] + (size_t)first_image * E.pyr.slot_stride[0],
                                    (size_t)E.pyr.slot_stride[0], E.W, E.H, E.W, masks, mask_pitch, kps, desc32, cap, counts);
    if (rc == PSLOT_OK || rc == PSLOT_ERR_CAPACITY) { // the device holds the complete result either way
        ctx->res_batch = b; ctx->res_submits = b->submits; ctx->res_first = first_image; ctx->res_n = n_images;
    }
    return rc;
}

int pslot_cvorb_stereo_match_resident(pslot_cvorb *ctx, pslot_batch *b, float bf, float max_d, float *u_right, float *depth, int cap)
{
    if (!ctx || !b || !u_right || !depth) return fail(PSLOT_ERR_ARG, "null argument");
    pslot_cvorb &C = *ctx;
    Engine &E = b->E;
    const int np = b->last_pairs;
    if (C.res_batch != b || C.res_submits != b->submits || C.res_first != 0 || C.res_n != 2 * np || np <= 0)
        return fail(PSLOT_ERR_STATE, "the context does not hold the cv::ORB features of all %d images of this batch's last submit", 2 * np);
    if (C.nlevels > E.p.nlevels) return fail(PSLOT_ERR_ARG, "cv::ORB has more levels (%d) than the batch's pyramids (%d)", C.nlevels, E.p.nlevels);
    if (cap < 1) return fail(PSLOT_ERR_ARG, "cap < 1");
    CU(cudaSetDevice(C.device));
    const int K = C.g.slot_kp;
    const int *hc = reinterpret_cast<const int *>(C.h_out); // counts of the resident call, still in the staging block
    int max_l = 0, max_r = 1;
    for (int p = 0; p < np; ++p) { max_l = std::max(max_l, hc[p]); max_r = std::max(max_r, hc[np + p]); }
    if (max_r > 16000) return fail(PSLOT_ERR_CAPACITY, "too many right keypoints for the stereo matcher (%d > 16000)", max_r);
    if (max_l > cap) return fail(PSLOT_ERR_CAPACITY, "need room for %d results per pair, caller gave %d", max_l, cap);
    const size_t plane = 4 * (size_t)np * K;
    if (5 * plane > C.stereo_bytes) {
        CU(cudaStreamSynchronize(C.st));
        cudaFree(C.d_stereo);
        C.d_stereo = nullptr; C.stereo_bytes = 0;
        CU(cudaMalloc(&C.d_stereo, 5 * plane));
        C.stereo_bytes = 5 * plane;
    }
    StereoArgs a{};
    a.kps_l = C.b.out_kps; a.kps_r = C.b.out_kps + (size_t)np * K;
    a.desc_l = C.b.out_desc; a.desc_r = C.b.out_desc + (size_t)np * K * 32;
    a.n_l = C.b.out_count; a.n_r = C.b.out_count + np;
    a.K = K;
    a.pyr_l = E.pyr; a.pyr_r = E.pyr; // pair p: slot p of the left views, slot npairs + p of the right views
    for (int l = 0; l < E.g.nlevels; ++l) a.pyr_r.lv[l] = E.pyr.lv[l] + (size_t)np * E.pyr.slot_stride[l];
    a.bf = bf; a.max_d = max_d;
    a.u_right = reinterpret_cast<float *>(C.d_stereo); a.depth = reinterpret_cast<float *>(C.d_stereo + plane);
    a.best_dist = reinterpret_cast<int *>(C.d_stereo + 2 * plane); a.best_idx = reinterpret_cast<int *>(C.d_stereo + 3 * plane);
    a.sad = reinterpret_cast<int *>(C.d_stereo + 4 * plane);
    a.err_flag = C.b.err;
    a.counters = nullptr;
    CU(cudaStreamSynchronize(E.st)); // the batch's pyramids are complete and idle
    launch_stereo(E.g, a, np, (max_r + 63) & ~63, C.st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PSLOT_ERR_CUDA, "stereo launch failed: %s", cudaGetErrorString(e));
    const size_t wbytes = 4 * (size_t)std::min(cap, K);
    CU(cudaMemcpy2DAsync(u_right, 4 * (size_t)cap, a.u_right, 4 * (size_t)K, wbytes, (size_t)np, cudaMemcpyDeviceToHost, C.st));
    CU(cudaMemcpy2DAsync(depth, 4 * (size_t)cap, a.depth, 4 * (size_t)K, wbytes, (size_t)np, cudaMemcpyDeviceToHost, C.st));
    CU(cudaStreamSynchronize(C.st));
    if (*(volatile int *)C.h_err) {
        const int f = *C.h_err;
        *C.h_err = 0;
        return fail(PSLOT_ERR_CUDA, "stereo matcher on the cv::ORB features: device error flag 0x%x", f);
    }
    return PSLOT_OK;
}

int pslot_cvorb_detect_and_compute(pslot_cvorb *ctx, const uint8_t *img, int w, int h, ptrdiff_t pitch, const uint8_t *mask,
                                   ptrdiff_t mask_pitch, pslot_keypoint *kps, uint8_t *desc32, int cap, int *n)
{
    if (!ctx || !n) return fail(PSLOT_ERR_ARG, "ctx / n is null");
    *n = 0;
    if (!img || w <= 0 || h <= 0) return PSLOT_OK; // empty image: no keypoints
    return pslot_cvorb_detect_and_compute_batch(ctx, 1, &img, w, h, pitch, mask ? &mask : nullptr, mask_pitch, kps, desc32,
                                                cap < 0 ? 0 : cap, n);
}

// ---- bag of words ------------------------------------------------------------------------------------------------------
int pslot_vocab_create(pslot_vocab **voc, int n_nodes, int k, int L, int scoring, int weighting, const uint8_t *desc32,
                       const int32_t *child_start, const int32_t *child, const int32_t *word_id, const double *weight,
                       int device)
{
    if (!voc) return fail(PSLOT_ERR_ARG, "voc is null");
    *voc = nullptr;
    if (n_nodes < 2 || !desc32 || !child_start || !child || !word_id || !weight) return fail(PSLOT_ERR_ARG, "empty vocabulary");
    if (L < 1 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3) return fail(PSLOT_ERR_ARG, "bad L / scoring / weighting");
    // the tree must be walkable: child lists inside the array, every child a later node than its parent (file order), the
    // root has children, every leaf carries a word
    if (child_start[0] != 0 || child_start[n_nodes] != n_nodes - 1 || child_start[1] <= 0)
        return fail(PSLOT_ERR_ARG, "child_start must run from 0 to n_nodes - 1 and the root must have children");
    for (int i = 0; i < n_nodes; ++i) {
        if (child_start[i + 1] < child_start[i]) return fail(PSLOT_ERR_ARG, "child_start not ascending at node %d", i);
        if (child_start[i + 1] - child_start[i] > 65535) return fail(PSLOT_ERR_ARG, "node %d has more than 65535 children", i);
        for (int c = child_start[i]; c < child_start[i + 1]; ++c)
            if (child[c] <= i || child[c] >= n_nodes) return fail(PSLOT_ERR_ARG, "child %d of node %d is not a later node", child[c], i);
        if (child_start[i + 1] == child_start[i] && word_id[i] < 0) return fail(PSLOT_ERR_ARG, "leaf %d has no word id", i);
    }
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PSLOT_ERR_CUDA, "CUDA device %d not available (%d visible)", device, ndev);
    pslot_vocab *v = new (std::nothrow) pslot_vocab();
    if (!v) return fail(PSLOT_ERR_ARG, "out of host memory");
    v->n_nodes = n_nodes; v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->device = device;
    v->word_id.assign(word_id, word_id + n_nodes);
    v->weight.assign(weight, weight + n_nodes);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&v->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&v->d_desc, 32 * (size_t)n_nodes);
    if (e == cudaSuccess) e = cudaMalloc(&v->d_child_start, 4 * (size_t)(n_nodes + 1));
    if (e == cudaSuccess) e = cudaMalloc(&v->d_child, 4 * (size_t)n_nodes);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_desc, desc32, 32 * (size_t)n_nodes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_child_start, child_start, 4 * (size_t)(n_nodes + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_child, child, 4 * (size_t)(n_nodes - 1), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete v; return fail(PSLOT_ERR_CUDA, "pslot_vocab_create: %s", cudaGetErrorString(e)); }
    *voc = v;
    return PSLOT_OK;
}

void pslot_vocab_destroy(pslot_vocab *voc)
{
    if (!voc) return;
    cudaSetDevice(voc->device);
    if (voc->st) cudaStreamSynchronize(voc->st);
    delete voc;
}

int pslot_bow_transform(pslot_vocab *voc, pslot_orb *ctx, const uint8_t *desc32, int n, int levelsup, int32_t *leaf_node,
                        int32_t *nid)
{
    if (!voc) return fail(PSLOT_ERR_ARG, "voc is null");
    if (n < 0 || (n > 0 && (!leaf_node || !nid))) return fail(PSLOT_ERR_ARG, "bad n / null outputs");
    if (n == 0) return PSLOT_OK;
    pslot_vocab &V = *voc;
    CU(cudaSetDevice(V.device));
    if (n > V.cap) {
        CU(cudaStreamSynchronize(V.st));
        cudaFree(V.d_feat); cudaFree(V.d_leaf); cudaFree(V.d_nid);
        V.d_feat = nullptr; V.d_leaf = V.d_nid = nullptr; V.cap = 0;
        const int cap = std::max(n, 4096);
        CU(cudaMalloc(&V.d_feat, 32 * (size_t)cap)); CU(cudaMalloc(&V.d_leaf, 4 * (size_t)cap)); CU(cudaMalloc(&V.d_nid, 4 * (size_t)cap));
        V.cap = cap;
    }
    const uint8_t *feat = V.d_feat;
    if (desc32) {
        CU(cudaMemcpyAsync(V.d_feat, desc32, 32 * (size_t)n, cudaMemcpyHostToDevice, V.st));
    } else {
        if (!ctx) return fail(PSLOT_ERR_ARG, "desc32 is null and no extractor context given");
        Engine &E = ctx->E;
        if (!E.has_pyramid || E.device != V.device) return fail(PSLOT_ERR_STATE, "the extractor holds no features on this device");
        CU(cudaStreamSynchronize(E.st));
        if (n > E.g.kp_cap_total) return fail(PSLOT_ERR_ARG, "n exceeds the extractor's keypoint capacity");
        feat = E.d_out_desc; // the last pslot_orb_extract left its descriptors here (slot 0)
    }
    launch_bow_transform(V.d_desc, V.d_child_start, V.d_child, feat, n, V.L - levelsup, V.d_leaf, V.d_nid, V.st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PSLOT_ERR_CUDA, "bow launch failed: %s", cudaGetErrorString(e));
    CU(cudaMemcpyAsync(leaf_node, V.d_leaf, 4 * (size_t)n, cudaMemcpyDeviceToHost, V.st));
    CU(cudaMemcpyAsync(nid, V.d_nid, 4 * (size_t)n, cudaMemcpyDeviceToHost, V.st));
    CU(cudaStreamSynchronize(V.st));
    return PSLOT_OK;
}

int pslot_bow_vectors(const pslot_vocab *voc, const int32_t *leaf_node, const int32_t *nid, int n, int32_t *bow_id,
                      double *bow_val, int *n_words, int32_t *fv_node, int32_t *fv_start, int32_t *fv_feat, int *n_fv_nodes)
{
    if (!voc || !n_words || !n_fv_nodes || n < 0) return fail(PSLOT_ERR_ARG, "null argument");
    *n_words = *n_fv_nodes = 0;
    if (n > 0 && (!leaf_node || !nid || !bow_id || !bow_val || !fv_node || !fv_start || !fv_feat)) return fail(PSLOT_ERR_ARG, "null array");
    // DBoW2's order of operations (TemplatedVocabulary.h:1132-1201, BowVector.cpp, FeatureVector.cpp) without its node-based
    // maps: the non-stopped features are ordered by (key, feature index) - a std::map visits keys ascending, and both
    // addWeight's running sum and FeatureVector's push_back see a key's features in feature order
    const bool tf = voc->weighting == 0 || voc->weighting == 1;
    std::vector<std::pair<uint32_t, int32_t>> by_word, by_node; // (key, feature)
    by_word.reserve((size_t)n); by_node.reserve((size_t)n);
    for (int i = 0; i < n; ++i) {
        const int leaf = leaf_node[i];
        if (leaf < 0 || leaf >= voc->n_nodes || voc->word_id[leaf] < 0) return fail(PSLOT_ERR_ARG, "leaf_node[%d] = %d is not a word", i, leaf);
        if (!(voc->weight[leaf] > 0)) continue; // stopped word
        by_word.emplace_back((uint32_t)voc->word_id[leaf], i);
        by_node.emplace_back((uint32_t)nid[i], i);
    }
    std::sort(by_word.begin(), by_word.end());
    std::sort(by_node.begin(), by_node.end());
    int k = 0;
    for (size_t i = 0; i < by_word.size();) {
        const uint32_t id = by_word[i].first;
        const double w = voc->weight[leaf_node[by_word[i].second]];
        double acc = w; // addWeight: v[id] = w, then += w per further feature; addIfNotExist: w once
        size_t e = i + 1;
        for (; e < by_word.size() && by_word[e].first == id; ++e)
            if (tf) acc += w;
        bow_id[k] = (int32_t)id; bow_val[k] = acc; ++k;
        i = e;
    }
    const bool must = voc->scoring != 5, l2 = voc->scoring == 1; // ScoringObject.h: DOT_PRODUCT does not normalise, L2_NORM uses L2
    if (tf && k > 0 && !must) {
        const double nd = (double)k;
        for (int i = 0; i < k; ++i) bow_val[i] /= nd;
    }
    if (must) {
        double norm = 0.0;
        if (!l2) for (int i = 0; i < k; ++i) norm += std::fabs(bow_val[i]);
        else { for (int i = 0; i < k; ++i) norm += bow_val[i] * bow_val[i]; norm = std::sqrt(norm); }
        if (norm > 0.0) for (int i = 0; i < k; ++i) bow_val[i] /= norm;
    }
    int j = 0, pos = 0;
    for (size_t i = 0; i < by_node.size(); ++i) {
        if (i == 0 || by_node[i].first != by_node[i - 1].first) { fv_node[j] = (int32_t)by_node[i].first; fv_start[j] = pos; ++j; }
        fv_feat[pos++] = by_node[i].second;
    }
    if (n > 0) fv_start[j] = pos;
    *n_words = k; *n_fv_nodes = j;
    return PSLOT_OK;
}

int pslot_cvorb_debug(pslot_cvorb *ctx, int what, int level, void *out, float *aux, size_t cap_bytes, int *w, int *h, int *n)
{
    if (!ctx || !out || !w || !h || !n) return fail(PSLOT_ERR_ARG, "null argument");
    pslot_cvorb &C = *ctx;
    if (C.W == 0) return fail(PSLOT_ERR_STATE, "no image processed yet");
    if (level < 0 || level >= C.g.nlevels) return fail(PSLOT_ERR_ARG, "level out of range");
    CU(cudaSetDevice(C.device));
    CU(cudaStreamSynchronize(C.st));
    const CvLevel &L = C.g.lv[level];
    *w = L.w; *h = L.h; *n = 0;
    if (what == 0 || what == 1) {
        if (cap_bytes < (size_t)L.w * L.h) return fail(PSLOT_ERR_CAPACITY, "need %zu bytes", (size_t)L.w * L.h);
        CU(cudaMemcpy2D(out, L.w, (what ? C.b.score : C.b.pyr) + L.off, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost));
        *n = L.w * L.h;
        return PSLOT_OK;
    }
    if (what != 2 && what != 3) return fail(PSLOT_ERR_ARG, "what must be 0..3");
    int counters[CV_CNT_TOTAL];
    CU(cudaMemcpy(counters, C.b.counters, sizeof(counters), cudaMemcpyDeviceToHost));
    const int cnt = std::min(counters[(what == 2 ? CV_CNT_CAND : CV_CNT_KEPT) + level], L.cand_cap);
    if (cap_bytes < 4 * (size_t)cnt) return fail(PSLOT_ERR_CAPACITY, "need %zu bytes", 4 * (size_t)cnt);
    CU(cudaMemcpy(out, (what == 2 ? C.b.cand : C.b.kept) + L.cand_base, 4 * (size_t)cnt, cudaMemcpyDeviceToHost));
    if (what == 3 && aux) CU(cudaMemcpy(aux, C.b.kept_resp + L.cand_base, 4 * (size_t)cnt, cudaMemcpyDeviceToHost));
    *n = cnt;
    return PSLOT_OK;
}

// ---- throughput API -------------------------------------------------------------------------------
int pslot_batch_create(pslot_batch **b, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th,
                       int width, int height, int max_pairs, int device, void *stream)
{
    if (!b) return fail(PSLOT_ERR_ARG, "b is null");
    *b = nullptr;
    if (max_pairs < 1 || max_pairs > 16384) return fail(PSLOT_ERR_ARG, "max_pairs must be in [1,16384]");
    if (width <= 0 || height <= 0) return fail(PSLOT_ERR_ARG, "bad image size");
    pslot_batch *c = new (std::nothrow) pslot_batch();
    if (!c) return fail(PSLOT_ERR_ARG, "out of host memory");
    Params p{nfeatures, scale_factor, nlevels, ini_th, min_th};
    int rc = c->E.init(p, device, stream);
    if (!rc) rc = c->E.configure(width, height, 2 * max_pairs);
    if (!rc && c->E.g.kp_cap_total > 16000) // the stereo kernel stages every right key of a pair in shared memory
        rc = fail(PSLOT_ERR_CAPACITY, "nfeatures too large for the batched stereo matcher (%d keypoint slots > 16000)", c->E.g.kp_cap_total);
    if (rc) { delete c; return rc; }
    c->max_pairs = max_pairs;
    *b = c;
    return PSLOT_OK;
}

void pslot_batch_destroy(pslot_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->E.device);
    if (b->E.st) cudaStreamSynchronize(b->E.st);
    delete b;
}

int pslot_batch_max_keypoints(const pslot_batch *b) { return b ? b->E.g.kp_cap_total : fail(PSLOT_ERR_ARG, "b is null"); }
void *pslot_batch_stream(const pslot_batch *b) { return b ? (void *)b->E.st : nullptr; }
int pslot_batch_kernel_launches(const pslot_batch *b)
{
    if (!b) return fail(PSLOT_ERR_ARG, "b is null");
    // ingest (re-pitch / colour conversion, one per view when used) + pyramid + FAST + quadtree (2) + describe + 2 stereo
    return b->launches ? b->launches : b->E.p.nlevels - 1 + 3 + (b->E.p.nlevels > 2 ? 1 : 0) + 2 + 2;
}

// Staged uploads go through the shared upload stream when one is set (pslot_batch_set_copy_stream).
static cudaStream_t upload_begin(pslot_batch *b)
{
    if (!b->copy_st) return b->E.st;
    if (b->stage_busy) cudaStreamWaitEvent(b->copy_st, b->ev_stage_free, 0); // previous ingest kernel has drained the stage
    return b->copy_st;
}
static void upload_end(pslot_batch *b)
{
    if (!b->copy_st) return;
    cudaEventRecord(b->ev_uploaded, b->copy_st);
    cudaStreamWaitEvent(b->E.st, b->ev_uploaded, 0);
}

static int batch_submit_impl(pslot_batch *b, const uint8_t *left, const uint8_t *right, ptrdiff_t pitch,
                             ptrdiff_t image_stride, int npairs, int host_input, int channels, int rgb, float bf,
                             float max_d, const pslot_batch_out *out)
{
    if (!b || !left || !right) return fail(PSLOT_ERR_ARG, "null argument");
    if (npairs < 1 || npairs > b->max_pairs) return fail(PSLOT_ERR_ARG, "npairs must be in [1,%d]", b->max_pairs);
    if (channels != 1 && channels != 3 && channels != 4) return fail(PSLOT_ERR_ARG, "channels must be 1, 3 or 4");
    Engine &E = b->E;
    if (pitch < (ptrdiff_t)E.W * channels || image_stride < pitch * E.H) return fail(PSLOT_ERR_ARG, "bad pitch / image stride");
    if (channels != 1 && !host_input) return fail(PSLOT_ERR_ARG, "colour input must be host memory");
    CU(cudaSetDevice(E.device));
    const cudaMemcpyKind kind = host_input ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    const LevelGeom &L0 = E.g.lv[0];
    cudaEvent_t *ev = b->profiling ? b->ev : nullptr;
    if (ev) cudaEventRecord(ev[0], E.st);
    const uint8_t *views[2] = {left, right};
    int ingest_launches = 0;
    for (int v = 0; v < 2; ++v) {
        uint8_t *dst = E.pyr.lv[0] + (size_t)v * npairs * E.pyr.slot_stride[0];
        if (channels != 1) {
            // colour frames: tight upload, then one conversion kernel per view writes level 0 directly
            const size_t per_img = (size_t)E.W * E.H * channels;
            const size_t view_bytes = align_up(per_img * b->max_pairs + 16, 256); // keeps each view word-aligned
            const size_t need = 2 * view_bytes;
            if (b->cstage_bytes < need) {
                cudaFree(b->d_cstage);
                b->d_cstage = nullptr; b->cstage_bytes = 0;
                CU(cudaMalloc(&b->d_cstage, need));
                b->cstage_bytes = need;
            }
            uint8_t *stg = b->d_cstage + (size_t)v * view_bytes;
            cudaStream_t up = upload_begin(b);
            if (pitch == (ptrdiff_t)E.W * channels && image_stride == pitch * E.H) {
                CU(cudaMemcpyAsync(stg, views[v], per_img * npairs, cudaMemcpyHostToDevice, up));
            } else {
                for (int i = 0; i < npairs; ++i)
                    CU(cudaMemcpy2DAsync(stg + (size_t)i * per_img, (size_t)E.W * channels, views[v] + (size_t)i * image_stride,
                                         pitch, (size_t)E.W * channels, E.H, cudaMemcpyHostToDevice, up));
            }
            upload_end(b);
            ingest_launches += launch_gray_ingest(stg, per_img, channels, rgb, dst, E.W, E.H, L0.pitch, E.pyr.slot_stride[0], npairs, E.st);
        } else if (pitch == E.W && image_stride == pitch * E.H && pitch != L0.pitch) {
            // tightly packed gray frames: (host) one linear copy at full PCIe rate into the staging buffer, then a
            // re-pitch kernel into the padded level-0 rows (an SM kernel: keeps the copy engines free for H2D / D2H)
            const uint8_t *tight = views[v];
            if (host_input) {
                const size_t view_bytes = align_up((size_t)E.W * E.H * b->max_pairs + 16, 256);
                if (!b->d_stage) CU(cudaMalloc(&b->d_stage, 2 * view_bytes));
                uint8_t *stg = b->d_stage + (size_t)v * view_bytes;
                cudaStream_t up = upload_begin(b);
                CU(cudaMemcpyAsync(stg, views[v], (size_t)E.W * E.H * npairs, cudaMemcpyHostToDevice, up));
                upload_end(b);
                tight = stg;
            }
            ingest_launches += launch_gray_ingest(tight, (size_t)E.W * E.H, 1, 0, dst, E.W, E.H, L0.pitch, E.pyr.slot_stride[0], npairs, E.st);
        } else if (image_stride == pitch * E.H) {
            CU(cudaMemcpy2DAsync(dst, L0.pitch, views[v], pitch, E.W, (size_t)E.H * npairs, kind, E.st));
        } else {
            for (int i = 0; i < npairs; ++i)
                CU(cudaMemcpy2DAsync(dst + (size_t)i * E.pyr.slot_stride[0], L0.pitch, views[v] + (size_t)i * image_stride,
                                     pitch, E.W, E.H, kind, E.st));
        }
    }
    if (b->copy_st) { // the staging buffers may be overwritten once the ingest kernels above have run
        cudaEventRecord(b->ev_stage_free, E.st);
        b->stage_busy = true;
    }
    int rc = E.run_extract(2 * npairs, ev);
    if (rc < 0) return rc;
    rc += ingest_launches;
    const int K = E.g.kp_cap_total;
    StereoArgs a{};
    a.kps_l = E.d_out_kps; a.kps_r = E.d_out_kps + (size_t)npairs * K;
    a.desc_l = E.d_out_desc; a.desc_r = E.d_out_desc + (size_t)npairs * K * 32;
    a.n_l = E.d_out_count; a.n_r = E.d_out_count + npairs;
    a.K = K;
    a.pyr_l = E.pyr; a.pyr_r = E.pyr;
    for (int l = 0; l < E.g.nlevels; ++l) a.pyr_r.lv[l] = E.pyr.lv[l] + (size_t)npairs * E.pyr.slot_stride[l];
    a.bf = bf; a.max_d = max_d;
    a.u_right = E.d_uright; a.depth = E.d_depth; a.best_dist = E.d_best_dist; a.best_idx = E.d_best_idx; a.sad = E.d_sad;
    a.err_flag = E.d_err;
    a.counters = b->profiling ? b->d_counters : nullptr;
    rc += launch_stereo(E.g, a, npairs, K, E.st, ev ? ev[6] : nullptr);
    if (ev) cudaEventRecord(ev[7], E.st);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PSLOT_ERR_CUDA, "stereo launch failed: %s", cudaGetErrorString(e));
    b->launches = rc;
    b->last_pairs = npairs;
    ++b->submits;
    if (out) {
        const size_t ni = 2 * (size_t)npairs;
        if (out->counts) CU(cudaMemcpyAsync(out->counts, E.d_out_count, 4 * ni, cudaMemcpyDeviceToHost, E.st));
        if (out->kps) CU(cudaMemcpyAsync(out->kps, E.d_out_kps, sizeof(pslot_keypoint) * ni * K, cudaMemcpyDeviceToHost, E.st));
        if (out->desc) CU(cudaMemcpyAsync(out->desc, E.d_out_desc, 32 * ni * K, cudaMemcpyDeviceToHost, E.st));
        if (out->u_right) CU(cudaMemcpyAsync(out->u_right, E.d_uright, 4 * (size_t)npairs * K, cudaMemcpyDeviceToHost, E.st));
        if (out->depth) CU(cudaMemcpyAsync(out->depth, E.d_depth, 4 * (size_t)npairs * K, cudaMemcpyDeviceToHost, E.st));
    }
    if (ev) cudaEventRecord(ev[8], E.st);
    b->profiled = ev != nullptr;
    return PSLOT_OK;
}

int pslot_batch_submit(pslot_batch *b, const uint8_t *left, const uint8_t *right, ptrdiff_t pitch, ptrdiff_t image_stride,
                       int npairs, int host_input, float bf, float max_d, const pslot_batch_out *out)
{
    return batch_submit_impl(b, left, right, pitch, image_stride, npairs, host_input, 1, 0, bf, max_d, out);
}

int pslot_batch_submit_color(pslot_batch *b, const uint8_t *left, const uint8_t *right, ptrdiff_t pitch,
                             ptrdiff_t image_stride, int npairs, int channels, int rgb_order, float bf, float max_d,
                             const pslot_batch_out *out)
{
    return batch_submit_impl(b, left, right, pitch, image_stride, npairs, 1, channels, rgb_order, bf, max_d, out);
}

int pslot_batch_wait(pslot_batch *b)
{
    if (!b) return fail(PSLOT_ERR_ARG, "b is null");
    CU(cudaSetDevice(b->E.device));
    return check_err_flag(b->E);
}

int pslot_batch_set_copy_stream(pslot_batch *b, void *stream)
{
    if (!b) return fail(PSLOT_ERR_ARG, "b is null");
    CU(cudaSetDevice(b->E.device));
    CU(cudaStreamSynchronize(b->E.st));
    if (stream && !b->ev_uploaded) {
        CU(cudaEventCreateWithFlags(&b->ev_uploaded, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&b->ev_stage_free, cudaEventDisableTiming));
    }
    b->copy_st = (cudaStream_t)stream;
    b->stage_busy = false;
    return PSLOT_OK;
}

int pslot_batch_set_profiling(pslot_batch *b, int enable)
{
    if (!b) return fail(PSLOT_ERR_ARG, "b is null");
    CU(cudaSetDevice(b->E.device));
    if (enable) {
        for (auto &e : b->ev)
            if (!e) CU(cudaEventCreate(&e));
        if (!b->d_counters) {
            CU(cudaMalloc(&b->d_counters, 16));
            CU(cudaMemset(b->d_counters, 0, 16));
        }
    }
    b->profiling = enable != 0;
    b->profiled = false;
    return PSLOT_OK;
}

int pslot_batch_stage_times(pslot_batch *b, float *ms)
{
    if (!b || !ms) return fail(PSLOT_ERR_ARG, "null argument");
    if (!b->profiled) return fail(PSLOT_ERR_STATE, "no profiled submit to report");
    CU(cudaSetDevice(b->E.device));
    CU(cudaEventSynchronize(b->ev[PSLOT_NUM_STAGES]));
    for (int i = 0; i < PSLOT_NUM_STAGES; ++i) CU(cudaEventElapsedTime(&ms[i], b->ev[i], b->ev[i + 1]));
    return PSLOT_OK;
}

int pslot_batch_stats(pslot_batch *b, int64_t stats[4])
{
    if (!b || !stats) return fail(PSLOT_ERR_ARG, "null argument");
    Engine &E = b->E;
    const int np = b->last_pairs;
    if (np <= 0) return fail(PSLOT_ERR_STATE, "no batch has been submitted");
    CU(cudaSetDevice(E.device));
    CU(cudaStreamSynchronize(E.st));
    const size_t K = (size_t)E.g.kp_cap_total;
    std::vector<int> cells((size_t)2 * np * E.g.cells_per_slot), counts((size_t)2 * np), bd((size_t)np * K);
    std::vector<float> ur((size_t)np * K);
    CU(cudaMemcpy(cells.data(), E.d_cell_count, 4 * cells.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(counts.data(), E.d_out_count, 4 * counts.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(bd.data(), E.d_best_dist, 4 * bd.size(), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(ur.data(), E.d_uright, 4 * ur.size(), cudaMemcpyDeviceToHost));
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    for (int v : cells) stats[0] += v;
    for (int v : counts) stats[1] += v;
    for (int p = 0; p < np; ++p)
        for (int i = 0; i < counts[p]; ++i) {
            stats[2] += bd[p * K + i] < 75;
            stats[3] += ur[p * K + i] >= 0.f;
        }
    return PSLOT_OK;
}

int pslot_batch_stereo_counters(pslot_batch *b, int64_t out[2])
{
    if (!b || !out) return fail(PSLOT_ERR_ARG, "null argument");
    if (!b->d_counters) return fail(PSLOT_ERR_STATE, "profiling has not been enabled on this context");
    CU(cudaSetDevice(b->E.device));
    CU(cudaStreamSynchronize(b->E.st));
    unsigned long long h[2];
    CU(cudaMemcpy(h, b->d_counters, 16, cudaMemcpyDeviceToHost));
    CU(cudaMemset(b->d_counters, 0, 16));
    out[0] = (int64_t)h[0]; out[1] = (int64_t)h[1];
    return PSLOT_OK;
}

static int check_grid_spec(const pslot_grid_spec *spec)
{
    if (!spec) return fail(PSLOT_ERR_ARG, "spec is null");
    if (spec->cols < 1 || spec->rows < 1 || (long long)spec->cols * spec->rows > 16384)
        return fail(PSLOT_ERR_ARG, "grid must have 1..16384 cells");
    return PSLOT_OK;
}

int pslot_assign_grid(pslot_orb *ctx, const pslot_keypoint *kps_un, int n, const pslot_grid_spec *spec,
                      int32_t *cell_start, int32_t *indices, int *n_in_grid)
{
    if (!ctx) return fail(PSLOT_ERR_ARG, "ctx is null");
    if (n < 0 || (n > 0 && !kps_un)) return fail(PSLOT_ERR_ARG, "bad keypoint array");
    if (int rc = check_grid_spec(spec)) return rc;
    Engine &E = ctx->E;
    CU(cudaSetDevice(E.device));
    const int C = spec->cols * spec->rows;
    if (n > ctx->gk_cap) {
        cudaFree(ctx->d_gk); cudaFree(ctx->d_gidx);
        ctx->d_gk = nullptr; ctx->d_gidx = nullptr; ctx->gk_cap = 0;
        const int cap = std::max(n, 2048);
        CU(cudaMalloc(&ctx->d_gk, sizeof(pslot_keypoint) * (size_t)cap));
        CU(cudaMalloc(&ctx->d_gidx, sizeof(int) * (size_t)cap));
        ctx->gk_cap = cap;
    }
    if (C + 1 > ctx->gcs_cap) {
        cudaFree(ctx->d_gcs);
        ctx->d_gcs = nullptr; ctx->gcs_cap = 0;
        CU(cudaMalloc(&ctx->d_gcs, sizeof(int) * (size_t)(C + 1)));
        ctx->gcs_cap = C + 1;
    }
    cudaStream_t st = E.st;
    ctx->grid_n = -1;
    ctx->grid_token = 0;
    if (n) CU(cudaMemcpyAsync(ctx->d_gk, kps_un, sizeof(pslot_keypoint) * (size_t)n, cudaMemcpyHostToDevice, st));
    launch_assign_grid(ctx->d_gk, nullptr, n, (size_t)std::max(n, 1), *spec, ctx->d_gcs, ctx->d_gidx, 1, st);
    CU(cudaGetLastError());
    int total = 0;
    CU(cudaMemcpyAsync(&total, ctx->d_gcs + C, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (cell_start) CU(cudaMemcpyAsync(cell_start, ctx->d_gcs, sizeof(int) * (size_t)(C + 1), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (indices && total) CU(cudaMemcpy(indices, ctx->d_gidx, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost));
    if (n_in_grid) *n_in_grid = total;
    static std::atomic<uint64_t> generation{0}; // process-wide, so tokens of different contexts never collide either
    ctx->grid_n = n;
    ctx->grid_token = ++generation;
    ctx->grid_spec = *spec;
    ctx->h_gangle.resize((size_t)n);
    for (int i = 0; i < n; ++i) ctx->h_gangle[i] = kps_un[i].angle;
    return PSLOT_OK;
}

// ---- Frame::AssignFeatures ---------------------------------------------------------------------------------------------
int pslot_assign_features(pslot_orb *ctx, const pslot_keypoint *kps, int n, const uint8_t *mask, int w, int h, ptrdiff_t mask_pitch,
                          const int32_t *object_ids, int n_objects, const float *depth, float depth_limit, int32_t *cls,
                          int32_t *order, int32_t *class_start, int32_t *effective)
{
    if (!ctx) return fail(PSLOT_ERR_ARG, "ctx is null");
    if (n < 0 || (n > 0 && (!kps || !cls || !order))) return fail(PSLOT_ERR_ARG, "bad feature arrays");
    if (n_objects < 0 || n_objects > 256 || (n_objects > 0 && !object_ids)) return fail(PSLOT_ERR_ARG, "n_objects must be in [0,256]");
    if (!class_start) return fail(PSLOT_ERR_ARG, "class_start is null");
    if (w <= 0 || h <= 0) return fail(PSLOT_ERR_ARG, "bad mask size");
    if (mask && mask_pitch < w) return fail(PSLOT_ERR_ARG, "pitch < width");
    Engine &E = ctx->E;
    CU(cudaSetDevice(E.device));
    cudaStream_t st = E.st;
    if (mask) {
        const size_t need = (size_t)w * h;
        if (need > ctx->pmask_bytes) {
            CU(cudaStreamSynchronize(st));
            cudaFree(ctx->d_pmask);
            ctx->d_pmask = nullptr; ctx->pmask_bytes = 0;
            CU(cudaMalloc(&ctx->d_pmask, need));
            ctx->pmask_bytes = need;
        }
        CU(cudaMemcpy2DAsync(ctx->d_pmask, (size_t)w, mask, (size_t)mask_pitch, (size_t)w, (size_t)h, cudaMemcpyHostToDevice, st));
        ctx->pmask_w = w; ctx->pmask_h = h;
    } else if (!ctx->d_pmask || ctx->pmask_w != w || ctx->pmask_h != h) {
        return fail(PSLOT_ERR_STATE, "mask is null and no mask of this size was uploaded by an earlier call");
    }
    // arena layout
    const size_t N = (size_t)std::max(n, 1), J = (size_t)std::max(n_objects, 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_k = take(sizeof(pslot_keypoint) * N), o_d = take(4 * N), o_id = take(4 * J), o_cls = take(4 * N),
                 o_ord = take(4 * N), o_cs = take(4 * (J + 2)), o_eff = take(4 * J);
    if (off > ctx->part_bytes) {
        CU(cudaStreamSynchronize(st));
        cudaFree(ctx->d_part);
        ctx->d_part = nullptr; ctx->part_bytes = 0;
        const size_t cap = std::max<size_t>(2 * off, 1 << 18);
        CU(cudaMalloc(&ctx->d_part, cap));
        ctx->part_bytes = cap;
    }
    unsigned char *A = ctx->d_part;
    if (n) CU(cudaMemcpyAsync(A + o_k, kps, sizeof(pslot_keypoint) * (size_t)n, cudaMemcpyHostToDevice, st));
    if (n && depth) CU(cudaMemcpyAsync(A + o_d, depth, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    if (n_objects) CU(cudaMemcpyAsync(A + o_id, object_ids, 4 * (size_t)n_objects, cudaMemcpyHostToDevice, st));
    PartArgs a{};
    a.kps = reinterpret_cast<const pslot_keypoint *>(A + o_k); a.counts = nullptr; a.fixed_n = n; a.kp_stride = 0;
    a.mask = ctx->d_pmask; a.w = w; a.h = h; a.pitch = w; a.mask_stride = 0;
    a.object_ids = reinterpret_cast<const int *>(A + o_id); a.n_objects = n_objects; a.ids_stride = 0;
    a.depth = depth ? reinterpret_cast<const float *>(A + o_d) : nullptr; a.depth_stride = 0; a.depth_limit = depth_limit;
    a.cls = reinterpret_cast<int *>(A + o_cls); a.order = reinterpret_cast<int *>(A + o_ord); a.out_stride = 0;
    a.class_start = reinterpret_cast<int *>(A + o_cs);
    a.effective = reinterpret_cast<int *>(A + o_eff);
    launch_assign_features(a, 1, st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(class_start, A + o_cs, 4 * (size_t)(n_objects + 2), cudaMemcpyDeviceToHost, st));
    if (n) CU(cudaMemcpyAsync(cls, A + o_cls, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (effective && n_objects) CU(cudaMemcpyAsync(effective, A + o_eff, 4 * (size_t)n_objects, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const int kept = class_start[n_objects + 1];
    if (kept > 0) CU(cudaMemcpy(order, A + o_ord, 4 * (size_t)kept, cudaMemcpyDeviceToHost));
    return PSLOT_OK;
}

int pslot_batch_assign_features(pslot_batch *b, const uint8_t *masks, ptrdiff_t mask_pitch, ptrdiff_t mask_image_stride,
                                const int32_t *object_ids, int n_objects, int32_t *cls, int32_t *order, int32_t *class_start)
{
    if (!b) return fail(PSLOT_ERR_ARG, "b is null");
    if (b->last_pairs <= 0) return fail(PSLOT_ERR_STATE, "no batch has been submitted");
    if (!masks || !cls || !order || !class_start) return fail(PSLOT_ERR_ARG, "null argument");
    if (n_objects < 0 || n_objects > 256 || (n_objects > 0 && !object_ids)) return fail(PSLOT_ERR_ARG, "n_objects must be in [0,256]");
    Engine &E = b->E;
    if (mask_pitch < E.W || mask_image_stride < mask_pitch * E.H) return fail(PSLOT_ERR_ARG, "bad mask pitch / stride");
    CU(cudaSetDevice(E.device));
    const int np = b->last_pairs;
    const size_t K = (size_t)E.g.kp_cap_total, P = (size_t)b->max_pairs, J = (size_t)std::max(n_objects, 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_m = take((size_t)E.W * E.H * P), o_id = take(4 * J * P), o_cls = take(4 * K * P), o_ord = take(4 * K * P),
                 o_cs = take(4 * (J + 2) * P);
    if (off > b->part_bytes) {
        CU(cudaStreamSynchronize(E.st));
        cudaFree(b->d_part);
        b->d_part = nullptr; b->part_bytes = 0;
        CU(cudaMalloc(&b->d_part, off));
        CU(cudaMemset(b->d_part, 0, off));
        b->part_bytes = off;
    }
    unsigned char *A = b->d_part;
    if (mask_image_stride == mask_pitch * E.H)
        CU(cudaMemcpy2DAsync(A + o_m, (size_t)E.W, masks, (size_t)mask_pitch, (size_t)E.W, (size_t)E.H * np, cudaMemcpyHostToDevice, E.st));
    else // images not back to back: one copy per image
        for (int i = 0; i < np; ++i)
            CU(cudaMemcpy2DAsync(A + o_m + (size_t)i * E.W * E.H, (size_t)E.W, masks + (size_t)i * mask_image_stride, (size_t)mask_pitch,
                                 (size_t)E.W, (size_t)E.H, cudaMemcpyHostToDevice, E.st));
    if (n_objects) CU(cudaMemcpyAsync(A + o_id, object_ids, 4 * (size_t)n_objects * np, cudaMemcpyHostToDevice, E.st));
    PartArgs a{};
    // left images are slots 0..np-1 of the batch: counts and keypoint rows are the first np entries
    a.kps = E.d_out_kps; a.counts = E.d_out_count; a.fixed_n = 0; a.kp_stride = K;
    a.mask = A + o_m; a.w = E.W; a.h = E.H; a.pitch = E.W; a.mask_stride = (size_t)E.W * E.H;
    a.object_ids = reinterpret_cast<const int *>(A + o_id); a.n_objects = n_objects; a.ids_stride = (size_t)n_objects;
    a.depth = nullptr; a.depth_stride = 0; a.depth_limit = 0.f;
    a.cls = reinterpret_cast<int *>(A + o_cls); a.order = reinterpret_cast<int *>(A + o_ord); a.out_stride = K;
    a.class_start = reinterpret_cast<int *>(A + o_cs);
    a.effective = nullptr;
    launch_assign_features(a, np, E.st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(class_start, A + o_cs, 4 * (size_t)(n_objects + 2) * np, cudaMemcpyDeviceToHost, E.st));
    CU(cudaMemcpyAsync(cls, A + o_cls, 4 * K * np, cudaMemcpyDeviceToHost, E.st));
    CU(cudaMemcpyAsync(order, A + o_ord, 4 * K * np, cudaMemcpyDeviceToHost, E.st));
    return PSLOT_OK;
}

uint64_t pslot_grid_token(const pslot_orb *ctx) { return ctx ? ctx->grid_token : 0; }

int pslot_features_in_area(pslot_orb *ctx, const pslot_area_query *queries, int nq, int cap, int32_t *counts,
                           int32_t *indices)
{
    if (!ctx) return fail(PSLOT_ERR_ARG, "ctx is null");
    if (ctx->grid_n < 0) return fail(PSLOT_ERR_STATE, "pslot_assign_grid has not been called on this context");
    if (nq < 0 || cap < 0 || (nq > 0 && (!queries || !counts)) || (nq > 0 && cap > 0 && !indices))
        return fail(PSLOT_ERR_ARG, "bad query arguments");
    if (nq == 0) return PSLOT_OK;
    Engine &E = ctx->E;
    CU(cudaSetDevice(E.device));
    if (nq > ctx->q_cap) {
        cudaFree(ctx->d_q); cudaFree(ctx->d_qcount);
        ctx->d_q = nullptr; ctx->d_qcount = nullptr; ctx->q_cap = 0;
        const int c = std::max(nq, 1024);
        CU(cudaMalloc(&ctx->d_q, sizeof(pslot_area_query) * (size_t)c));
        CU(cudaMalloc(&ctx->d_qcount, sizeof(int) * (size_t)c));
        ctx->q_cap = c;
    }
    const size_t need = (size_t)nq * (size_t)std::max(cap, 1);
    if (need > ctx->qout_cap) {
        cudaFree(ctx->d_qout);
        ctx->d_qout = nullptr; ctx->qout_cap = 0;
        CU(cudaMalloc(&ctx->d_qout, sizeof(int) * need));
        ctx->qout_cap = need;
    }
    cudaStream_t st = E.st;
    CU(cudaMemcpyAsync(ctx->d_q, queries, sizeof(pslot_area_query) * (size_t)nq, cudaMemcpyHostToDevice, st));
    launch_features_in_area(ctx->d_gk, ctx->d_gcs, ctx->d_gidx, ctx->grid_spec, ctx->d_q, nq, cap, ctx->d_qcount,
                            ctx->d_qout, st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(counts, ctx->d_qcount, sizeof(int) * (size_t)nq, cudaMemcpyDeviceToHost, st));
    if (cap > 0) CU(cudaMemcpyAsync(indices, ctx->d_qout, sizeof(int) * (size_t)nq * cap, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return PSLOT_OK;
}

// Common part of the projection searches: upload the flattened points and frame arrays, run sweeps of
// k_search_by_projection until a sweep changes no match (see grid.cu), download match_of_pt.
static int run_projection(pslot_orb *ctx, const std::vector<ProjPoint> &pts, const uint8_t *pt_desc,
                          const uint8_t *frame_desc, const float *u_right, const uint8_t *occupied, int n, int mode,
                          float nn_ratio, int32_t *match_of_pt, int th_dist = 100 /* ORBmatcher::TH_HIGH */)
{
    const int n_pts = (int)pts.size();
    Engine &E = ctx->E;
    CU(cudaSetDevice(E.device));
    // one scratch block: points | their descriptors | frame descriptors | uRight | occupied | 2 claim arrays |
    // 2 match arrays | changed flags
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_mp = take(sizeof(ProjPoint) * (size_t)n_pts), o_md = take(32 * (size_t)n_pts);
    const size_t o_fd = take(32 * (size_t)n), o_ur = take(4 * (size_t)n), o_oc = take((size_t)n);
    const size_t o_c0 = take(4 * (size_t)n), o_c1 = take(4 * (size_t)n);
    const size_t o_m0 = take(4 * (size_t)n_pts), o_m1 = take(4 * (size_t)n_pts), o_ch = take(64);
    if (off > ctx->proj_bytes) {
        cudaFree(ctx->d_proj);
        ctx->d_proj = nullptr; ctx->proj_bytes = 0;
        CU(cudaMalloc(&ctx->d_proj, off));
        ctx->proj_bytes = off;
    }
    uint8_t *base = (uint8_t *)ctx->d_proj;
    cudaStream_t st = E.st;
    CU(cudaMemcpyAsync(base + o_mp, pts.data(), sizeof(ProjPoint) * (size_t)n_pts, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + o_md, pt_desc, 32 * (size_t)n_pts, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + o_fd, frame_desc, 32 * (size_t)n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + o_ur, u_right, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    if (occupied) CU(cudaMemcpyAsync(base + o_oc, occupied, (size_t)n, cudaMemcpyHostToDevice, st));
    else CU(cudaMemsetAsync(base + o_oc, 0, (size_t)n, st));
    int *claim[2] = {(int *)(base + o_c0), (int *)(base + o_c1)};
    int *match[2] = {(int *)(base + o_m0), (int *)(base + o_m1)};
    CU(cudaMemsetAsync(claim[0], 0x7f, 4 * (size_t)n, st));  // "no map point has claimed this feature"
    CU(cudaMemsetAsync(match[0], 0xfe, 4 * (size_t)n_pts, st)); // differs from every possible match: the first sweep counts as a change
    ProjArgs a{};
    a.kps = ctx->d_gk; a.cell_start = ctx->d_gcs; a.indices = ctx->d_gidx; a.s = ctx->grid_spec;
    a.pts = (const ProjPoint *)(base + o_mp); a.pt_desc = base + o_md; a.n_pts = n_pts;
    a.frame_desc = base + o_fd; a.u_right = (const float *)(base + o_ur); a.occupied = base + o_oc;
    a.mode = mode; a.nn_ratio = nn_ratio; a.th_dist = th_dist;
    // Sweeps are enqueued four at a time and the host looks at their "changed" flags once per group: a sweep past the
    // fixed point recomputes the same matches, so overshooting is harmless and saves a round trip per sweep.
    const int GROUP = 4;
    int cur = 0, sweeps = 0;
    for (bool done = false; !done;) {
        if (sweeps > n_pts + GROUP) return fail(PSLOT_ERR_STATE, "projection search did not converge"); // cannot happen: see grid.cu
        int *flags = (int *)(base + o_ch);
        CU(cudaMemsetAsync(flags, 0, 4 * GROUP, st));
        for (int k = 0; k < GROUP; ++k) {
            a.claim_prev = claim[cur]; a.claim_next = claim[cur ^ 1];
            a.match_prev = match[cur]; a.match_next = match[cur ^ 1];
            a.changed = flags + k;
            CU(cudaMemsetAsync(a.claim_next, 0x7f, 4 * (size_t)n, st));
            launch_search_by_projection(a, st);
            cur ^= 1;
        }
        CU(cudaGetLastError());
        int changed[GROUP];
        CU(cudaMemcpyAsync(changed, flags, 4 * GROUP, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (int k = 0; k < GROUP; ++k) {
            ++sweeps;
            if (!changed[k]) { done = true; break; } // the later sweeps of the group reproduced this one
        }
    }
    CU(cudaMemcpy(match_of_pt, match[cur], 4 * (size_t)n_pts, cudaMemcpyDeviceToHost));
    ctx->last_sweeps = sweeps;
    return PSLOT_OK;
}

static int check_projection_frame(pslot_orb *ctx, int n_pts, const void *pts, const void *desc, const void *out,
                                  const uint8_t *frame_desc, const float *u_right, int n)
{
    if (!ctx) return fail(PSLOT_ERR_ARG, "ctx is null");
    if (ctx->grid_n < 0) return fail(PSLOT_ERR_STATE, "pslot_assign_grid has not been called on this context");
    if (n != ctx->grid_n) return fail(PSLOT_ERR_ARG, "n (%d) differs from the gridded keypoint count (%d)", n, ctx->grid_n);
    if (n_pts < 0 || (n_pts > 0 && (!pts || !desc || !out)) || (n > 0 && (!frame_desc || !u_right)))
        return fail(PSLOT_ERR_ARG, "null array");
    return PSLOT_OK;
}

static int search_by_projection_impl(pslot_orb *ctx, const pslot_map_point *mps, const uint8_t *mp_desc, int n_mp,
                                     const uint8_t *frame_desc, const float *u_right, const uint8_t *occupied, int n,
                                     const float *scale_factors, int nlevels, float th, float nn_ratio,
                                     int32_t *match_of_mp, int *nmatches, float fixed_window, int level_up, int th_dist);

int pslot_search_by_projection(pslot_orb *ctx, const pslot_map_point *mps, const uint8_t *mp_desc, int n_mp,
                               const uint8_t *frame_desc, const float *u_right, const uint8_t *occupied, int n,
                               const float *scale_factors, int nlevels, float th, float nn_ratio,
                               int32_t *match_of_mp, int *nmatches)
{
    return search_by_projection_impl(ctx, mps, mp_desc, n_mp, frame_desc, u_right, occupied, n, scale_factors, nlevels, th,
                                     nn_ratio, match_of_mp, nmatches, 0.f, 0, 100 /* TH_HIGH */);
}

// ORBmatcher::SearchByProjection(F, nOrder, vpMapObjectPoints, th) (ORBmatcher.cc:157-246): the local-map search over one
// detected object's own gridded feature set, with a fixed 5 px window, levels [L - 1, L + 1] and TH_HIGH_FORDYNAMIC
int pslot_search_by_projection_obj(pslot_orb *ctx, const pslot_map_point *mps, const uint8_t *mp_desc, int n_mp,
                                   const uint8_t *frame_desc, const float *u_right, const uint8_t *occupied, int n,
                                   const float *scale_factors, int nlevels, float th, float nn_ratio,
                                   int32_t *match_of_mp, int *nmatches)
{
    return search_by_projection_impl(ctx, mps, mp_desc, n_mp, frame_desc, u_right, occupied, n, scale_factors, nlevels, th,
                                     nn_ratio, match_of_mp, nmatches, 5.f, 1, 130 /* TH_HIGH_FORDYNAMIC */);
}

static int search_by_projection_impl(pslot_orb *ctx, const pslot_map_point *mps, const uint8_t *mp_desc, int n_mp,
                                     const uint8_t *frame_desc, const float *u_right, const uint8_t *occupied, int n,
                                     const float *scale_factors, int nlevels, float th, float nn_ratio,
                                     int32_t *match_of_mp, int *nmatches, float fixed_window, int level_up, int th_dist)
{
    if (int rc = check_projection_frame(ctx, n_mp, mps, mp_desc, match_of_mp, frame_desc, u_right, n)) return rc;
    if (!scale_factors || nlevels < 1 || nlevels > PSLOT_MAX_LEVELS) return fail(PSLOT_ERR_ARG, "bad scale factor table");
    if (nmatches) *nmatches = 0;
    if (n_mp == 0) return PSLOT_OK;
    std::vector<ProjPoint> pts((size_t)n_mp);
    for (int i = 0; i < n_mp; ++i) {
        const pslot_map_point &m = mps[i];
        ProjPoint &p = pts[i];
        p = ProjPoint{};
        p.flags = m.flags & 3;
        if (!(m.flags & 1)) continue;
        if (m.level < 0 || m.level >= nlevels) return fail(PSLOT_ERR_ARG, "map point %d: predicted level %d out of range", i, m.level);
        // The size of the window depends on the viewing direction (:85-92); float products, one rounding each
        float r = ((double)m.view_cos > 0.998) ? 2.5f : 4.0f; // RadiusByViewingCos (:252-258)
        if (th != 1.0f) r *= th;
        p.x = m.proj_x; p.y = m.proj_y; p.xr = m.proj_xr;
        p.er_limit = r * scale_factors[m.level];          // er > r * F.mvScaleFactors[nPredictedLevel] (:118, :215)
        p.radius = fixed_window > 0 ? fixed_window : p.er_limit; // (:92) / the literal 5 px of the object overload (:179)
        p.min_level = m.level - 1; p.max_level = m.level + level_up; // (:92) / (:179)
    }
    if (n == 0) { // no features: nothing can match
        for (int i = 0; i < n_mp; ++i) match_of_mp[i] = -1;
        return PSLOT_OK;
    }
    if (int rc = run_projection(ctx, pts, mp_desc, frame_desc, u_right, occupied, n, 0, nn_ratio, match_of_mp, th_dist)) return rc;
    if (nmatches) {
        int c = 0;
        for (int i = 0; i < n_mp; ++i) c += match_of_mp[i] >= 0;
        *nmatches = c;
    }
    return PSLOT_OK;
}

static int rotation_check(pslot_orb *ctx, const pslot_proj_point *pts_in, int n_pts, int check_orientation, float factor,
                          int32_t *match_of_pt, uint8_t *rejected, int *nmatches);

// ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:2529-2654) after the
// projection: the last-frame machinery without the stereo test, every association (pre-existing or made on the way)
// blocking a feature, acceptance at ORBdist, 30-degree rotation bins (`factor = 1.0f/HISTO_LENGTH`, :2542)
int pslot_search_by_projection_kf(pslot_orb *ctx, const pslot_proj_point *pts_in, const uint8_t *pt_desc, int n_pts,
                                  const uint8_t *frame_desc, const uint8_t *occupied, int n, int orb_dist,
                                  int check_orientation, int32_t *match_of_pt, uint8_t *rejected, int *nmatches)
{
    std::vector<float> no_stereo((size_t)std::max(n, 1), -1.0f); // mvuRight <= 0 everywhere: the stereo test never applies
    if (int rc = check_projection_frame(ctx, n_pts, pts_in, pt_desc, match_of_pt, frame_desc, no_stereo.data(), n)) return rc;
    if (n_pts > 0 && !rejected) return fail(PSLOT_ERR_ARG, "rejected is null");
    if (orb_dist < 0 || orb_dist > 256) return fail(PSLOT_ERR_ARG, "ORBdist must be in [0, 256]");
    if (nmatches) *nmatches = 0;
    if (n_pts == 0) return PSLOT_OK;
    for (int i = 0; i < n_pts; ++i) rejected[i] = 0;
    if (n == 0) {
        for (int i = 0; i < n_pts; ++i) match_of_pt[i] = -1;
        return PSLOT_OK;
    }
    std::vector<ProjPoint> pts((size_t)n_pts);
    for (int i = 0; i < n_pts; ++i) {
        const pslot_proj_point &q = pts_in[i];
        ProjPoint &p = pts[i];
        p = ProjPoint{};
        p.flags = (q.flags & 1) | 2; // CurrentFrame.mvpMapPoints[bestIdx2] = pMP blocks the feature whatever pMP observes (:2596)
        p.x = q.u; p.y = q.v; p.xr = 0.f; p.radius = q.radius; p.er_limit = q.radius;
        p.min_level = q.min_level; p.max_level = q.max_level;
    }
    if (int rc = run_projection(ctx, pts, pt_desc, frame_desc, no_stereo.data(), occupied, n, 1, 0.0f, match_of_pt, orb_dist)) return rc;
    return rotation_check(ctx, pts_in, n_pts, check_orientation, 1.0f / 30 /* (:2542) */, match_of_pt, rejected, nmatches);
}

int pslot_search_by_projection_last(pslot_orb *ctx, const pslot_proj_point *pts_in, const uint8_t *pt_desc, int n_pts,
                                    const uint8_t *frame_desc, const float *u_right, const uint8_t *occupied, int n,
                                    int check_orientation, int32_t *match_of_pt, uint8_t *rejected, int *nmatches)
{
    if (int rc = check_projection_frame(ctx, n_pts, pts_in, pt_desc, match_of_pt, frame_desc, u_right, n)) return rc;
    if (n_pts > 0 && !rejected) return fail(PSLOT_ERR_ARG, "rejected is null");
    if (nmatches) *nmatches = 0;
    if (n_pts == 0) return PSLOT_OK;
    for (int i = 0; i < n_pts; ++i) rejected[i] = 0;
    if (n == 0) {
        for (int i = 0; i < n_pts; ++i) match_of_pt[i] = -1;
        return PSLOT_OK;
    }
    std::vector<ProjPoint> pts((size_t)n_pts);
    for (int i = 0; i < n_pts; ++i) {
        const pslot_proj_point &q = pts_in[i];
        ProjPoint &p = pts[i];
        p = ProjPoint{};
        p.flags = q.flags & 3;
        p.x = q.u; p.y = q.v; p.xr = q.ur; p.radius = q.radius; p.er_limit = q.radius; // er > radius (:1692)
        p.min_level = q.min_level; p.max_level = q.max_level;
    }
    if (int rc = run_projection(ctx, pts, pt_desc, frame_desc, u_right, occupied, n, 1, 0.0f, match_of_pt)) return rc;
    return rotation_check(ctx, pts_in, n_pts, check_orientation, 30 / 360.0f /* factor (:1633) */, match_of_pt, rejected, nmatches);
}

// rotation consistency of the best-only searches (:1720-1753, :2616-2651), sequential and tiny: on the host
static int rotation_check(pslot_orb *ctx, const pslot_proj_point *pts_in, int n_pts, int check_orientation, float factor,
                          int32_t *match_of_pt, uint8_t *rejected, int *nmatches)
{
    int nm = 0;
    for (int i = 0; i < n_pts; ++i) nm += match_of_pt[i] >= 0;
    if (check_orientation) {
        const int HISTO_LENGTH = 30;
        std::vector<int> rotHist[HISTO_LENGTH];
        for (int i = 0; i < n_pts; ++i) {
            if (match_of_pt[i] < 0) continue;
            float rot = pts_in[i].angle - ctx->h_gangle[(size_t)match_of_pt[i]];
            if (rot < 0.0) rot += 360.0f;
            int bin = (int)round(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            if (bin < 0 || bin >= HISTO_LENGTH) return fail(PSLOT_ERR_ARG, "keypoint angle out of [0, 360]"); // the reference asserts
            rotHist[bin].push_back(i);
        }
        // ORBmatcher::ComputeThreeMaxima (:2658-2699)
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) {
            const int sz = (int)rotHist[i].size();
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; ind3 = ind2; ind2 = i; }
            else if (sz > max3) { max3 = sz; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int j : rotHist[i]) { rejected[j] = 1; --nm; }
    }
    if (nmatches) *nmatches = nm;
    return PSLOT_OK;
}

int pslot_debug_projection_sweeps(const pslot_orb *ctx) { return ctx ? ctx->last_sweeps : fail(PSLOT_ERR_ARG, "ctx is null"); }

int pslot_batch_assign_grid(pslot_batch *b, const pslot_grid_spec *spec, int32_t *cell_start, int32_t *indices)
{
    if (!b) return fail(PSLOT_ERR_ARG, "b is null");
    if (int rc = check_grid_spec(spec)) return rc;
    if (b->last_pairs <= 0) return fail(PSLOT_ERR_STATE, "no batch has been submitted");
    Engine &E = b->E;
    CU(cudaSetDevice(E.device));
    const int C = spec->cols * spec->rows, np = b->last_pairs;
    const size_t K = (size_t)E.g.kp_cap_total;
    if (!b->d_gidx) {
        CU(cudaMalloc(&b->d_gidx, sizeof(int) * K * (size_t)b->max_pairs));
        CU(cudaMemset(b->d_gidx, 0, sizeof(int) * K * (size_t)b->max_pairs));
    }
    if (C + 1 > b->gcs_cells) {
        CU(cudaStreamSynchronize(E.st));
        cudaFree(b->d_gcs);
        b->d_gcs = nullptr; b->gcs_cells = 0;
        CU(cudaMalloc(&b->d_gcs, sizeof(int) * (size_t)(C + 1) * (size_t)b->max_pairs));
        b->gcs_cells = C + 1;
    }
    // left images are slots 0..np-1 of the batch: counts and keypoint rows are the first np entries
    launch_assign_grid(E.d_out_kps, E.d_out_count, 0, K, *spec, b->d_gcs, b->d_gidx, np, E.st);
    CU(cudaGetLastError());
    if (cell_start) CU(cudaMemcpyAsync(cell_start, b->d_gcs, sizeof(int) * (size_t)(C + 1) * np, cudaMemcpyDeviceToHost, E.st));
    if (indices) CU(cudaMemcpyAsync(indices, b->d_gidx, sizeof(int) * K * np, cudaMemcpyDeviceToHost, E.st));
    return PSLOT_OK;
}

int pslot_batch_device_out(const pslot_batch *b, pslot_batch_out *dev)
{
    if (!b || !dev) return fail(PSLOT_ERR_ARG, "null argument");
    dev->counts = b->E.d_out_count; dev->kps = b->E.d_out_kps; dev->desc = b->E.d_out_desc;
    dev->u_right = b->E.d_uright; dev->depth = b->E.d_depth;
    return PSLOT_OK;
}

} // extern "C"
