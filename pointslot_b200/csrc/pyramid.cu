// pyramid.cu — K1: ComputePyramid (reference src/ORBextractor.cc:1107-1132).
//
// Level l is cv::resize(level l-1, INTER_LINEAR) for CV_8UC1, i.e. OpenCV's 11-bit fixed-point
// bilinear kernel; the taps (source indices + coefficients) are computed on the host exactly as
// OpenCV does and the device evaluates the integer formula
//     dst = ( ((c0y*(H0>>4))>>16) + ((c1y*(H1>>4))>>16) + 2 ) >> 2,   H = s0*c0x + s1*c1x.
// Integer-only, hence bit-exact.  Levels are chained (each from the previous one, :1120), so a batch
// needs nlevels-1 dependent launches; each launch covers every slot of the batch.
//
// Mapping: a warp covers 32 consecutive output columns (so one load instruction touches ~38 consecutive
// source bytes = one or two 32-byte sectors) and each thread walks RROWS consecutive output rows,
// re-using the horizontally interpolated source row shared by neighbouring output rows (scale 1.2:
// consecutive output rows share a source row 4 times out of 5).
//
// Bound: L1/LSU wavefronts on byte loads (gather at a non-integer stride); algorithmic HBM bytes per
// launch = S * (P_{l-1} + P_l).
#include "pslot_internal.h"

namespace {

constexpr int RX = 128;   // threads per CTA = output columns per CTA
constexpr int RROWS = 8;  // output rows per thread

__global__ void __launch_bounds__(RX)
k_resize_level(const uint8_t *__restrict__ src, int spitch, unsigned long long sstride,
               uint8_t *__restrict__ dst, int dw, int dh, int dpitch, unsigned long long dstride,
               const ResizeTap *__restrict__ xt, const ResizeTap *__restrict__ yt)
{
    const int dx = blockIdx.x * RX + threadIdx.x;
    if (dx >= dw) return;
    const int dy0 = blockIdx.y * RROWS;
    const uint8_t *s = src + (size_t)blockIdx.z * sstride;
    uint8_t *d = dst + (size_t)blockIdx.z * dstride + dx;
    const ResizeTap tx = xt[dx];
    const int x0 = tx.s0, x1 = tx.s1, cx0 = tx.c0, cx1 = tx.c1;
    int cached_row = -1, cached_h = 0;
#pragma unroll
    for (int r = 0; r < RROWS; ++r) {
        const int dy = dy0 + r;
        if (dy >= dh) break;
        const ResizeTap ty = yt[dy];
        int h0, h1;
        if (ty.s0 == cached_row) {
            h0 = cached_h;
        } else {
            const uint8_t *p = s + (size_t)ty.s0 * spitch;
            h0 = (int)__ldg(p + x0) * cx0 + (int)__ldg(p + x1) * cx1;
        }
        if (ty.s1 == ty.s0) {
            h1 = h0;
        } else {
            const uint8_t *p = s + (size_t)ty.s1 * spitch;
            h1 = (int)__ldg(p + x0) * cx0 + (int)__ldg(p + x1) * cx1;
        }
        cached_row = ty.s1;
        cached_h = h1;
        int v = ((((int)ty.c0 * (h0 >> 4)) >> 16) + (((int)ty.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        d[(size_t)dy * dpitch] = (uint8_t)v;
    }
}

} // namespace

int launch_pyramid(const Geom &g, const PyrPtrs &pyr, const ResizeTap *taps, int nslots, cudaStream_t st)
{
    int launches = 0;
    for (int l = 1; l < g.nlevels; ++l) {
        const LevelGeom &S = g.lv[l - 1], &D = g.lv[l];
        dim3 grid((D.w + RX - 1) / RX, (D.h + RROWS - 1) / RROWS, nslots);
        k_resize_level<<<grid, RX, 0, st>>>(pyr.lv[l - 1], S.pitch, pyr.slot_stride[l - 1], pyr.lv[l], D.w, D.h,
                                            D.pitch, pyr.slot_stride[l], taps + D.tab_x, taps + D.tab_y);
        ++launches;
    }
    return launches;
}
