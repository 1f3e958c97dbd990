// pyramid.cu — K1: ComputePyramid (reference src/ORBextractor.cc:1107-1132).
//
// Level l is cv::resize(level l-1, INTER_LINEAR) for CV_8UC1, i.e. OpenCV's 11-bit fixed-point
// bilinear kernel; the taps (source indices + coefficients) are computed on the host exactly as
// OpenCV does and the device evaluates the integer formula
//     dst = ( ((c0y*(H0>>4))>>16) + ((c1y*(H1>>4))>>16) + 2 ) >> 2,   H = s0*c0x + s1*c1x.
// Integer-only, hence bit-exact.  Levels are chained (each from the previous one, :1120), so a batch
// needs nlevels-1 dependent launches; each launch covers every slot of the batch.
//
// Bound: HBM/L2 streaming.  Algorithmic bytes per launch = S * (P_{l-1} + P_l).
#include "pslot_internal.h"

namespace {

constexpr int RX = 64; // threads along x, 4 output pixels each
constexpr int RY = 4;

__global__ void __launch_bounds__(RX *RY)
k_resize_level(const uint8_t *__restrict__ src, int spitch, unsigned long long sstride,
               uint8_t *__restrict__ dst, int dw, int dh, int dpitch, unsigned long long dstride,
               const ResizeTap *__restrict__ xt, const ResizeTap *__restrict__ yt)
{
    const int dx0 = (blockIdx.x * RX + threadIdx.x) * 4;
    const int dy = blockIdx.y * RY + threadIdx.y;
    if (dx0 >= dw || dy >= dh) return;
    const uint8_t *s = src + (size_t)blockIdx.z * sstride;
    uint8_t *d = dst + (size_t)blockIdx.z * dstride;
    const ResizeTap ty = yt[dy];
    const uint8_t *r0 = s + (size_t)ty.s0 * spitch;
    const uint8_t *r1 = s + (size_t)ty.s1 * spitch;
    // the tap table is padded to a multiple of 4 entries, so a 2 x 16-byte vector read is in bounds
    const uint4 ta = __ldg(reinterpret_cast<const uint4 *>(xt + dx0));
    const uint4 tb = __ldg(reinterpret_cast<const uint4 *>(xt + dx0 + 2));
    const uint32_t tw[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int s0 = (int)(tw[2 * k] & 0xFFFFu), s1 = (int)(tw[2 * k] >> 16);
        const int c0 = (int)(short)(tw[2 * k + 1] & 0xFFFFu), c1 = (int)(short)(tw[2 * k + 1] >> 16);
        const int h0 = (int)__ldg(r0 + s0) * c0 + (int)__ldg(r0 + s1) * c1;
        const int h1 = (int)__ldg(r1 + s0) * c0 + (int)__ldg(r1 + s1) * c1;
        int v = ((((int)ty.c0 * (h0 >> 4)) >> 16) + (((int)ty.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        out |= (uint32_t)v << (8 * k);
    }
    // pitch is a multiple of 128, so the word store is aligned; bytes past dw land in row padding
    *reinterpret_cast<uint32_t *>(d + (size_t)dy * dpitch + dx0) = out;
}

} // namespace

int launch_pyramid(const Geom &g, const PyrPtrs &pyr, const ResizeTap *taps, int nslots, cudaStream_t st)
{
    int launches = 0;
    for (int l = 1; l < g.nlevels; ++l) {
        const LevelGeom &S = g.lv[l - 1], &D = g.lv[l];
        dim3 block(RX, RY);
        dim3 grid((D.w + RX * 4 - 1) / (RX * 4), (D.h + RY - 1) / RY, nslots);
        k_resize_level<<<grid, block, 0, st>>>(pyr.lv[l - 1], S.pitch, pyr.slot_stride[l - 1], pyr.lv[l], D.w, D.h,
                                               D.pitch, pyr.slot_stride[l], taps + D.tab_x, taps + D.tab_y);
        ++launches;
    }
    return launches;
}
