// ingest.cu — colour -> gray ingest: the cv::cvtColor(RGB/BGR(A) -> GRAY) calls that
// Tracking::GrabImageStereo makes right before building the Frame (reference src/Tracking.cc:1014-1039),
// moved onto the device so that colour KITTI frames (Camera.gray: 0) are uploaded once and converted
// straight into pyramid level 0.
//
// Arithmetic = OpenCV 4.13 RGB2Gray<uchar>: Y = (R*9798 + G*19235 + B*3735 + 2^14) >> 15 (BT.601, 15-bit
// fixed point; OpenCV 3.x used the 14-bit triple 4899/9617/1868, which differs in the last bit for ~0.25 %
// of colours — the oracle pins the 4.13 behaviour exhaustively over all 2^24 colours against cv2).
//
// Mapping: one thread per 4 output pixels; the 12 / 16 source bytes are fetched as aligned 32-bit words of
// the tightly packed staging image and re-aligned with funnel shifts.  Bound: HBM stream;
// algorithmic bytes = S * (channels + 1) * P_0.
#include "pslot_internal.h"

namespace {

__device__ __forceinline__ uint32_t gray_of(uint32_t c0, uint32_t c1, uint32_t c2, bool rgb)
{
    const uint32_t r = rgb ? c0 : c2, b = rgb ? c2 : c0;
    return (r * 9798u + c1 * 19235u + b * 3735u + (1u << 14)) >> 15;
}

// aligned 32-bit load that never touches a byte at or beyond `limit` (the end of the caller's buffer)
__device__ __forceinline__ uint32_t load_word_safe(const uint8_t *src, size_t a, size_t limit)
{
    if (a + 4 <= limit) return __ldg(reinterpret_cast<const uint32_t *>(src + a));
    uint32_t v = 0;
    for (int i = 0; i < 4; ++i)
        if (a + i < limit) v |= (uint32_t)__ldg(src + a + i) << (8 * i);
    return v;
}

template <int CH>
__global__ void __launch_bounds__(256)
k_gray_ingest(const uint8_t *__restrict__ src, size_t src_image_stride, size_t limit, uint8_t *__restrict__ dst, int w,
              int h, int dpitch, unsigned long long dstride, int rgb)
{
    const int gx = blockIdx.x * blockDim.x + threadIdx.x; // group of 4 pixels
    const int y = blockIdx.y;
    const int x0 = gx * 4;
    if (x0 >= w) return;
    // tight layout: row pitch = w * CH, images back to back; `src` itself is 256-byte aligned (cudaMalloc),
    // an image start need not be 4-byte aligned, so the word alignment is taken on the offset from `src`
    const size_t off = (size_t)blockIdx.z * src_image_stride + ((size_t)y * w + x0) * CH;
    const size_t a = off & ~(size_t)3;
    const int sh = (int)(off - a) * 8;
    // words covering the 4 pixels (the word after the last needed byte is only touched if it exists)
    uint32_t v[CH + 1];
#pragma unroll
    for (int i = 0; i <= CH; ++i) v[i] = load_word_safe(src, a + 4 * i, limit);
    uint32_t q[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) q[i] = __funnelshift_r(v[i], v[i + 1], sh);
    uint32_t g[4];
    if (CH == 3) { // bytes: p0 = q0[0..2], p1 = q0[3] q1[0..1], p2 = q1[2..3] q2[0], p3 = q2[1..3]
        g[0] = gray_of(q[0] & 0xFF, (q[0] >> 8) & 0xFF, (q[0] >> 16) & 0xFF, rgb);
        g[1] = gray_of(q[0] >> 24, q[1] & 0xFF, (q[1] >> 8) & 0xFF, rgb);
        g[2] = gray_of((q[1] >> 16) & 0xFF, q[1] >> 24, q[2] & 0xFF, rgb);
        g[3] = gray_of((q[2] >> 8) & 0xFF, (q[2] >> 16) & 0xFF, q[2] >> 24, rgb);
    } else if (CH == 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) g[i] = gray_of(q[i] & 0xFF, (q[i] >> 8) & 0xFF, (q[i] >> 16) & 0xFF, rgb);
    } else { // CH == 1: gray input, plain re-pitch of tightly packed rows into the padded level-0 rows
#pragma unroll
        for (int i = 0; i < 4; ++i) g[i] = (q[0] >> (8 * i)) & 0xFF;
    }
    uint8_t *o = dst + (size_t)blockIdx.z * dstride + (size_t)y * dpitch + x0;
    if (x0 + 3 < w) {
        *reinterpret_cast<uint32_t *>(o) = g[0] | (g[1] << 8) | (g[2] << 16) | (g[3] << 24); // dpitch % 128 == 0
    } else {
        for (int i = 0; x0 + i < w; ++i) o[i] = (uint8_t)g[i];
    }
}

// Gray input: plain re-pitch of tightly packed rows into the padded level-0 rows.  One warp per row: a lane fetches the two
// aligned 16-byte chunks that hold its 16 output pixels, the byte offset of the row inside its chunk is the same for the
// whole row (warp-uniform switch over its word part, funnel shifts for the rest), one 16-byte store per lane into the
// 128-byte aligned destination row.  ~15 instructions per 16 bytes: an HBM stream (the first version - five guarded word
// loads and an integer division per thread - was issue-bound at 84 % and half the HBM rate).
constexpr int RW_WARPS = 4; // rows per CTA
__global__ void __launch_bounds__(RW_WARPS * 32)
k_gray_repitch(const uint8_t *__restrict__ src, size_t src_image_stride, size_t limit, uint8_t *__restrict__ dst, int w,
               int h, int dpitch, unsigned long long dstride)
{
    const int lane = threadIdx.x & 31;
    const int y = blockIdx.x * RW_WARPS + (threadIdx.x >> 5);
    if (y >= h) return;
    const uint8_t *p = src + (size_t)blockIdx.z * src_image_stride + (size_t)y * w; // the row's first pixel, any alignment
    const uint8_t *end = src + limit;
    const int s = (int)((uintptr_t)p & 15); // the row starts s bytes into an aligned 16-byte chunk
    const uint8_t *pa = p - s;
    const int sh = 8 * (s & 3);
    uint8_t *o = dst + (size_t)blockIdx.z * dstride + (size_t)y * dpitch;
    const int groups = (w + 15) >> 4;
    if (pa < src || pa + 16 * (size_t)(groups + 1) > end) {
        // the first / last row of the caller's buffer: the aligned chunks would reach outside [src, src + limit)
        for (int x = lane; x < w; x += 32) o[x] = __ldg(p + x);
        return;
    }
    for (int g = lane; g < groups; g += 32) {
        const uint4 *pc = reinterpret_cast<const uint4 *>(pa + 16 * (size_t)g);
        const uint4 c0 = __ldg(pc), c1 = __ldg(pc + 1);
        const uint32_t v[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        uint4 q;
        switch (s >> 2) { // warp-uniform
        case 0: q.x = __funnelshift_r(v[0], v[1], sh); q.y = __funnelshift_r(v[1], v[2], sh); q.z = __funnelshift_r(v[2], v[3], sh); q.w = __funnelshift_r(v[3], v[4], sh); break;
        case 1: q.x = __funnelshift_r(v[1], v[2], sh); q.y = __funnelshift_r(v[2], v[3], sh); q.z = __funnelshift_r(v[3], v[4], sh); q.w = __funnelshift_r(v[4], v[5], sh); break;
        case 2: q.x = __funnelshift_r(v[2], v[3], sh); q.y = __funnelshift_r(v[3], v[4], sh); q.z = __funnelshift_r(v[4], v[5], sh); q.w = __funnelshift_r(v[5], v[6], sh); break;
        default: q.x = __funnelshift_r(v[3], v[4], sh); q.y = __funnelshift_r(v[4], v[5], sh); q.z = __funnelshift_r(v[5], v[6], sh); q.w = __funnelshift_r(v[6], v[7], sh); break;
        }
        const int x0 = 16 * g;
        if (x0 + 15 < w) {
            *reinterpret_cast<uint4 *>(o + x0) = q;
        } else { // row tail: do not spill the next row's pixels into the padding
            const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
            for (int i = 0; x0 + i < w; ++i) o[x0 + i] = (uint8_t)(qq[i >> 2] >> (8 * (i & 3)));
        }
    }
}

} // namespace

int launch_gray_ingest(const uint8_t *src, size_t src_image_stride, int channels, int rgb, uint8_t *dst, int w, int h,
                       int dpitch, unsigned long long dstride, int nimages, cudaStream_t st)
{
    dim3 block(256);
    dim3 grid(((w + 3) / 4 + 255) / 256, h, nimages);
    const size_t limit = (size_t)(nimages - 1) * src_image_stride + (size_t)w * h * channels; // end of the last image
    if (channels == 1)
        k_gray_repitch<<<dim3((h + RW_WARPS - 1) / RW_WARPS, 1, nimages), RW_WARPS * 32, 0, st>>>(src, src_image_stride, limit, dst, w,
                                                                                                 h, dpitch, dstride);
    else if (channels == 3)
        k_gray_ingest<3><<<grid, block, 0, st>>>(src, src_image_stride, limit, dst, w, h, dpitch, dstride, rgb);
    else
        k_gray_ingest<4><<<grid, block, 0, st>>>(src, src_image_stride, limit, dst, w, h, dpitch, dstride, rgb);
    return 1;
}
