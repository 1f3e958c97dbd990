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

// Gray input: plain re-pitch of tightly packed rows into the padded level-0 rows, 16 pixels per thread (five
// aligned source words re-aligned with funnel shifts, one 16-byte store: destination rows are 128-byte aligned).
__global__ void __launch_bounds__(128)
k_gray_repitch(const uint8_t *__restrict__ src, size_t src_image_stride, size_t limit, uint8_t *__restrict__ dst, int w,
               int h, int dpitch, unsigned long long dstride)
{
    // threads are flattened over (row, group of 16 pixels) so that no lane idles on the ragged row end
    const int groups = (w + 15) >> 4;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= groups * h) return;
    const int y = idx / groups, x0 = (idx - y * groups) * 16;
    const size_t off = (size_t)blockIdx.z * src_image_stride + (size_t)y * w + x0;
    const size_t a = off & ~(size_t)3;
    const int sh = (int)(off - a) * 8;
    uint32_t v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = load_word_safe(src, a + 4 * i, limit);
    uint4 q;
    q.x = __funnelshift_r(v[0], v[1], sh); q.y = __funnelshift_r(v[1], v[2], sh);
    q.z = __funnelshift_r(v[2], v[3], sh); q.w = __funnelshift_r(v[3], v[4], sh);
    uint8_t *o = dst + (size_t)blockIdx.z * dstride + (size_t)y * dpitch + x0;
    if (x0 + 15 < w) {
        *reinterpret_cast<uint4 *>(o) = q;
    } else { // row tail: do not spill the next row's pixels into the padding
        const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
        for (int i = 0; x0 + i < w; ++i) o[i] = (uint8_t)(qq[i >> 2] >> (8 * (i & 3)));
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
        k_gray_repitch<<<dim3((((w + 15) / 16) * h + 127) / 128, 1, nimages), 128, 0, st>>>(src, src_image_stride, limit,
                                                                                          dst, w, h, dpitch, dstride);
    else if (channels == 3)
        k_gray_ingest<3><<<grid, block, 0, st>>>(src, src_image_stride, limit, dst, w, h, dpitch, dstride, rgb);
    else
        k_gray_ingest<4><<<grid, block, 0, st>>>(src, src_image_stride, limit, dst, w, h, dpitch, dstride, rgb);
    return 1;
}
