// Which issue pipe do candidate min/max forms use on sm_100a?  Each op alone, then interleaved with VIMNMX3:
// interleaved time ~ max -> different pipes, ~ sum -> same pipe.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int OP, bool MIX> __global__ void k(unsigned *out, int iters)
{
    unsigned a[8], m[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x; m[i] = a[i] ^ 0x1234567u; }
    unsigned b = (threadIdx.x ^ 0x5bd1e995u) & 0x3fff3fffu, c = (blockIdx.x * 0x9E3779B9u + 77u) & 0x3fff3fffu;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) { __half2 x = *(__half2 *)&a[i], y = *(__half2 *)&b; x = __hmax2(x, y); a[i] = *(unsigned *)&x; }
            else if (OP == 1) { float x = __uint_as_float(a[i]), y = __uint_as_float(b); a[i] = __float_as_uint(fmaxf(x, y)); }
            else if (OP == 2) a[i] = a[i] * b + c;                       // IMAD
            else if (OP == 3) { __half2 x = *(__half2 *)&a[i], y = *(__half2 *)&b, z = *(__half2 *)&c; x = __hfma2(x, y, z); a[i] = *(unsigned *)&x; }
            else if (OP == 4) a[i] = __vabsdiffu2(a[i], b);
            else if (OP == 5) a[i] = __dp4a(a[i], b, c);                 // IDP4A
            if (MIX) m[i] = __vimax3_u16x2(m[i], b, c);
            a[i] ^= 0; 
        }
        b = (b + it) & 0x3fff3fffu; c ^= b;
    }
    unsigned s = 0;
    for (int i = 0; i < 8; ++i) s ^= a[i] ^ m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}
__global__ void kv(unsigned *out, int iters)
{
    unsigned m[8];
    for (int i = 0; i < 8; ++i) m[i] = threadIdx.x * 2654435761u + i;
    unsigned b = threadIdx.x ^ 0x5bd1e995u, c = blockIdx.x * 0x9E3779B9u + 77u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = __vimax3_u16x2(m[i], b, c);
        b += it; c ^= b;
    }
    unsigned s = 0;
    for (int i = 0; i < 8; ++i) s ^= m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}
template <typename F> float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(16); cudaEventRecord(e0); f(4096); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
    unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int B = 148 * 8, T = 256;
    printf("VIMNMX3 alone          %.3f ms\n", timeit([&](int n) { kv<<<B, T>>>(d, n); }));
#define RUN(OP, name) printf("%-10s alone %.3f ms   mixed with VIMNMX3 %.3f ms\n", name, timeit([&](int n) { k<OP, false><<<B, T>>>(d, n); }), timeit([&](int n) { k<OP, true><<<B, T>>>(d, n); }));
    RUN(0, "HMNMX2") RUN(1, "FMNMX") RUN(2, "IMAD") RUN(3, "HFMA2") RUN(4, "VABSDIFF2") RUN(5, "IDP4A")
    return 0;
}
