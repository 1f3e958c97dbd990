// 2-input vs 3-input packed min/max issue rate on sm_100a, with operand patterns the compiler cannot merge.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(unsigned *out, int iters)
{
    unsigned a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
    unsigned b = threadIdx.x ^ 0x5bd1e995u, c = blockIdx.x * 0x9E3779B9u + 77u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = __vminu2(__vmaxu2(a[i], b), c);                       // 2 x 2-input (max then min)
            else if (OP == 1) a[i] = __vimin3_u16x2(__vimax3_u16x2(a[i], b, c), b, c); // 2 x 3-input
            else if (OP == 2) a[i] = min(max(a[i], b), c);                             // 2 x 32-bit 2-input
            else if (OP == 3) { a[i] = __vmaxu2(a[i], a[(i + 1) & 7]); a[i] = __vminu2(a[i], a[(i + 3) & 7]); } // register-register
        }
        b += 0x00010001u; c -= 0x00010001u;
    }
    unsigned s = 0;
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b + c;
}
template <int OP> void run(const char *name, unsigned *d)
{
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(d, 16);
    cudaEventRecord(e0); k<OP><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads / 32 * iters * 16; // 16 min/max instructions per iteration
    printf("%-34s %8.3f ms  %5.2f cycles per warp-instruction per SMSP (1.965 GHz)\n", name, ms, ms * 1e-3 * 1.965e9 * 148 * 4 / ops);
}
int main()
{
    unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("VIMNMX.U16x2 2-input max,min", d); run<1>("VIMNMX3.U16x2 3-input", d); run<2>("IMNMX 32-bit 2-input", d); run<3>("VIMNMX.U16x2 reg-reg", d);
    return 0;
}
