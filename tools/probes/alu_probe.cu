// Issue-rate probe for the integer instructions the FAST kernel leans on (sm_100a).
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
            if (OP == 0) a[i] = __vimax3_u16x2(a[i], b, c);
            else if (OP == 1) a[i] = __vmaxu2(a[i], b);
            else if (OP == 2) a[i] = __byte_perm(a[i], b, 0x5140);
            else if (OP == 3) a[i] = (a[i] & b) ^ c;             // LOP3
            else if (OP == 4) a[i] = a[i] * b + c;               // IMAD
            else if (OP == 5) a[i] = __funnelshift_r(a[i], b, 8); // SHF
            else if (OP == 6) a[i] = a[i] + b + c;               // IADD3
            else if (OP == 7) a[i] = max(a[i], b);               // IMNMX/VIMNMX 32
            else if (OP == 8) { a[i] = __vimax3_u16x2(a[i], b, c); b = b * a[(i + 1) & 7] + c; } // mix VIMNMX3 + IMAD
        }
        b += it; c ^= b;
    }
    unsigned s = 0;
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}
template <int OP> void run(const char *name, unsigned *d)
{
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(d, 16);
    cudaEventRecord(e0); k<OP><<<blocks, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double winst = (double)blocks * threads / 32 * iters * 8;
    // per SM sub-partition: 148*4 issue ports
    printf("%-22s %8.3f ms  %6.3f warp-inst/clk/SMSP (at 1.965 GHz)\n", name, ms, winst / (ms * 1e-3) / 1.965e9 / (148 * 4));
}
int main()
{
    unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("VIMNMX3.U16x2", d); run<1>("VIMNMX.U16x2", d); run<2>("PRMT", d); run<3>("LOP3", d); run<4>("IMAD", d);
    run<5>("SHF", d); run<6>("IADD3", d); run<7>("IMNMX.U32", d); run<8>("VIMNMX3 + IMAD mix", d);
    return 0;
}
