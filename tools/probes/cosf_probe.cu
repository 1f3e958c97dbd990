// cosf_probe.cu — does the device's cosf / sinf equal the host libm's for every argument cv::ORB can produce?
// computeOrbDescriptors (OpenCV orb.cpp) steers the BRIEF pattern with a = cosf(angle * (float)(CV_PI / 180.f)),
// b = sinf(same) for angle in [0, 360); this sweeps every float32 angle in that range (1.14e9 values) and counts the
// arguments for which the two implementations differ (next row).  Not part of the product.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fopenmp -o cosf_probe cosf_probe.cu && ./cosf_probe
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

__global__ void k(uint32_t first, uint32_t n, float fac, float *c, float *s, int mode)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = __uint_as_float(first + i) * fac;
    if (mode == 0) { c[i] = cosf(a); s[i] = sinf(a); }
    else { c[i] = (float)cos((double)a); s[i] = (float)sin((double)a); } // double evaluation, rounded once more
}

int main()
{
    const float fac = (float)(3.1415926535897932384626433832795 / 180.f);
    uint32_t last;
    {
        float f = 360.0f;
        memcpy(&last, &f, 4);
    }
    const uint32_t CH = 1u << 24;
    float *dc, *ds;
    cudaMalloc(&dc, 4ull * CH);
    cudaMalloc(&ds, 4ull * CH);
    std::vector<float> hc(CH), hs(CH);
  for (int mode = 0; mode < 2; ++mode) {
    unsigned long long total = 0, badc = 0, bads = 0;
    uint32_t worst = 0;
    for (uint64_t first = 0; first <= last; first += CH) {
        const uint32_t n = (uint32_t)std::min<uint64_t>(CH, (uint64_t)last + 1 - first);
        k<<<(n + 255) / 256, 256>>>((uint32_t)first, n, fac, dc, ds, mode);
        cudaMemcpy(hc.data(), dc, 4ull * n, cudaMemcpyDeviceToHost);
        cudaMemcpy(hs.data(), ds, 4ull * n, cudaMemcpyDeviceToHost);
        unsigned long long bc = 0, bs = 0;
#pragma omp parallel for reduction(+ : bc, bs)
        for (long long i = 0; i < (long long)n; ++i) {
            uint32_t u = (uint32_t)first + (uint32_t)i;
            float ang;
            memcpy(&ang, &u, 4);
            const float a = ang * fac;
            const float c = cosf(a), s = sinf(a);
            bc += memcmp(&c, &hc[i], 4) != 0;
            bs += memcmp(&s, &hs[i], 4) != 0;
        }
        if (bc + bs && !worst) worst = (uint32_t)first;
        badc += bc; bads += bs; total += n;
    }
    printf("angles %llu: cos mismatches %llu, sin mismatches %llu (%s vs host libm cosf/sinf)\n", total, badc, bads,
           mode == 0 ? "device cosf/sinf" : "device (float)cos((double)x)");
  }
    return 0;
}
