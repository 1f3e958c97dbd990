// Standalone TMA probe: 3-D u8 tensor (x, y, slot), box load at an unaligned x coordinate.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../pointslot_b200/csrc/tma.cuh"

struct Maps { CUtensorMap lv[4]; };

__global__ void probe(const __grid_constant__ Maps maps, int l, int bw, int bh, int cx, int cy, int cz, uint8_t* out)
{
    __shared__ __align__(128) uint8_t tile[80 * 66];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { tma::mbar_init(&bar, 1); tma::fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        tma::mbar_arrive_expect_tx(&bar, bw * bh);
        tma::load_3d(tile, &maps.lv[l], &bar, cx, cy, cz);
    }
    tma::mbar_wait(&bar, 0);
    __syncthreads();
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}

int main()
{
    const int W = 333, H = 222, S = 3, pitch = 384;
    std::vector<uint8_t> h((size_t)pitch * H * S);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 7 + i / pitch) & 0xFF);
    uint8_t *d, *dout;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 80 * 66);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeFn enc = (EncodeFn)fn;
    Maps maps;
    const int bw = 48, bh = 38;
    for (int l = 0; l < 4; ++l) {
        cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)S};
        cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * H};
        cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&maps.lv[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode l=%d -> %d\n", l, (int)r);
    }
    const int tests[][3] = {{16, 16, 0}, {15, 16, 1}, {301, 200, 2}, {46, 47, 1}};
    for (auto& t : tests) {
        probe<<<1, 128>>>(maps, 2, bw, bh, t[0], t[1], t[2], dout);
        cudaError_t e = cudaDeviceSynchronize();
        printf("coords (%d,%d,%d): %s\n", t[0], t[1], t[2], cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint8_t> o(bw * bh);
        cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
            int gx = t[0] + x, gy = t[1] + y;
            uint8_t exp = (gx < W && gy < H) ? h[(size_t)t[2] * pitch * H + (size_t)gy * pitch + gx] : 0;
            bad += o[y * bw + x] != exp;
        }
        printf("   mismatches %d\n", bad);
    }
    return 0;
}
