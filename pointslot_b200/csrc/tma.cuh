// tma.cuh — minimal TMA (cp.async.bulk.tensor) + mbarrier wrappers for sm_100a, inline PTX.
// Tensor maps are created on the host (engine.cu, cuTensorMapEncodeTiled through
// cudaGetDriverEntryPoint) and passed to kernels as __grid_constant__ parameters.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "pslot_internal.h"

struct TmapSet {
    CUtensorMap lv[PSLOT_MAX_LEVELS]; // level l as a (x, y, slot) u8 tensor, box = FAST cell tile
};
// pyramid: level l as the SOURCE of level l + 1 (box = the source rectangle of one 128 x 64 output tile) and as a
// DESTINATION (box = one output tile)
struct PyrMaps {
    CUtensorMap src[PSLOT_MAX_LEVELS], dst[PSLOT_MAX_LEVELS];
    int box_w[PSLOT_MAX_LEVELS], box_h[PSLOT_MAX_LEVELS]; // source box of level l (16-byte multiple x rows)
};

namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}

// make the barrier initialisation visible to the async (TMA) proxy
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 3-D tiled load: box of the tensor map at element coordinates (c0, c1, c2) -> dense shared tile.
__device__ __forceinline__ void load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// 3-D tiled store: dense shared tile -> box of the tensor map at (c0, c1, c2); elements outside the tensor are not written.
// The generic-proxy writes that filled the tile must be ordered before it with fence_proxy_async().
__device__ __forceinline__ void store_3d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the committed bulk stores have finished READING shared memory (the CTA may then exit or reuse the tile)
__device__ __forceinline__ void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void prefetch_map(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

} // namespace tma
