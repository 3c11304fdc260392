// ehb_tma.cuh -- Blackwell / Hopper bulk-asynchronous data movement (TMA) used by the rasterizer, as thin PTX wrappers.
//
//   tensor maps  : masks are a [items][H][W] f32 tensor; a 32 x 32 tile of one item is ONE instruction
//                  (cp.async.bulk.tensor.3d, SASS UTMASTG): the empty tiles of a frame (85 % of it) are zero-filled from a
//                  4 KB shared-memory tile, composed tiles are written from their staging tile.  Out-of-image parts of a
//                  tile (the top row of tiles of a 720-row image, partial columns) are clipped by the hardware.
//   bulk copies  : contiguous blocks (parked 4.3 KB batch blocks, 128-B triangle records, rows of a depth-plane window) go
//                  global -> shared with cp.async.bulk (SASS UBLKCP) and complete on an mbarrier: one instruction per block
//                  instead of one load + one store per 4 bytes, nothing staged in registers.
// Requirements of the hardware path: 16-byte aligned addresses and sizes; a tensor map needs 16-byte aligned row pitch
// (W % 4 == 0 for f32).  Callers fall back to plain loads / stores when that does not hold.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t ehb_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ehb_mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ehb_smem_addr(bar)), "r"(count) : "memory");
}
// makes mbarrier initialisation visible to the asynchronous proxy (the copy engines)
__device__ __forceinline__ void ehb_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders generic-proxy writes to shared memory before later asynchronous-proxy reads of it (TMA stores)
__device__ __forceinline__ void ehb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void ehb_mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ehb_smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void ehb_mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "EHB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra EHB_DONE;\n"
        "bra EHB_WAIT;\n"
        "EHB_DONE:\n"
        "}" ::"r"(ehb_smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// global -> shared, `bytes` (multiple of 16, both addresses 16-byte aligned), completes `bytes` on the mbarrier
__device__ __forceinline__ void ehb_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ehb_smem_addr(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(ehb_smem_addr(bar))
                 : "memory");
}

// shared -> global tile store through a 3-D tensor map: coordinates (x, y, z) of the tile's first element
__device__ __forceinline__ void ehb_tma_store_3d(const CUtensorMap* map, const void* smem_src, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                 "r"(ehb_smem_addr(smem_src)), "r"(x), "r"(y), "r"(z)
                 : "memory");
}
__device__ __forceinline__ void ehb_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores of this thread have finished READING their shared-memory source
__device__ __forceinline__ void ehb_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void ehb_prefetch_tensormap(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
