// tc_tma.cuh -- the Blackwell copy engine for the tiled kernels: 2-D tensor maps (host), mbarrier and
// cp.async.bulk.tensor wrappers (device).  A tile is moved by ONE instruction issued by one lane; its arrival is
// counted on an mbarrier; a staging buffer is handed back through a shared-memory counter and the lane whose warp
// hands it back last issues the copy of the tile after next -- no producer warp (a ninth warp would cost every
// warp 16 registers: the register file is split over the four schedulers) and nobody blocks on "empty".
#pragma once

#include <cuda.h> // CUtensorMap (types only: the encoder is looked up through cudaGetDriverEntryPoint)

#include "tc_common.cuh"

namespace tc {

// ---- host (tc_api.cu) ----
bool tma_available();
// 2-D map over a pitched array of `elem_bytes` (4 or 8) elements: dim0 elements per row, dim1 rows, box0 x box1
// elements per copy, zero fill outside.  Cached by every argument.  NB the copy itself needs the first byte of a box row
// 16-byte aligned (coordinate 0 x element size a multiple of 16): a violation surfaces as "illegal instruction".
int tensor_map_2d(const void* base, int elem_bytes, unsigned long long dim0, unsigned long long dim1, unsigned long long pitch_bytes,
                  unsigned box0, unsigned box1, CUtensorMap* out);

// ---- device ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
                     smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void tma_load_tile(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"((unsigned long long)map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// generic-proxy reads / writes of a buffer are ordered before the async-proxy writes of the copy that refills it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// returns the number of warps that had handed the buffer back before this one
__device__ __forceinline__ int release_count(int* counter)
{
    int prev;
    asm volatile("atom.acq_rel.cta.shared::cta.add.s32 %0, [%1], 1;\n" : "=r"(prev) : "r"(smem_u32(counter)) : "memory");
    return prev;
}

} // namespace tc
