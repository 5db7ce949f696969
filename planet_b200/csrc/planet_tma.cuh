// planet_tma.cuh -- bulk asynchronous copies shared memory -> global memory (the TMA unit's
// non-tensor form, PTX cp.async.bulk, SASS UBLKCP) and the fences / group waits around them.
//
// Used where a kernel produces a contiguous run of finished output on chip -- K1's rebased strip
// of one quad (8 144 B), K2's 128-sample tile of a height map (512 B) -- and wants it to leave the
// SM as ONE request instead of a store instruction per thread.  The destination is any global
// address: local HBM, or a peer GPU's memory mapped through CUDA IPC (the stores then travel over
// NVLink), which is how K2 gathers finished height maps into every GPU's buffer while it computes.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace planet {
namespace tma {

// Writes made to shared memory by ordinary st.shared must be made visible to the async proxy
// before a bulk copy reads them: every writing thread fences, then the CTA / warp synchronises,
// then ONE thread issues the copy.
__device__ __forceinline__ void fence_smem_writes()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// dst: 16-byte aligned global address; src: 16-byte aligned shared address; bytes % 16 == 0
__device__ __forceinline__ void store_bulk(void *dst, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(__cvta_generic_to_global(dst)), "r"((uint32_t)__cvta_generic_to_shared(src)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

// all but the latest N groups of this thread have finished READING shared memory (the staging
// buffer may be overwritten)
template <int N> __device__ __forceinline__ void wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
// all but the latest N groups of this thread are complete (their writes are performed)
template <int N> __device__ __forceinline__ void wait_all() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }

// ---- global -> shared bulk loads, completion on an mbarrier (for kernels that only forward data) ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t phase)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(phase) : "memory");
    return ok != 0;
}
// src: 16-byte aligned global address; dst: 16-byte aligned shared address; bytes % 16 == 0
__device__ __forceinline__ void load_bulk(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes),
                    "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// data another thread wrote through ordinary stores (and this thread has acquired) is about to be
// read through the async proxy
__device__ __forceinline__ void fence_before_async_reads() { asm volatile("fence.proxy.async;" ::: "memory"); }

} // namespace tma
} // namespace planet
