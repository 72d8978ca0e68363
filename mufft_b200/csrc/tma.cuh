// Bulk asynchronous copies (TMA) and the mbarrier they complete on, as small inline functions:
// the PTX in the CUDA build, a synchronous stand-in under tests/emu (MUFFT_EMU) so that kernels
// built on them run in the no-GPU test suite too.
//
//   tma_load_1d   cp.async.bulk            contiguous bytes, global -> shared
//   tma_load_2d   cp.async.bulk.tensor.2d  a box of a row-major 2D array described by a tensor map
//                                          (host side: cuTensorMapEncodeTiled, fft2d_inst.cu)
// Both count the bytes they deliver on an mbarrier in shared memory; consumers wait on its phase.
#pragma once

#include "fft_engine.cuh"

#ifndef MUFFT_EMU
#include <cuda.h> // CUtensorMap
#endif

namespace mufft {

#ifdef MUFFT_EMU
// Emulated tensor map of a row-major 2D array of 8-byte elements.
struct alignas(64) tensor_map_t {
    const unsigned char *base;
    size_t pitch_bytes;     // between consecutive rows
    unsigned box_w, box_h;  // box size in elements / rows
};
#define MUFFT_GRID_CONSTANT
// Emulated mbarrier word: low 32 bits = completed phases, high 32 bits = bytes still expected.
// Only the issuing thread touches the high half; copies are synchronous, so a phase completes
// inside the last tma_load_* of its expect_tx.
__device__ __forceinline__ void tma_mbar_init(uint64_t *bar, uint32_t) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
// (bytes may be completed by other threads BEFORE the expectation is posted - remote st.async
// senders of rowsc_kernel.cuh -: whichever operation brings the pending count to zero ends the phase)
__device__ __forceinline__ void tma_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    const uint64_t now = __atomic_add_fetch(bar, (uint64_t)bytes << 32, __ATOMIC_SEQ_CST);
    if ((now >> 32) == 0) __atomic_fetch_add(bar, 1ull, __ATOMIC_SEQ_CST);
}
__device__ __forceinline__ void emu_tx_done(uint64_t *bar, uint32_t bytes)
{
    const uint64_t left = __atomic_sub_fetch(bar, (uint64_t)bytes << 32, __ATOMIC_SEQ_CST);
    if ((left >> 32) == 0) __atomic_fetch_add(bar, 1ull, __ATOMIC_SEQ_CST);
}
__device__ __forceinline__ void tma_mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (((uint32_t)__atomic_load_n(bar, __ATOMIC_SEQ_CST) & 1u) == parity) std::this_thread::yield();
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    memcpy(dst, src, bytes);
    emu_tx_done(bar, bytes);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const tensor_map_t *map, int x, int y, uint64_t *bar)
{
    for (unsigned r = 0; r < map->box_h; r++)
        memcpy(static_cast<unsigned char *>(dst) + (size_t)r * map->box_w * 8, map->base + (size_t)(y + r) * map->pitch_bytes + (size_t)x * 8, (size_t)map->box_w * 8);
    emu_tx_done(bar, map->box_w * map->box_h * 8);
}
__device__ __forceinline__ void tma_fence_mbar_init() {}
__device__ __forceinline__ void tma_fence_reads_before_refill() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
__device__ __forceinline__ void tma_fence_generic_to_async() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

#else // ---- CUDA ------------------------------------------------------------------------------------

typedef CUtensorMap tensor_map_t;
#define MUFFT_GRID_CONSTANT __grid_constant__

__device__ __forceinline__ uint32_t tma_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(count));
}
// Makes the initialised mbarrier visible to the async proxy before the first copy names it.
__device__ __forceinline__ void tma_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// One arrival + `bytes` expected from the copies that follow.
__device__ __forceinline__ void tma_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tma_smem_u32(bar)), "r"(bytes) : "memory");
}
// Returns once the phase with parity `parity` has completed.
__device__ __forceinline__ void tma_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(tma_smem_u32(bar)), "r"(parity) : "memory");
}
// 16-byte aligned addresses, size a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tma_smem_u32(dst)), "l"(src), "r"(bytes), "r"(tma_smem_u32(bar)) : "memory");
}
// Box at element column x, row y of the array behind `map` (a __grid_constant__ kernel parameter).
__device__ __forceinline__ void tma_load_2d(void *dst, const tensor_map_t *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(tma_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(tma_smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
// This thread's generic-proxy reads of a landing buffer, ordered before the async-proxy write that
// refills it (issued by another thread after a barrier).
__device__ __forceinline__ void tma_fence_reads_before_refill() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Generic-proxy writes made visible to this thread (its own, or other SMs' after an acquire),
// ordered before the async-proxy reads of the copies it issues next.
__device__ __forceinline__ void tma_fence_generic_to_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

#endif

} // namespace mufft
