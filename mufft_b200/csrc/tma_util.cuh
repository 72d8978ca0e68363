// Bulk-copy (TMA) and mbarrier helpers shared by fft2d_tma_kernel.cuh, rows_prefetch_kernel.cuh and long_kernel.cuh:
// thin wrappers over the PTX instructions, plus a wait that cannot hang the device.
#pragma once

#include "fft_engine.cuh"

#ifndef MUFFT_EMU
#include <cuda.h>
#include <cstdio>

// Fraction of the lines stored with the evict_last policy that really get it (the rest: normal).  The L2 keeps evict_last
// lines in a set-aside of ~24 MB (tools/l2_ceiling_probe.cu): an intermediate larger than that may be better off with
// only a part of it pinned.
#ifndef MUFFT_EVICT_LAST_FRACTION
#define MUFFT_EVICT_LAST_FRACTION 1.0
#endif
#define MUFFT_STR_(x) #x
#define MUFFT_STR(x) MUFFT_STR_(x)

namespace mufft {

// A wait that sees no progress for this long ends in a trap instead of hanging the device: a protocol error, never a slow
// memory system (waits are microseconds).  Long enough to survive a preempted context (time slicing, a debugger).
#ifndef MUFFT_TMA_WATCHDOG_NS
#define MUFFT_TMA_WATCHDOG_NS 20000000000ull
#endif
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(void *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// A protocol error must not hang the device: MUFFT_TMA_WATCHDOG_NS without progress end the wait (false).
// SLEEP: nanoseconds between polls (0 for the compute warps).
template <int SLEEP>
__device__ __forceinline__ bool mbar_wait(void *bar, unsigned parity)
{
    if (mbar_try_wait(bar, parity)) return true;
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if constexpr (SLEEP > 0) __nanosleep(SLEEP);
        if ((++spins & (SLEEP > 0 ? 1023u : 65535u)) == 0 && global_ns() - t0 > MUFFT_TMA_WATCHDOG_NS) return false;
    }
    return true;
}
#ifdef MUFFT_TMA_DEBUG
#define MUFFT_TMA_WAIT(SLEEP, bar, parity, what, p0_, p1_) do { if (!mbar_wait<SLEEP>(bar, parity)) { if ((threadIdx.x & 31) == 0) printf("cta %d tid %d: timeout %s b=%u use=%u\n", (int)blockIdx.x, (int)threadIdx.x, what, (unsigned)(p0_), (unsigned)(p1_)); return; } } while (0)
#else
#define MUFFT_TMA_WAIT(SLEEP, bar, parity, what, p0_, p1_) do { if (!mbar_wait<SLEEP>(bar, parity)) __trap(); } while (0)
#endif

template <int KIND>
__device__ __forceinline__ unsigned long long l2_policy()
{
    unsigned long long pol = 0;
    if constexpr (KIND == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if constexpr (KIND == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, " MUFFT_STR(MUFFT_EVICT_LAST_FRACTION) ";" : "=l"(pol));
    return pol;
}
template <int KIND>
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, unsigned bytes, void *bar, unsigned long long pol)
{
    if constexpr (KIND == 0)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
template <int KIND>
__device__ __forceinline__ void bulk_store_1d(void *dst, const void *src, unsigned bytes, unsigned long long pol)
{
    if constexpr (KIND == 0)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
    else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes), "l"(pol) : "memory");
}
template <int KIND>
__device__ __forceinline__ void tensor_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, void *bar, unsigned long long pol)
{
    if constexpr (KIND == 0)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                     ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
template <int KIND>
__device__ __forceinline__ void tensor_store_2d(const CUtensorMap *map, int c0, int c1, const void *src, unsigned long long pol)
{
    if constexpr (KIND == 0)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                     ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;"
                     ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(src)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
} // namespace mufft
#endif // MUFFT_EMU
