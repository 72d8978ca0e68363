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
// After mbarrier.init: make the barriers visible to the async proxy.
__device__ __forceinline__ void mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async_all();
}
} // namespace mufft

#else // MUFFT_EMU: the same interface on the CPU emulation of tests/emu (TEST INFRASTRUCTURE: index maps and protocols only)
#include <chrono>
#include <cstdio>

// What cuTensorMapEncodeTiled describes, as far as the kernels use it: a 2D array of floats, boxes of box[0] x box[1].
struct alignas(64) CUtensorMap {
    void *base;
    unsigned long long dims[2];   // floats per row that exist, rows
    unsigned long long row_bytes; // pitch
    unsigned box[2];
    unsigned char pad[128 - 8 - 16 - 8 - 8];
};

namespace mufft {

#ifndef MUFFT_TMA_WATCHDOG_NS
#define MUFFT_TMA_WATCHDOG_NS 60000000000ull
#endif
#ifndef MUFFT_EVICT_LAST_FRACTION
#define MUFFT_EVICT_LAST_FRACTION 1.0
#endif
// An mbarrier in its 8 bytes: [63:52] arrival count, [51:40] arrivals pending in this phase, [39:32] phase, [31:0] transaction
// bytes still expected (signed: a copy may complete before the expect_tx that announces it).
namespace emu_mbar {
inline unsigned long long pack(unsigned count, unsigned pending, unsigned phase, int tx)
{
    return ((unsigned long long)count << 52) | ((unsigned long long)(pending & 0xfff) << 40) | ((unsigned long long)(phase & 0xff) << 32) | (unsigned)tx;
}
inline void update(void *bar, int arrivals, int tx_delta)
{
    unsigned long long *w = static_cast<unsigned long long *>(bar);
    unsigned long long old = __atomic_load_n(w, __ATOMIC_SEQ_CST), now;
    do {
        const unsigned count = (unsigned)(old >> 52), phase = (unsigned)(old >> 32) & 0xff;
        unsigned pending = (unsigned)(old >> 40) & 0xfff;
        int tx = (int)(unsigned)old + tx_delta;
        pending -= (unsigned)arrivals;
        if (pending == 0 && tx == 0) now = pack(count, count, phase + 1, 0); // the phase completes
        else now = pack(count, pending, phase, tx);
    } while (!__atomic_compare_exchange_n(w, &old, now, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
}
} // namespace emu_mbar
inline void mbar_init(void *bar, unsigned count) { __atomic_store_n(static_cast<unsigned long long *>(bar), emu_mbar::pack(count, count, 0, 0), __ATOMIC_SEQ_CST); }
inline void mbar_init_fence() {}
inline void mbar_arrive(void *bar) { emu_mbar::update(bar, 1, 0); }
inline void mbar_arrive_expect_tx(void *bar, unsigned bytes) { emu_mbar::update(bar, 1, (int)bytes); }
inline bool mbar_try_wait(void *bar, unsigned parity)
{
    const unsigned phase = (unsigned)(__atomic_load_n(static_cast<unsigned long long *>(bar), __ATOMIC_SEQ_CST) >> 32) & 0xff;
    if ((phase & 1) != parity) return true;
    std::this_thread::yield();
    return false;
}
inline unsigned long long global_ns()
{
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
template <int SLEEP>
inline bool mbar_wait(void *bar, unsigned parity)
{
    if (mbar_try_wait(bar, parity)) return true;
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity))
        if ((++spins & 1023u) == 0 && global_ns() - t0 > MUFFT_TMA_WATCHDOG_NS) return false;
    return true;
}
#define MUFFT_TMA_WAIT(SLEEP, bar, parity, what, p0_, p1_) do { if (!mbar_wait<SLEEP>(bar, parity)) { fprintf(stderr, "emu: cta %u tid %u: timeout %s %u %u\n", blockIdx.x, threadIdx.x, what, (unsigned)(p0_), (unsigned)(p1_)); __trap(); } } while (0)
template <int KIND> inline unsigned long long l2_policy() { return 0; }
// The copies happen at once (the emulation has no asynchrony to offer); completion is signalled as the TMA unit would.
template <int KIND> inline void bulk_load_1d(void *dst, const void *src, unsigned bytes, void *bar, unsigned long long)
{
    std::memcpy(dst, src, bytes);
    emu_mbar::update(bar, 0, -(int)bytes);
}
template <int KIND> inline void bulk_store_1d(void *dst, const void *src, unsigned bytes, unsigned long long) { std::memcpy(dst, src, bytes); }
template <int KIND> inline void tensor_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, void *bar, unsigned long long)
{
    float *d = static_cast<float *>(dst);
    for (unsigned r = 0; r < map->box[1]; r++)
        for (unsigned c = 0; c < map->box[0]; c++) {
            const unsigned long long x = (unsigned long long)c0 + c, y = (unsigned long long)c1 + r;
            const bool in = x < map->dims[0] && y < map->dims[1]; // out of bounds: zero fill
            d[r * map->box[0] + c] = in ? *reinterpret_cast<const float *>(static_cast<const char *>(map->base) + y * map->row_bytes + x * 4) : 0.0f;
        }
    emu_mbar::update(bar, 0, -(int)(map->box[0] * map->box[1] * 4));
}
template <int KIND> inline void tensor_store_2d(const CUtensorMap *map, int c0, int c1, const void *src, unsigned long long)
{
    const float *s = static_cast<const float *>(src);
    for (unsigned r = 0; r < map->box[1]; r++)
        for (unsigned c = 0; c < map->box[0]; c++) {
            const unsigned long long x = (unsigned long long)c0 + c, y = (unsigned long long)c1 + r;
            if (x < map->dims[0] && y < map->dims[1]) // out of bounds: dropped
                *reinterpret_cast<float *>(static_cast<char *>(map->base) + y * map->row_bytes + x * 4) = s[r * map->box[0] + c];
        }
}
inline void bulk_commit() {}
inline void bulk_wait_read() {}
inline void bulk_wait_all() {}
inline void fence_proxy_async_smem() {}
inline void fence_proxy_async_all() {}
} // namespace mufft
#endif // MUFFT_EMU
