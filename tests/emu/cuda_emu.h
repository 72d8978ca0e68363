// TEST INFRASTRUCTURE ONLY - never linked into libmufft_b200.so, never on a product path.
//
// A small CPU emulation of the CUDA execution model, just large enough to run the kernel
// templates of mufft_b200/csrc/*.cuh unchanged under g++ (-DMUFFT_EMU): one OS thread per CUDA
// thread, real barriers for __syncthreads / named barriers / __syncwarp, per-CTA shared memory,
// thread-block clusters with distributed shared memory, global atomics.  tests/test_emu.py uses it
// to check the index maps, the barrier placement and the inter-CTA protocols of the kernels
// against the CPU oracle in the `-m "not gpu"` suite, where no GPU exists.
//
// OS threads are scheduled far more erratically than warps, so a missing barrier shows up here
// as a wrong result rather than passing by the luck of lock-step execution.  What it cannot
// check: performance, and anything that depends on the hardware memory model beyond
// sequentially consistent atomics.
#pragma once

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct alignas(8) float2 {
    float x, y;
};
static inline float2 make_float2(float x, float y) { return float2{ x, y }; }
struct alignas(16) uint4 {
    unsigned x, y, z, w;
};
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{ x, y, z, w }; }
#define __grid_constant__

namespace emu {

struct idx3 {
    unsigned x = 0, y = 0, z = 0;
};

// Reusable barrier for `count` participants (count is given at every call, as bar.sync does).
struct Barrier {
    std::mutex m;
    std::condition_variable cv;
    unsigned waiting = 0, generation = 0;
    // bar.arrive: counts as a participant of the current phase but does not wait for it to complete
    void arrive(unsigned count)
    {
        std::unique_lock<std::mutex> lock(m);
        if (++waiting == count) {
            waiting = 0;
            generation++;
            cv.notify_all();
        }
    }
    void sync(unsigned count)
    {
        std::unique_lock<std::mutex> lock(m);
        const unsigned g = generation;
        if (++waiting == count) {
            waiting = 0;
            generation++;
            cv.notify_all();
        } else {
            cv.wait(lock, [&] { return generation != g; });
        }
    }
};

struct Cluster;
struct Cta {
    idx3 block_idx;
    unsigned threads = 0;
    unsigned rank_in_cluster = 0;
    Cluster *cluster = nullptr;
    std::vector<unsigned char> dyn_smem;     // 128-byte aligned view in `smem`
    unsigned char *smem = nullptr;
    alignas(16) unsigned char static_smem[1024];
    Barrier named[16];                       // [0] = __syncthreads
    std::unique_ptr<Barrier[]> warp;         // __syncwarp
};
struct Cluster {
    std::vector<Cta *> ctas;
    Barrier barrier;
    unsigned threads_total = 0;
};

extern thread_local idx3 t_thread_idx, t_block_idx;
extern thread_local Cta *t_cta;
extern thread_local unsigned t_static_off;
extern idx3 g_block_dim, g_grid_dim;

inline unsigned char *dyn_smem() { return t_cta->smem; }
// Bump allocation inside the CTA's static shared area: every thread executes the same
// declarations in the same order, so every thread gets the same addresses.
template <typename T>
inline T *static_smem(unsigned count)
{
    unsigned off = (t_static_off + alignof(T) - 1) & ~(unsigned)(alignof(T) - 1);
    t_static_off = off + count * sizeof(T);
    return reinterpret_cast<T *>(t_cta->static_smem + off);
}

inline void bar_sync(unsigned id, unsigned count) { t_cta->named[id].sync(count); }
inline void bar_arrive(unsigned id, unsigned count) { t_cta->named[id].arrive(count); }
inline void sync_threads() { t_cta->named[0].sync(t_cta->threads); }
inline void sync_warp()
{
    const unsigned w = t_thread_idx.x / 32;
    const unsigned left = t_cta->threads - w * 32;
    t_cta->warp[w].sync(left < 32 ? left : 32);
}
// thread-block clusters
inline unsigned cluster_ctarank() { return t_cta->rank_in_cluster; }
inline unsigned cluster_nctarank() { return t_cta->cluster ? (unsigned)t_cta->cluster->ctas.size() : 1; }
inline void cluster_sync()
{
    if (t_cta->cluster) t_cta->cluster->barrier.sync(t_cta->cluster->threads_total);
    else sync_threads();
}
// address of the same dynamic-shared-memory offset in CTA `rank` of the cluster
template <typename T>
inline T *map_shared_rank(T *p, unsigned rank)
{
    Cta *peer = t_cta->cluster ? t_cta->cluster->ctas[rank] : t_cta;
    return reinterpret_cast<T *>(peer->smem + (reinterpret_cast<unsigned char *>(p) - t_cta->smem));
}

// Runs `kernel` for every CTA of a grid.  CTAs execute `concurrent` at a time (1 = one after the
// other; persistent kernels that talk to each other through global memory need all of them at once).
// cluster_size > 1 groups consecutive CTAs (along x) into clusters; a cluster always runs together.
void launch(idx3 grid, unsigned block_threads, size_t dyn_smem_bytes, unsigned concurrent, unsigned cluster_size,
            const std::function<void()> &kernel);

} // namespace emu

#define threadIdx emu::t_thread_idx
#define blockIdx emu::t_block_idx
#define blockDim emu::g_block_dim
#define gridDim emu::g_grid_dim

static inline void __syncthreads() { emu::sync_threads(); }
static inline void __syncwarp() { emu::sync_warp(); }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
[[noreturn]] static inline void __trap() { std::abort(); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T> static inline void __stcs(T *p, T v) { *p = v; }
template <typename T> static inline void __stcg(T *p, T v) { *p = v; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{ a.x + b.x, a.y + b.y }; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
