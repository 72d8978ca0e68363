// Row transforms of 32768 and 65536 points in ONE kernel on a thread-block cluster: 2 or 4 CTAs of
// 512 threads hold the transform in registers (32 values per thread, as the 16384-point shape) and
// exchange between passes through DISTRIBUTED shared memory.  The transform is read once and
// written once, where the two-kernel split (rowst_kernel.cuh) makes two passes over HBM.
//
// Exchange layout: position pos of the Stockham scatter belongs to the thread that will gather it,
// g = pos mod T (T = N/32 threads in the cluster), i.e. to CTA g / 512, at local element
// (pos / T) * 512 + g % 512.  Scatters are therefore remote stores (st.shared::cluster, fire and
// forget), gathers are local and conflict-free; a cluster barrier (release / acquire) sits between
// them, and one in front of every scatter keeps it from overtaking the peers' previous gather.
// Twiddles are derived (MUFFT_TW_DERIVE): a cluster barrier empties L1 like any cluster-scope fence,
// and the second-pass table alone is 254 KiB.
//
// Two flavours of the "data has landed" side of an exchange:
//   ASYNC = false  plain remote stores + a release / acquire cluster barrier (in SASS the release is a
//                  MEMBAR.ALL.GPU, which also waits for the global stores of the previous transform);
//   ASYNC = true   st.async remote stores that count their bytes on an mbarrier in the DESTINATION CTA
//                  (each CTA expects exactly its 128 KiB per exchange and waits on its own mbarrier);
//                  the cluster barrier in front of the scatter only has to keep writers behind the
//                  peers' previous gather and is a relaxed arrive.
//
// Replaces for these lengths, like rowst_kernel.cuh, the reference's chain of radix passes over the
// whole array (kernel.c:395-453 driven by fft.c:811-850).  Opt-in until measured
// (mufft_cuda_set_cluster_rows).
#pragma once

#include "abi.h"
#include "fft_engine.cuh"
#include "rows_kernel.cuh"
#include "tma.cuh"

namespace mufft {

#define MUFFT_ROWSC_THREADS 512
__host__ __device__ constexpr int rowsc_cluster(int n) { return n / (MUFFT_ROWSC_THREADS * 32); } // CTAs per transform
__host__ __device__ constexpr bool rowsc_supported(int n) { return engine_regs(n) == 32 && (rowsc_cluster(n) == 2 || rowsc_cluster(n) == 4); }
// local exchange area: 32 * 512 elements, one padding element per 32; then the receive mbarrier (ASYNC)
__host__ __device__ constexpr int rowsc_smem_elems() { return 32 * MUFFT_ROWSC_THREADS + MUFFT_ROWSC_THREADS + 2; }
__host__ __device__ constexpr size_t rowsc_smem_bytes() { return (size_t)rowsc_smem_elems() * sizeof(float2) + 16; }
#define MUFFT_ROWSC_EXCHANGE_BYTES (32u * MUFFT_ROWSC_THREADS * (unsigned)sizeof(float2)) // what a CTA receives per exchange

// ---- cluster primitives (PTX in the CUDA build, tests/emu stand-ins otherwise) ---------------------
__device__ __forceinline__ unsigned cluster_cta_rank()
{
#ifdef MUFFT_EMU
    return emu::cluster_ctarank();
#else
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
#endif
}
__device__ __forceinline__ unsigned cluster_id_x()
{
#ifdef MUFFT_EMU
    return blockIdx.x / emu::cluster_nctarank();
#else
    unsigned r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
#endif
}
__device__ __forceinline__ unsigned cluster_count_x()
{
#ifdef MUFFT_EMU
    return gridDim.x / emu::cluster_nctarank();
#else
    unsigned r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
#endif
}
// Every thread of every CTA of the cluster; orders shared-memory accesses (also remote ones) across it.
__device__ __forceinline__ void cluster_barrier()
{
#ifdef MUFFT_EMU
    emu::cluster_sync();
#else
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}
// Arrive without memory ordering (the accesses it separates are complete register-consumed loads), wait.
__device__ __forceinline__ void cluster_barrier_relaxed()
{
#ifdef MUFFT_EMU
    emu::cluster_sync();
#else
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
#endif
}
#ifdef MUFFT_EMU
typedef cpx *dsmem_ptr; // address of the exchange area in a CTA of the cluster
typedef uint64_t *dsmem_bar;
__device__ __forceinline__ dsmem_ptr dsmem_map(cpx *local, unsigned rank) { return emu::map_shared_rank(local, rank); }
__device__ __forceinline__ dsmem_bar dsmem_map_bar(uint64_t *local, unsigned rank) { return emu::map_shared_rank(local, rank); }
__device__ __forceinline__ void dsmem_store(dsmem_ptr base, int elem, cpx v) { base[elem] = v; }
__device__ __forceinline__ void dsmem_store_async(dsmem_ptr base, int elem, cpx v, dsmem_bar bar)
{
    base[elem] = v;
    emu_tx_done(bar, 8);
}
#else
typedef uint32_t dsmem_ptr; // shared::cluster window address
__device__ __forceinline__ dsmem_ptr dsmem_map(cpx *local, unsigned rank)
{
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"((uint32_t)__cvta_generic_to_shared(local)), "r"(rank));
    return out;
}
__device__ __forceinline__ void dsmem_store(dsmem_ptr base, int elem, cpx v)
{
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(base + (uint32_t)elem * 8u), "f"(v.x), "f"(v.y) : "memory");
}
typedef uint32_t dsmem_bar;
__device__ __forceinline__ dsmem_bar dsmem_map_bar(uint64_t *local, unsigned rank)
{
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"((uint32_t)__cvta_generic_to_shared(local)), "r"(rank));
    return out;
}
// 8 bytes to a peer's shared memory, counted on the mbarrier `bar` of that peer when they land
__device__ __forceinline__ void dsmem_store_async(dsmem_ptr base, int elem, cpx v, dsmem_bar bar)
{
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                 ::"r"(base + (uint32_t)elem * 8u), "f"(v.x), "f"(v.y), "r"(bar) : "memory");
}
#endif

// Scatter of a radix-32 pass for cluster thread g (one butterfly per thread): see the layout note at
// the top.  Output q goes to position pos = base + q*p, and `base` has zeros in the five bits q*p
// occupies, so every bit field of pos - owner CTA, row, thread inside the CTA, padding - is the
// thread's part OR a compile-time part of q: the thread computes one address and one owner, the
// 32 stores use immediate offsets.
template <int N, int PASS, int CL, bool ASYNC>
__device__ __forceinline__ void rowsc_scatter(const cpx (&v)[32], const dsmem_ptr (&area)[CL], const dsmem_bar (&bars)[CL], int g)
{
    constexpr int T = N / 32;
    constexpr int p = engine_span(N, PASS);
    constexpr int LGP = ilog2(p), LT = ilog2(T);
    static_assert(engine_regs(N) == 32 && engine_radix(N, PASS) == 32, "cluster scatter: one radix-32 butterfly per thread");
    const int k = g & (p - 1);
    const int base = (g - k) * 32 + k;
    const int own_base = (base & (T - 1)) / MUFFT_ROWSC_THREADS;
    const int elem_base = (base >> LT) * MUFFT_ROWSC_THREADS + (base & (MUFFT_ROWSC_THREADS - 1));
    const int pbase = pad_pos<5>(elem_base);
    // owner = own_base | (compile-time part of q): the areas this thread can hit, by that part
    dsmem_ptr sel[CL];
    dsmem_bar selb[CL];
#pragma unroll
    for (int j = 0; j < CL; j++) {
        sel[j] = area[0];
        selb[j] = bars[0];
#pragma unroll
        for (int c = 1; c < CL; c++) {
            sel[j] = (own_base | j) == c ? area[c] : sel[j];
            selb[j] = (own_base | j) == c ? bars[c] : selb[j];
        }
    }
    static_for<32>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        constexpr int QP = q << LGP;
        constexpr int OWN_Q = (QP & (T - 1)) / MUFFT_ROWSC_THREADS;
        constexpr int ELEM_Q = (QP >> LT) * MUFFT_ROWSC_THREADS + (QP & (MUFFT_ROWSC_THREADS - 1));
        if constexpr (ASYNC) dsmem_store_async(sel[OWN_Q], pbase + ELEM_Q + (ELEM_Q >> 5), v[q], selb[OWN_Q]);
        else dsmem_store(sel[OWN_Q], pbase + ELEM_Q + (ELEM_Q >> 5), v[q]);
    });
}

template <int N, int DIR, int IN_MODE, int OUT_MODE, bool ASYNC = false>
__global__ void __launch_bounds__(MUFFT_ROWSC_THREADS, 1) fft_rows_cluster_kernel(const mcu_rows_args a)
{
    static_assert(IN_MODE == MCU_IN_COMPLEX || IN_MODE == MCU_IN_HALF, "cluster rows: plain or zero-padded complex input");
    static_assert(OUT_MODE == MCU_OUT_COMPLEX || OUT_MODE == MCU_OUT_PEERS, "cluster rows: complex output");
    constexpr int R = 32;
    constexpr int T = N / R;
    constexpr int CL = rowsc_cluster(N);
    constexpr int NP = engine_num_passes(N);
    MUFFT_DYNAMIC_SMEM(smem_raw);
    cpx *sm = reinterpret_cast<cpx *>(smem_raw);

    const int tid = threadIdx.x;
    const unsigned rank = cluster_cta_rank();
    const int g = (int)rank * MUFFT_ROWSC_THREADS + tid; // thread index inside the transform
    const cpx *__restrict__ tw = static_cast<const cpx *>(a.tw);
    uint64_t *rx = reinterpret_cast<uint64_t *>(smem_raw + (size_t)rowsc_smem_elems() * sizeof(float2)); // receive mbarrier (ASYNC)
    dsmem_ptr area[CL];
    dsmem_bar bars[CL];
#pragma unroll
    for (int c = 0; c < CL; c++) {
        area[c] = dsmem_map(sm, (unsigned)c);
        bars[c] = dsmem_map_bar(rx, (unsigned)c);
    }
    uint32_t parity = 0;
    if constexpr (ASYNC) {
        if (tid == 0) {
            tma_mbar_init(rx, 1);
            tma_fence_mbar_init();
        }
        cluster_barrier(); // every mbarrier of the cluster is initialised before anyone sends
    }

    // every CTA of a cluster walks the same transforms: the barriers below stay matched
    for (size_t b = cluster_id_x(); b < a.batch; b += cluster_count_x()) {
        const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + b * a.in_stride;
        cpx v[R];
#pragma unroll
        for (int e = 0; e < R; e++) {
            if constexpr (IN_MODE == MCU_IN_HALF) v[e] = e < R / 2 ? in[g + e * T] : make_float2(0.0f, 0.0f);
            else v[e] = in[g + e * T];
        }
        static_for<NP>([&](auto pc) {
            constexpr int PASS = decltype(pc)::value;
            engine_butterflies<N, DIR, PASS, 0, -1, MUFFT_TW_DERIVE>(v, tw, g);
            if constexpr (PASS + 1 < NP) {
                if constexpr (ASYNC) {
                    cluster_barrier_relaxed(); // the peers have gathered what the previous exchange (or transform) left here
                    if (tid == 0) tma_mbar_expect_tx(rx, MUFFT_ROWSC_EXCHANGE_BYTES);
                    rowsc_scatter<N, PASS, CL, true>(v, area, bars, g);
                    tma_mbar_wait(rx, parity); // this CTA's 128 KiB have landed
                    parity ^= 1;
                } else {
                    cluster_barrier(); // the peers have gathered what the previous exchange (or transform) left here
                    rowsc_scatter<N, PASS, CL, false>(v, area, bars, g);
                    cluster_barrier(); // every remote store has landed
                }
#pragma unroll
                for (int e = 0; e < R; e++) v[e] = sm[pad_pos<5>(e * MUFFT_ROWSC_THREADS + tid)];
            }
        });
        if constexpr (OUT_MODE == MCU_OUT_PEERS) {
            const size_t row = (a.peers.row0 + b) * a.peers.pitch;
            const unsigned mask = (1u << a.peers.shift) - 1u;
#pragma unroll
            for (int e = 0; e < R; e++) {
                const unsigned x = (unsigned)(g + e * T);
                static_cast<cpx *>(a.peers.base[x >> a.peers.shift])[row + (x & mask)] = v[e];
            }
        } else {
            cpx *__restrict__ out = static_cast<cpx *>(a.out) + b * a.out_stride;
#pragma unroll
            for (int e = 0; e < R; e++) out[g + e * T] = v[e];
        }
    }
    cluster_barrier(); // no CTA leaves while a peer may still store into its shared memory
}

} // namespace mufft
