// Row transforms of 32768 .. 65536 points (and, as an alternative shape, 16384) in ONE launch and
// ONE pass over HBM: the four-step split N = N1 * N2 with both steps inside a persistent CTA and the
// intermediate parked in an L2-resident scratch slot that belongs to the CTA.
//
//   x[n1*N2 + n2]  --A-->  Y[k1][n2] = w_N^(k1 n2) * sum_n1 x[n1*N2 + n2] w_N1^(n1 k1)
//                  --B-->  X[k1 + N1*k2] = sum_n2 Y[k1][n2] w_N2^(n2 k2)
//
//   step A  one thread per column n2: the N1 (16 or 32) inputs are N2 elements apart, so consecutive
//           lanes read consecutive addresses; the whole N1-point transform is one register butterfly
//           (no shared memory, no barrier).  The inter-stage twiddles w_N^(k1 n2), k1 < N1, are
//           powers of w_N^n2: log2(N1) entries of the master root table are loaded, the others are
//           products (TwTree of fft_engine.cuh).  Y goes to the CTA's scratch slot, coalesced.
//   step B  TR rows k1 at a time: N2-point row transforms on the shared-memory engine.  X has stride
//           N1 along k2, so the last engine pass changes the thread mapping from "lanes walk along
//           a row" to "lanes walk across the TR rows" (the exchange goes through shared memory
//           anyway): a warp then stores runs of TR consecutive k1 (64 or 128 bytes) per k2.
//
// The scratch slot (N complex per resident CTA, 38 MB for 32768 points on 148 SMs x 1 CTA, 76 MB x 2)
// is written and read back by the same CTA a few microseconds apart, so it lives in the 126 MB L2: HBM
// sees each transform's input once and its output once (16 N bytes), where the two-kernel split
// (cols_kernel.cuh + rowst_kernel.cuh) moves 32 N bytes and the reference's log8(N) = 5 passes move 80 N
// (ref fft.c:811-850 over kernel.c:395-453).  Two things keep it that way (ncu on the first version,
// profiles/r02b_long32768.txt: 2.03 GB written for 1.07 GB of output - every scratch line was eventually
// written back, dead):
//   * once a group of rows has been read back, its scratch lines are DISCARDED (discard.global.L2 = drop the
//     dirty lines from L2 without a write-back): the intermediate never reaches HBM and the live footprint
//     is half a slot per CTA;
//   * input loads and output stores carry streaming hints.  (Pulling the next transform's input into L2 by a
//     bulk prefetch during step B was measured twice and is a loss: 0.54 -> 0.45, and 0.565 -> 0.495 with the scratch
//     stored evict_last, profiles/r02ag_ab_long_l2prefetch_rejected.txt.)
#pragma once

#include "abi.h"
#include "fft_engine.cuh"
#include "rows_kernel.cuh"
#include "tma_util.cuh"

namespace mufft {

#ifndef MUFFT_LONG_TR
#define MUFFT_LONG_TR 8 // rows per group of step B (8: 64-byte output runs; 16: 128-byte runs, twice the CTA)
#endif
#ifndef MUFFT_LONG_PACKED_A
#define MUFFT_LONG_PACKED_A 1 // packed complex adds (FADD2) in the register transforms of step A (measured: 65536 points 0.450 -> 0.460)
#endif
#ifndef MUFFT_LONG_MIN_BLOCKS
#define MUFFT_LONG_MIN_BLOCKS 2 // CTAs of 256 threads per SM
#endif
// Store of an intermediate that is read back from L2 a few microseconds later: the line is marked evict_last, so that the
// streaming input / output lines passing through L2 push it out less easily (measured, profiles/r02o_ab.txt: 32768 points
// 0.544 -> 0.564 of the roofline, 65536 points 0.437 -> 0.450, long columns 0.442 -> 0.446; -DMUFFT_SCRATCH_EVICT_LAST=0
// gives plain stores).
#ifndef MUFFT_SCRATCH_EVICT_LAST
#define MUFFT_SCRATCH_EVICT_LAST 1
#endif
__device__ __forceinline__ unsigned long long scratch_policy()
{
#if MUFFT_SCRATCH_EVICT_LAST && !defined(MUFFT_EMU)
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, " MUFFT_STR(MUFFT_EVICT_LAST_FRACTION) ";" : "=l"(p));
    return p;
#else
    return 0;
#endif
}
__device__ __forceinline__ void st_scratch(cpx *p, cpx v, unsigned long long policy)
{
#if MUFFT_SCRATCH_EVICT_LAST && !defined(MUFFT_EMU)
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(policy) : "memory");
#else
    (void)policy;
    *p = v;
#endif
}
// Drop 128 bytes (aligned) from L2 without writing them back: the data is dead.
__device__ __forceinline__ void discard_l2_128(const void *p)
{
#ifdef MUFFT_EMU
    (void)p;
#else
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
#endif
}
// Rows per group of step B.  Local output: 8 (64-byte runs; 16 measured the same).  Scatter to peer slabs over NVLink:
// 16, so that every store run is a whole 128-byte line.  16 rows need a 1024-thread CTA and 270 KB of exchange area at
// n2 = 2048: 8 there either way.
#ifndef MUFFT_LONG_TR2
#define MUFFT_LONG_TR2 8 // rows per group at n2 = 2048 (4: two CTAs of 256 threads per SM instead of one of 512, 32-byte runs)
#endif
__host__ __device__ constexpr int long_tile_rows(int n2, bool peers = false) { return n2 <= 1024 ? (peers ? 16 : MUFFT_LONG_TR) : MUFFT_LONG_TR2; }
__host__ __device__ constexpr int long_cta_threads(int n2, bool peers = false) { return long_tile_rows(n2, peers) * engine_threads(n2); }
// Row pitch of the exchange area, in cpx: after the mapping switch a half-warp reads TR rows x 16/TR
// consecutive elements; pitch = 16/TR (mod 16) for TR <= 8, odd for TR = 16 puts them in 16 different bank pairs.
__host__ __device__ constexpr int long_row_pitch(int n2, bool peers = false)
{
    int se = engine_smem_elems(n2);
    const int tr = long_tile_rows(n2, peers);
    if (tr >= 16) return se | 1;
    while (se % 16 != 16 / tr) se++;
    return se;
}
__host__ __device__ constexpr size_t long_smem_bytes(int n2, bool peers = false) { return (size_t)long_tile_rows(n2, peers) * long_row_pitch(n2, peers) * sizeof(float2); }
__host__ __device__ constexpr bool long_supported(int n1, int n2, bool peers = false)
{
    return (n1 == 16 || n1 == 32) && engine_regs(n2) > 0 && engine_num_passes(n2) >= 2 && long_cta_threads(n2, peers) <= 1024 &&
           n2 % long_cta_threads(n2, peers) == 0 && n1 % long_tile_rows(n2, peers) == 0 && long_smem_bytes(n2, peers) <= 227 * 1024 &&
           long_tile_rows(n2, peers) <= 15 + (engine_threads(n2) <= 32 ? 64 : 0); // named barriers 1..15 when a row spans several warps
}
// ---- step A: N1-point transforms down the columns of the [N1][N2] view, `THREADS` threads -------------------
// One column: x = its N1 inputs -> transform, inter-stage twiddles w_N^(k1 c), into the scratch slot.
template <int N1, int N2, int DIR>
__device__ __forceinline__ void long_step_a_column(cpx (&x)[N1], int c, cpx *__restrict__ scratch, const cpx *__restrict__ roots, unsigned long long policy)
{
    constexpr int LG = ilog2(N1);
    dft_regs<N1, DIR, MUFFT_LONG_PACKED_A != 0>(x);
    cpx base[LG];
#pragma unroll
    for (int j = 0; j < LG; j++) base[j] = __ldg(roots + ((unsigned)c << j));
    TwTree<N1, LG, LG - 1, 0, true>::run(x, base, make_float2(1.0f, 0.0f));
#pragma unroll
    for (int k1 = 0; k1 < N1; k1++) st_scratch(scratch + k1 * N2 + c, x[k1], policy);
}
template <int N1, int N2, int DIR, int IN_MODE, int THREADS>
__device__ __forceinline__ void long_step_a(const cpx *__restrict__ in, cpx *__restrict__ scratch, const cpx *__restrict__ roots, int tid)
{
    const unsigned long long policy = scratch_policy();
#pragma unroll 1
    for (int c = tid; c < N2; c += THREADS) {
        cpx x[N1];
#pragma unroll
        for (int m = 0; m < N1; m++)
            x[m] = (IN_MODE == MCU_IN_HALF && m >= N1 / 2) ? make_float2(0.0f, 0.0f) : __ldcs(in + c + m * N2); // rows >= N1/2 are the zero half: never read
        long_step_a_column<N1, N2, DIR>(x, c, scratch, roots, policy);
    }
}

// ---- step B: N2-point transforms along the rows of Y, TR rows at a time; `sync` = barrier of the THREADS threads
// `last_gather` runs (every thread) once the last group's rows have left the exchange area for good.
template <int N1, int N2, int DIR, bool TO_PEERS, typename Sync, typename LastGather>
__device__ __forceinline__ void long_step_b(const mcu_long_args &a, size_t b, cpx *__restrict__ scratch, cpx *smem, const cpx *__restrict__ tw, int tid, Sync sync,
                                            LastGather last_gather)
{
    constexpr int THREADS = long_cta_threads(N2, TO_PEERS);
    constexpr int R = engine_regs(N2);
    constexpr int T = N2 / R;
    constexpr int TR = long_tile_rows(N2, TO_PEERS);
    constexpr int NP = engine_num_passes(N2);
    constexpr int SH = engine_pad_shift(N2);
    constexpr int SE = long_row_pitch(N2, TO_PEERS);
    const int slot = tid / T, t = tid - slot * T;   // passes 0 .. NP-2: lanes walk along a row
    const int slot2 = tid % TR, t2 = tid / TR;      // last pass: lanes walk across the TR rows
#pragma unroll 1
    for (int g = 0; g < N1 / TR; g++) {
        cpx v[R];
        {
            const cpx *__restrict__ y = scratch + (size_t)(g * TR + slot) * N2;
#pragma unroll
            for (int e = 0; e < R; e++) v[e] = __ldcg(y + t + e * T);
        }
        cpx *sm = smem + slot * SE;
        static_for<NP - 1>([&](auto pc) {
            constexpr int PASS = decltype(pc)::value;
            engine_butterflies<N2, DIR, PASS>(v, tw, t);
            if constexpr (PASS > 0) engine_sync<T, TR>(slot);
            engine_scatter<N2, PASS, 1>(v, sm, t);
            if constexpr (PASS + 1 < NP - 1) {
                engine_sync<T, TR>(slot);
                engine_gather<N2, 1>(v, sm, t);
            }
        });
        sync();
        {
            // every thread has its part of this group's rows in registers / shared memory: the scratch lines are dead
            // (measured at 32768 points: 0.375 -> 0.54 of the HBM roofline with the discard)
            const char *dead = reinterpret_cast<const char *>(scratch + (size_t)g * TR * N2);
            for (int l = tid; l < TR * N2 * (int)sizeof(cpx) / 128; l += THREADS) discard_l2_128(dead + (size_t)l * 128);
        }
        const cpx *sm2 = smem + slot2 * SE;
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = sm2[pad_pos<SH>(t2 + e * T)];
        if (g == N1 / TR - 1) last_gather();
        engine_butterflies<N2, DIR, NP - 1>(v, tw, t2);
        if constexpr (TO_PEERS) {
            // exchange step of a slab 2D transform: X[x] goes to the column slab of device x >> shift
            const size_t row = (a.peers.row0 + b) * a.peers.pitch;
            const unsigned cmask = (1u << a.peers.shift) - 1u;
#pragma unroll
            for (int e = 0; e < R; e++) {
                const unsigned x = (unsigned)(t2 + e * T) * N1 + g * TR + slot2;
                static_cast<cpx *>(a.peers.base[x >> a.peers.shift])[row + (x & cmask)] = v[e];
            }
        } else {
            cpx *__restrict__ out = static_cast<cpx *>(a.out) + b * a.out_batch + g * TR + slot2;
#pragma unroll
            for (int e = 0; e < R; e++) __stcs(out + (size_t)(t2 + e * T) * N1, v[e]);
        }
        sync(); // the exchange area is reused by the next group, the scratch slot by the next transform
    }
}

// One CTA = one transform at a time: step A, barrier, step B; two CTAs of 256 threads per SM drift against each other.
// (A variant with step A of transform j + 1 and step B of transform j on two halves of a 512-thread CTA - producer /
// consumer hand-over by named barriers, two scratch slots per CTA - measured the same 0.54 of the roofline at 32768
// points as this kernel with 8 or 16 rows per group: with ~110 registers per thread an SM holds 16 warps whichever way
// they are arranged, and ncu shows them waiting for loads a third of the time.  profiles/README.md keeps the numbers.)
// TMA_A: the inputs of step A arrive through the exchange area, which step A itself does not use.  Thread 0 asks the TMA
// unit for the N1 segments of the NEXT sweep of columns (THREADS columns, 64 / 128 KB) as soon as every thread has pulled
// its column of this sweep into registers; the copy lands while the butterflies run and the scratch is written.  The first
// sweep of the next transform is asked for when the last group of step B has left the exchange area.  HBM loads no longer
// stall the warps; needs 16-byte aligned transforms.
template <int N1, int N2, int DIR, int IN_MODE, bool TO_PEERS, bool TMA_A = false>
__global__ void __launch_bounds__(long_cta_threads(N2, TO_PEERS), long_cta_threads(N2, TO_PEERS) <= 256 ? MUFFT_LONG_MIN_BLOCKS : 1)
fft_long_kernel(const mcu_long_args a)
{
    constexpr int N = N1 * N2;
    constexpr int THREADS = long_cta_threads(N2, TO_PEERS);
    static_assert(long_supported(N1, N2, TO_PEERS), "shape");
    static_assert(IN_MODE == MCU_IN_COMPLEX || IN_MODE == MCU_IN_HALF, "input modes: complex, zero-padded upper half");
    MUFFT_DYNAMIC_SMEM(smem_raw);
    cpx *smem = reinterpret_cast<cpx *>(smem_raw);
    const int tid = threadIdx.x;
    const cpx *__restrict__ tw = static_cast<const cpx *>(a.tw);
    const cpx *__restrict__ roots = static_cast<const cpx *>(a.roots);
    cpx *__restrict__ scratch = static_cast<cpx *>(a.scratch) + (size_t)blockIdx.x * N;

    if constexpr (TMA_A) {
        constexpr int SWEEPS = N2 / THREADS;
        constexpr int ROWS_IN = IN_MODE == MCU_IN_HALF ? N1 / 2 : N1; // rows >= N1/2 are the zero half: never read
        constexpr unsigned SWEEP_BYTES = ROWS_IN * THREADS * sizeof(cpx);
        static_assert((size_t)N1 * THREADS * sizeof(cpx) <= long_smem_bytes(N2, TO_PEERS), "a sweep of columns fits the exchange area");
        unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + (long_smem_bytes(N2, TO_PEERS) + 15) / 16 * 16);
        const unsigned long long pol_in = l2_policy<1>(), policy = scratch_policy();
        auto ask = [&](size_t b, int sweep) { // thread 0, after a CTA barrier behind the last reads of the exchange area
            const cpx *src = static_cast<const cpx *>(a.in) + b * a.in_batch + sweep * THREADS;
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(full, SWEEP_BYTES);
#pragma unroll 1
            for (int m = 0; m < ROWS_IN; m++) bulk_load_1d<1>(smem + m * THREADS, src + (size_t)m * N2, THREADS * sizeof(cpx), full, pol_in);
        };
        if (tid == 0) {
            mbar_init(full, 1);
            mbar_init_fence();
            if (blockIdx.x < a.batch) ask(blockIdx.x, 0);
        }
        __syncthreads();
        unsigned phase = 0;
        for (size_t b = blockIdx.x; b < a.batch; b += gridDim.x) {
#pragma unroll 1
            for (int sweep = 0; sweep < SWEEPS; sweep++) {
                MUFFT_TMA_WAIT(0, full, phase, "sweep", sweep, phase);
                phase ^= 1;
                cpx x[N1];
#pragma unroll
                for (int m = 0; m < N1; m++) x[m] = m < ROWS_IN ? smem[m * THREADS + tid] : make_float2(0.0f, 0.0f);
                __syncthreads(); // every column of the sweep is in registers
                if (tid == 0 && sweep + 1 < SWEEPS) ask(b, sweep + 1);
                long_step_a_column<N1, N2, DIR>(x, sweep * THREADS + tid, scratch, roots, policy);
            }
            __syncthreads(); // Y is complete and visible to the whole CTA (the reads of step B bypass L1)
            long_step_b<N1, N2, DIR, TO_PEERS>(a, b, scratch, smem, tw, tid, [] { __syncthreads(); }, [&] {
                __syncthreads();
                if (tid == 0 && b + gridDim.x < a.batch) ask(b + gridDim.x, 0);
            });
        }
    } else {
        for (size_t b = blockIdx.x; b < a.batch; b += gridDim.x) {
            const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + b * a.in_batch;
            long_step_a<N1, N2, DIR, IN_MODE, THREADS>(in, scratch, roots, tid);
            __syncthreads(); // Y is complete and visible to the whole CTA (the reads of step B bypass L1)
            long_step_b<N1, N2, DIR, TO_PEERS>(a, b, scratch, smem, tw, tid, [] { __syncthreads(); }, [] {});
        }
    }
}

} // namespace mufft
