// Column transforms of 16384 and 32768 points (the column pass of a slab-decomposed 2D transform, BASELINE configs[4])
// in ONE persistent launch and ONE pass over HBM: the four-step split n = N1 x 1024 down the columns of a row-major
// [n][width] matrix, with the intermediate of a block of 32 columns kept in L2.
//
//   in[(y1*N2 + y2)][c]  --A-->  S[k1][y2][c] = w_n^(k1 y2) * sum_y1 in[y1*N2 + y2][c] w_N1^(y1 k1)
//                        --B-->  out[k1 + N1*k2][c] = sum_y2 S[k1][y2][c] w_N2^(y2 k2)
//
// The matrix is cut into blocks of CB = 32 columns (8 MiB of intermediate at 32768 rows).  Tickets, as in the fused 2D
// kernel (fft2d_kernel.cuh): the A tiles of block i are producer tickets, the B tiles of block i - lag consumer
// tickets; a B ticket waits until every A tile of its block has announced itself, an A ticket until the B tiles of
// the block that used its ring slot before are done.  Tickets are handed out in order and every wait is on earlier
// tickets: no deadlock whatever the number of resident CTAs.
//   A tile  16 rows y2 x 16 columns, one thread each: N1 loads N2 rows apart (64-byte row segments... 128 with 16 columns),
//           the whole N1-point transform as one register butterfly, inter-stage twiddles derived from log2 N1 roots
//           (as long_kernel.cuh), into the ring slot - no shared memory, no barrier;
//   B tile  one k1 x 8 columns: 1024-point column transforms on the engine (the fused 2D kernel's column tile),
//           stores to rows k1 + N1*k2; the tile's part of the slot is then DISCARDED from L2 (dead, never written back).
// In place (out == in) is fine: a block's B tiles start only when all its A tiles have read the block.
//
// Replaces, for these lengths, the two sweeps of the column kernel through scratch in HBM (32 n bytes of traffic per
// column instead of 16 n; measured 0.31 of the roofline, profiles/r02g_ab_cols_blocks.txt - L2-sized blocks of separate
// launches were measured too and lose to launch gaps).  The reference makes log8(n) = 5 sweeps (kernel.c:604-738).
#pragma once

#include "abi.h"
#include "fft_engine.cuh"
#include "fft2d_kernel.cuh" // ticket protocol helpers: ld_relaxed, red_relaxed_add, one-sided fences
#include "long_kernel.cuh"  // st_scratch

namespace mufft {

#define LONGCOLS_N2 1024
#define LONGCOLS_CB MUFFT_LONGCOLS_CB // columns per block (abi.h)
#define LONGCOLS_THREADS 256
__host__ __device__ constexpr bool longcols_supported(int n1) { return n1 == 16 || n1 == 32; } // 4 / 8 work too but lose to two sweeps
__host__ __device__ constexpr unsigned longcols_a_tiles() { return (LONGCOLS_N2 / 16) * (LONGCOLS_CB / 16); }
__host__ __device__ constexpr unsigned longcols_b_tiles(int n1) { return (unsigned)n1 * (LONGCOLS_CB / 8); }
__host__ __device__ constexpr size_t longcols_slot_elems(int n1) { return (size_t)n1 * LONGCOLS_N2 * LONGCOLS_CB; }
__host__ __device__ constexpr size_t longcols_smem_bytes() { return (size_t)8 * engine_smem_elems(LONGCOLS_N2) * sizeof(float2); }

template <int N1, int DIR>
__global__ void __launch_bounds__(LONGCOLS_THREADS, 2)
fft_longcols_kernel(const mcu_longcols_args a)
{
    constexpr int N2 = LONGCOLS_N2, CB = LONGCOLS_CB, THREADS = LONGCOLS_THREADS;
    constexpr int LG = ilog2(N1);
    constexpr int RY = engine_regs(N2), TY = N2 / RY; // 32 values on 32 threads per column
    constexpr unsigned NP = longcols_a_tiles(), NQ = longcols_b_tiles(N1);
    static_assert(TY * 8 == THREADS, "a B tile is 8 columns");
    MUFFT_DYNAMIC_SMEM(smem_raw);
    MUFFT_STATIC_SMEM_U32(s_ticket, 2);
    MUFFT_STATIC_SMEM_U32(s_wait, 2);
    cpx *smem = reinterpret_cast<cpx *>(smem_raw);

    const int tid = threadIdx.x;
    const cpx *__restrict__ tw = static_cast<const cpx *>(a.tw);
    const cpx *__restrict__ roots = static_cast<const cpx *>(a.roots);
    const unsigned blocks = (a.width + CB - 1) / CB;
    unsigned *ticket_counter = a.counters;
    unsigned *produced = a.counters + 1;          // [blocks] A tiles finished per block
    unsigned *consumed = a.counters + 1 + blocks; // [blocks] B tiles finished per block (frees the ring slot)
    const unsigned long long total = (unsigned long long)blocks * (NP + NQ);
    const unsigned lag = a.lag < blocks ? a.lag : blocks;
    const unsigned long long head = (unsigned long long)lag * NP;
    const unsigned long long body = (unsigned long long)(blocks - lag) * (NP + NQ);

    auto decode = [&](unsigned k, unsigned &block, unsigned &tile) -> bool {
        if (k < head) {
            block = k / NP; tile = k % NP;
            return true;
        } else if (k - head < body) {
            const unsigned long long kk = k - head;
            const unsigned s = lag + (unsigned)(kk / (NP + NQ)), j = (unsigned)(kk % (NP + NQ));
            if (j < NP) { block = s; tile = j; return true; }
            block = s - lag; tile = j - NP;
            return false;
        }
        const unsigned long long kk = k - head - body;
        block = blocks - lag + (unsigned)(kk / NQ); tile = (unsigned)(kk % NQ);
        return false;
    };
    unsigned *pending = nullptr; // thread 0: a finished tile whose announcement is held back until the next tile's loads fly
    auto release_pending = [&]() {
        if (pending != nullptr) {
            fence_release_gpu();
            red_relaxed_add(pending, 1u);
            pending = nullptr;
        }
    };
    auto dependency = [&](unsigned k, unsigned &need, bool &reads) -> const unsigned * {
        unsigned block, tile;
        if (!decode(k, block, tile)) { need = NP; reads = true; return produced + block; }
        reads = false; // a producer only overwrites its ring slot
        if (block >= a.ring) { need = NQ; return consumed + (block - a.ring); }
        return nullptr;
    };
    auto wait_for = [&](const unsigned *counter, unsigned need) {
        unsigned ns = 32;
        while (ld_relaxed(counter) < need) {
            __nanosleep(ns);
            if (ns < 1024) ns *= 2;
        }
    };
    auto publish = [&](unsigned k, int slot) {
        unsigned wait = 0;
        if (k < total) {
            unsigned need = 0;
            bool reads = false;
            const unsigned *dep = dependency(k, need, reads);
            if (dep != nullptr) {
                if (ld_relaxed(dep) < need) {
                    if (pending != nullptr) wait = 1;
                    else wait_for(dep, need);
                }
                if (reads) fence_acquire_gpu();
            }
        }
        s_ticket[slot] = k;
        s_wait[slot] = wait;
    };

    const unsigned long long policy = scratch_policy();
    unsigned next_ticket = 0;
    if (tid == 0) publish(atomicAdd(ticket_counter, 1u), 0);
    for (int it = 0;; it++) {
        __syncthreads();
        const unsigned k = s_ticket[it & 1];
        if (k >= total) {
            if (tid == 0) release_pending();
            break;
        }
        if (s_wait[it & 1]) {
            if (tid == 0) {
                unsigned need = 0;
                bool reads = false;
                const unsigned *dep = dependency(k, need, reads);
                release_pending();
                wait_for(dep, need);
                if (reads) fence_acquire_gpu();
            }
            __syncthreads();
        }
        if (tid == 0) next_ticket = atomicAdd(ticket_counter, 1u);
        unsigned block, tile;
        const bool producer = decode(k, block, tile);
        cpx *__restrict__ slot_base = static_cast<cpx *>(a.scratch) + (size_t)(block % a.ring) * longcols_slot_elems(N1);
        // slot layout [k1][column octet (CB/8)][y2][8 columns]: the 64 KiB a B tile reads are one contiguous run of whole lines

        if (producer) {
            // ---- A tile: 16 rows y2 x 16 columns, one N1-point register transform per thread -------------------
            const unsigned ct = tile % (CB / 16), yt = tile / (CB / 16);
            const int c = tid % 16;
            const unsigned y2 = yt * 16 + tid / 16;
            const unsigned cb = ct * 16 + c;                 // column inside the block
            const size_t col = (size_t)block * CB + cb;
            const bool active = col < a.width;
            const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + (size_t)y2 * a.in_pitch + col;
            cpx x[N1];
#pragma unroll
            for (int m = 0; m < N1; m++) x[m] = active ? __ldcs(in + (size_t)m * N2 * a.in_pitch) : make_float2(0.0f, 0.0f);
            if (tid == 0) release_pending();
            dft_regs<N1, DIR, MUFFT_LONG_PACKED_A != 0>(x);
            cpx base[LG];
#pragma unroll
            for (int j = 0; j < LG; j++) base[j] = __ldg(roots + (y2 << j));
            TwTree<N1, LG, LG - 1, 0, true>::run(x, base, make_float2(1.0f, 0.0f));
            cpx *__restrict__ s = slot_base + ((size_t)(cb / 8) * N2 + y2) * 8 + (cb % 8);
#pragma unroll
            for (int k1 = 0; k1 < N1; k1++) st_scratch(s + (size_t)k1 * (CB / 8) * N2 * 8, x[k1], policy);
            if (tid == 0) pending = produced + block;
        } else {
            // ---- B tile: row group k1, 8 columns: 1024-point column transforms on the engine ---------------------
            const unsigned k1 = tile / (CB / 8), octet = tile % (CB / 8);
            const int c = tid % 8, t = tid / 8;
            const size_t col = (size_t)block * CB + octet * 8 + c;
            const bool active = col < a.width;
            const cpx *__restrict__ s = slot_base + ((size_t)k1 * (CB / 8) + octet) * N2 * 8; // [y2][8]
            cpx v[RY];
#pragma unroll
            for (int e = 0; e < RY; e++) v[e] = __ldcg(s + (size_t)(t + e * TY) * 8 + c);
            if (tid == 0) release_pending();
            engine_run<N2, DIR, 8, THREADS, 1>(v, smem + c, tw, t, 0);
            if (active) {
                cpx *__restrict__ out = static_cast<cpx *>(a.out) + (size_t)k1 * a.out_pitch + col;
#pragma unroll
                for (int e = 0; e < RY; e++) __stcs(out + (size_t)(t + e * TY) * N1 * a.out_pitch, v[e]);
            }
            // every thread's loads completed before the engine's first exchange barrier: the tile's 64 KiB are dead
#ifndef MUFFT_EMU
            for (int l = tid; l < N2 * 8 * (int)sizeof(cpx) / 128; l += THREADS)
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(reinterpret_cast<const char *>(s) + (size_t)l * 128) : "memory");
#endif
            if (tid == 0) pending = consumed + block;
        }
        if (tid == 0) publish(next_ticket, (it + 1) & 1);
    }
}

} // namespace mufft
