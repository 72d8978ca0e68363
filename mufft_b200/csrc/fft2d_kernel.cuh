// Whole batched 2D C2C transforms in ONE persistent kernel: CTAs draw tickets from a global
// counter; a ticket is either a tile of rows or a tile of columns of one image.  Row tiles read the
// input and write the half-transformed image into the OUTPUT buffer; column tiles transform
// their columns of the output buffer in place.  Tickets are ordered
//     rows(0), rows(1), cols(0), rows(2), cols(1), ... rows(B-1), cols(B-2), cols(B-1)
// so the column pass of image i runs one image behind its row pass, while the 8 MiB (1024^2)
// intermediate is still resident in the 126 MB L2: HBM sees the input once and the output once.
//
// A column ticket of image i waits (thread 0 spins on an acquire load, with back-off) until all
// row tickets of image i have signalled completion.  Tickets are handed out in increasing order
// and row tickets never wait, so every wait is on tickets that are already running: no deadlock
// whatever the number of resident CTAs.
//
// Replaces the reference's 2D driver (fft.c:852-989): Ny*steps_x horizontal kernel calls plus
// steps_y vertical sweeps over the whole image (kernel.c:455-738), every one through memory.
#pragma once

#include "abi.h"
#include "fft_engine.cuh"
#include "rows_kernel.cuh"

namespace mufft {

#ifndef MUFFT_F2D_THREADS
#define MUFFT_F2D_THREADS 256
#endif
#ifndef MUFFT_F2D_MIN_BLOCKS
#define MUFFT_F2D_MIN_BLOCKS 2
#endif

// Shapes whose sides both run with 16 values per thread (256 and 512 points) need far fewer registers:
// at 3 CTAs per SM (80 registers) they compile without spills, at 4 (64 registers) with 36 bytes.  The
// default keeps them at MUFFT_F2D_MIN_BLOCKS until the extra occupancy has been measured
// (tools/ab_r2.txt: f2d_small3 / f2d_small4).
#ifndef MUFFT_F2D_MIN_BLOCKS_R16
#define MUFFT_F2D_MIN_BLOCKS_R16 MUFFT_F2D_MIN_BLOCKS
#endif
__host__ __device__ constexpr int f2d_min_blocks(int nx, int ny)
{
    return engine_regs(nx) <= 16 && engine_regs(ny) <= 16 ? MUFFT_F2D_MIN_BLOCKS_R16 : MUFFT_F2D_MIN_BLOCKS;
}
__host__ __device__ constexpr int f2d_rows_per_tile(int nx) { return MUFFT_F2D_THREADS / engine_threads(nx); }
__host__ __device__ constexpr int f2d_cols_per_tile(int ny) { return MUFFT_F2D_THREADS / engine_threads(ny); }
// A row ticket covers f2d_row_repeat() consecutive groups of f2d_rows_per_tile() rows, each warp
// working through its rows back to back without a CTA-wide barrier in between.  Measured on B200
// (profiles/README.md): 4 groups per ticket is SLOWER than 1 for 1024^2 x 256 (1.65 ms vs 1.27 ms;
// coarser tickets balance worse against the column tickets), so the default is 1; the knob stays
// for other shapes.  At least 16 row tickets per image are kept for load balance.
#ifndef MUFFT_F2D_ROW_REPEAT
#define MUFFT_F2D_ROW_REPEAT 1
#endif
__host__ __device__ constexpr int f2d_row_repeat(int nx, int ny)
{
    int k = ny / (16 * f2d_rows_per_tile(nx));
    return k < 1 ? 1 : (k > MUFFT_F2D_ROW_REPEAT ? MUFFT_F2D_ROW_REPEAT : k);
}
__host__ __device__ constexpr bool f2d_supported(int nx, int ny)
{
    return engine_regs(nx) > 0 && engine_regs(ny) > 0 && engine_threads(nx) <= MUFFT_F2D_THREADS &&
           engine_threads(ny) * 4 <= MUFFT_F2D_THREADS && (engine_threads(nx) <= 32 || f2d_rows_per_tile(nx) <= 15) &&
           ny % (f2d_rows_per_tile(nx) * f2d_row_repeat(nx, ny)) == 0 && nx % f2d_cols_per_tile(ny) == 0;
}
__host__ __device__ constexpr size_t f2d_xch_elems(int nx, int ny, int vx = 0)
{
    size_t r = (size_t)f2d_rows_per_tile(nx) * engine_smem_elems(nx, vx);
    size_t c = (size_t)f2d_cols_per_tile(ny) * engine_smem_elems(ny);
    return r > c ? r : c;
}
__host__ __device__ constexpr size_t f2d_smem_bytes(int nx, int ny, bool lite = false, int vx = 0)
{
    return f2d_xch_elems(nx, ny, vx) * sizeof(float2) + 16 +
           (lite ? (size_t)(engine_tw_total(nx, vx) + ((nx != ny || vx != 0) ? engine_tw_total(ny) : 0)) * sizeof(float2) : 0);
}
// Real-to-complex images (OM = MCU_OUT_R2C / MCU_OUT_R2C_FULL): NX is the complex length of a row
// transform (half the real row), the output rows have a pitch of 2 NX bins (ref fft.h:238-257) of which
// NX + 1 (all 2 NX with FULL) carry data and get column transforms.
__host__ __device__ constexpr int f2d_rows_variant(int nx, int om) { return rows_variant(nx, MCU_IN_COMPLEX, om); }
__host__ __device__ constexpr bool f2d_real_supported(int nx, int ny, int om)
{
    return f2d_supported(nx, ny) && f2d_row_repeat(nx, ny) == 1 &&
           engine_threads(nx, f2d_rows_variant(nx, om)) == engine_threads(nx); // same tile shape as the complex kernel
}

// Polling load (no L1 invalidation per iteration; the acquire fence follows once, after success).
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p)
{
#ifdef MUFFT_EMU
    return __atomic_load_n(p, __ATOMIC_RELAXED);
#else
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
// Release increment: the CTA's row stores (ordered before it by the preceding bar.sync) become
// visible to whoever acquires the counter.
__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v)
{
#ifdef MUFFT_EMU
    __atomic_fetch_add(p, v, __ATOMIC_RELEASE);
#else
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ void red_relaxed_add(unsigned *p, unsigned v)
{
#ifdef MUFFT_EMU
    __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
#else
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ void fence_acq_rel_gpu()
{
#ifdef MUFFT_EMU
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
#else
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}
// One-sided fences.  fence.acq_rel.gpu is MEMBAR.ALL.GPU + CCTL.IVALL in SASS: the second half throws
// away the SM's whole L1 - twiddle tables included - which a RELEASE does not need, and the first
// half waits for this thread's outstanding writes, which an ACQUIRE does not need.
__device__ __forceinline__ void fence_release_gpu()
{
#ifdef MUFFT_EMU
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
#else
    asm volatile("fence.release.gpu;" ::: "memory");
#endif
}
__device__ __forceinline__ void fence_acquire_gpu()
{
#ifdef MUFFT_EMU
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
#else
    asm volatile("fence.acquire.gpu;" ::: "memory");
#endif
}
// LITE kernels: one-sided fences and twiddle tables in shared memory (see fft2d_fused_kernel).
template <bool LITE>
__device__ __forceinline__ void f2d_fence_release()
{
    if constexpr (LITE) fence_release_gpu();
    else fence_acq_rel_gpu();
}
template <bool LITE>
__device__ __forceinline__ void f2d_fence_acquire()
{
    if constexpr (LITE) fence_acquire_gpu();
    else fence_acq_rel_gpu();
}
// Twiddle tables of both axes in shared memory (one table when the sides are equal and the rows use
// the general engine shape, vx == 0), in cpx.
__host__ __device__ constexpr int f2d_tw_elems(int nx, int ny, int vx = 0) { return engine_tw_total(nx, vx) + ((nx != ny || vx != 0) ? engine_tw_total(ny) : 0); }
// Copies them behind `dst`; the caller's next CTA barrier publishes them.
template <int NX, int NY, int VX = 0>
__device__ __forceinline__ void f2d_stage_twiddles(cpx *dst, const cpx *__restrict__ twx, const cpx *__restrict__ twy, int tid, int threads)
{
    for (int i = tid; i < engine_tw_total(NX, VX); i += threads) dst[i] = __ldg(twx + i);
    if constexpr (NX != NY || VX != 0)
        for (int i = tid; i < engine_tw_total(NY); i += threads) dst[engine_tw_total(NX, VX) + i] = __ldg(twy + i);
}
// Pull `bytes` (a multiple of 16, 16-byte aligned) into L2 ahead of use.
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned bytes)
{
#ifdef MUFFT_EMU
    (void)p; (void)bytes;
#else
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#endif
}

// FLAGS (kernel flavours, chosen per shape by the launcher from measurements):
//   F2D_SMEM_TW    twiddle tables staged in shared memory once per CTA.  ncu on the plain kernel
//                  (profiles/r01c_c2c2d1024_full.txt) shows the twiddle loads missing L1 more often
//                  than not - every fence.acq_rel.gpu of the ticket protocol invalidates it - so that
//                  they pull as many sectors from L2 as a whole pass over the data;
//   F2D_ONE_SIDED  release / acquire fences instead of fence.acq_rel.gpu in the ticket protocol;
//   F2D_DERIVE     only the power-of-two twiddles are loaded, the rest derived (MUFFT_TW_DERIVE);
//   F2D_EARLY      tickets are drawn TWO tiles ahead, so that the rows-done counter a column ticket
//                  depends on is read at the START of the tile before it and the answer is in a
//                  register when the ticket is handed to the CTA: the L2 round trip of that read leaves
//                  the critical path between two tiles (not measured yet).
//   F2D_OCC3       compiled for THREE CTAs per SM (80 registers per thread, the excess spilled; 24 warps per
//                  SM instead of 16).  ncu on the plain kernel: 25 % warp occupancy, issue slots 43 % busy,
//                  barrier + scoreboard stalls on top - more warps to hide them against the cost of the
//                  spills (not measured yet).  With F2D_SMEM_TW three exchange areas plus three table copies fit
//                  the shared memory of an SM only for 1024 x 1024 (3 x 75 600 bytes) and the 256 / 512-point shapes.
//   F2D_INTERLEAVE inside a round of tickets (the NR row tickets of image s and the NC column tickets of image
//                  s - lag) row and column tickets alternate (evenly spread when NR != NC) instead of all rows
//                  first: the two CTAs of an SM then tend to hold one tile of each kind - row tiles read HBM
//                  and write L2, column tiles read L2 and write HBM - instead of two of the same kind, and
//                  the HBM reads and writes are spread evenly over time (not measured yet).
#define F2D_SMEM_TW 1
#define F2D_ONE_SIDED 2
#define F2D_DERIVE 4
#define F2D_EARLY 8
#define F2D_OCC3 16
#define F2D_INTERLEAVE 32
#define F2D_LITE (F2D_SMEM_TW | F2D_ONE_SIDED)
__host__ __device__ constexpr int f2d_twm(int flags) { return ((flags & F2D_SMEM_TW) ? MUFFT_TW_SMEM : 0) | ((flags & F2D_DERIVE) ? MUFFT_TW_DERIVE : 0); }
// OM: MCU_OUT_COMPLEX = complex images; MCU_OUT_R2C / MCU_OUT_R2C_FULL = real images (forward only): the row
// tiles run the real-to-complex row transform (engine of NX points + the resolve epilogue of
// rows_kernel.cuh, ref kernel.c:47-82) into rows of pitch 2 NX, the column tiles transform the NX + 1
// (2 NX) columns that carry data; the last column tile is partly masked.
template <int NX, int NY, int DIR, int FLAGS = 0, int OM = MCU_OUT_COMPLEX>
__global__ void __launch_bounds__(MUFFT_F2D_THREADS, (FLAGS & F2D_OCC3) != 0 && f2d_min_blocks(NX, NY) < 3 ? 3 : f2d_min_blocks(NX, NY))
fft2d_fused_kernel(const mcu_2d_args a)
{
    constexpr int VX = f2d_rows_variant(NX, OM);
    constexpr int RX = engine_regs(NX, VX), TX = NX / RX, ROWS = f2d_rows_per_tile(NX);
    constexpr int RY = engine_regs(NY), TY = NY / RY, COLS = f2d_cols_per_tile(NY);
    constexpr int REP = f2d_row_repeat(NX, NY);
    constexpr int PITCH = OM == MCU_OUT_COMPLEX ? NX : 2 * NX;                               // output row pitch
    constexpr int WIDTH = OM == MCU_OUT_COMPLEX ? NX : (OM == MCU_OUT_R2C ? NX + 1 : 2 * NX); // columns to transform
    static_assert(OM == MCU_OUT_COMPLEX || (REP == 1 && DIR < 0 && MUFFT_F2D_THREADS / TX == ROWS), "real images: forward, one row group per ticket");
    constexpr unsigned NR = NY / (ROWS * REP); // row tickets per image
    constexpr unsigned NC = (WIDTH + COLS - 1) / COLS; // column tickets per image
    MUFFT_DYNAMIC_SMEM(smem_raw);
    MUFFT_STATIC_SMEM_U32(s_ticket, 2);
    MUFFT_STATIC_SMEM_U32(s_wait, 2);
    cpx *smem = reinterpret_cast<cpx *>(smem_raw);

    const int tid = threadIdx.x;
    const cpx *__restrict__ twx = static_cast<const cpx *>(a.tw_x);
    const cpx *__restrict__ twy = static_cast<const cpx *>(a.tw_y);
    constexpr bool LITE = (FLAGS & F2D_ONE_SIDED) != 0; // fences
    if constexpr ((FLAGS & F2D_SMEM_TW) != 0) {
        cpx *tws = smem + f2d_xch_elems(NX, NY, VX) + 2; // behind the exchange area (+16 bytes)
        f2d_stage_twiddles<NX, NY, VX>(tws, twx, twy, tid, MUFFT_F2D_THREADS);
        twx = tws;
        twy = (NX != NY || VX != 0) ? tws + engine_tw_total(NX, VX) : tws;
    }
    unsigned *ticket_counter = a.counters;
    unsigned *rows_done = a.counters + 1;
    const unsigned long long total = (unsigned long long)a.batch * (NR + NC);

    // Column tickets of image i are issued `lag` images after its row tickets, far enough that the
    // row tiles (taken by other CTAs) have normally finished when a column ticket is drawn.
    const unsigned lag = a.lag < a.batch ? a.lag : (unsigned)a.batch;
    const unsigned long long head = (unsigned long long)lag * NR;                  // rows only
    const unsigned long long body = (unsigned long long)(a.batch - lag) * (NR + NC); // rows + lagged columns

    // ticket -> (image, phase, tile)
    auto decode = [&](unsigned k, unsigned &image, unsigned &tile) -> bool {
        if (k < head) {
            image = k / NR; tile = k % NR;
            return true;
        } else if (k - head < body) {
            const unsigned long long kk = k - head;
            const unsigned s = lag + (unsigned)(kk / (NR + NC)), j = (unsigned)(kk % (NR + NC));
            if constexpr ((FLAGS & F2D_INTERLEAVE) != 0) {
                // r(j) = floor(j NR / (NR + NC)) row tickets precede position j; position j is a row ticket iff r(j + 1) > r(j)
                const unsigned r0 = (unsigned)((unsigned long long)j * NR / (NR + NC)), r1 = (unsigned)((unsigned long long)(j + 1) * NR / (NR + NC));
                if (r1 > r0) { image = s; tile = r0; return true; }
                image = s - lag; tile = j - r0;
                return false;
            }
            if (j < NR) { image = s; tile = j; return true; }
            image = s - lag; tile = j - NR;
            return false;
        }
        const unsigned long long kk = k - head - body;
        image = (unsigned)a.batch - lag + (unsigned)(kk / NC); tile = (unsigned)(kk % NC);
        return false;
    };
    // Rows of a finished row tile are announced LATE: a release needs the tile's stores to have
    // reached L2 (fence), and waiting for that right after issuing them would put a full store
    // round trip on the critical path of every row tile.  Thread 0 remembers the image instead
    // (`pending`) and releases it while the NEXT tile's loads are in flight, when the stores have
    // long landed.  (All threads' stores precede the CTA barrier at the top of the next
    // iteration, thread 0's fence follows it: fence cumulativity covers them.)
    unsigned pending = ~0u; // thread 0 only
    auto release_pending = [&]() {
        if (pending != ~0u) {
            f2d_fence_release<LITE>();
            red_relaxed_add(rows_done + pending, (unsigned)(ROWS * REP));
            pending = ~0u;
        }
    };
    auto wait_rows = [&](unsigned image) {
        unsigned ns = 32;
        while (ld_relaxed(rows_done + image) < (unsigned)NY) {
            __nanosleep(ns);
            if (ns < 1024) ns *= 2;
        }
    };
    // Thread 0 publishes a ticket to the CTA once its dependency is met (a column ticket needs all
    // row tickets of its image), so one CTA barrier per ticket is enough.  The exception: the
    // dependency is not met yet AND this CTA still holds back the rows of the tile it has just
    // stored (they may be the very rows it would wait for, and they cannot be released before the
    // next CTA barrier has covered every thread's stores).  Then the ticket is published with the
    // wait flag set and the wait happens after that barrier.
    // `seen`: a value of the image's rows-done counter read earlier (0 = none)
    auto publish = [&](unsigned k, int slot, unsigned seen = 0) {
        unsigned image, tile, wait = 0;
        if (k < total && !decode(k, image, tile)) {
            if (seen < (unsigned)NY && ld_relaxed(rows_done + image) < (unsigned)NY) {
                if (pending != ~0u) wait = 1;
                else wait_rows(image);
            }
            f2d_fence_acquire<LITE>();
        }
        s_ticket[slot] = k;
        s_wait[slot] = wait;
    };

    unsigned next_ticket = 0;
    unsigned ahead = 0, ahead_seen = 0; // F2D_EARLY, thread 0: the ticket after the current one and what is known of its dependency
    if (tid == 0) {
        publish(atomicAdd(ticket_counter, 1u), 0);
        if constexpr ((FLAGS & F2D_EARLY) != 0) ahead = atomicAdd(ticket_counter, 1u);
    }
    for (int it = 0;; it++) {
        __syncthreads(); // ticket visible; the previous ticket's users of the exchange area are done
        const unsigned k = s_ticket[it & 1];
        if (k >= total) {
            if (tid == 0) release_pending();
            break;
        }
        if (s_wait[it & 1]) { // rare: see publish
            if (tid == 0) {
                unsigned image, tile;
                decode(k, image, tile);
                release_pending();
                wait_rows(image);
                f2d_fence_acquire<LITE>();
            }
            __syncthreads();
        }
        // draw the next ticket now; its round trip to L2 overlaps this tile
        if constexpr ((FLAGS & F2D_EARLY) != 0) {
            if (tid == 0) {
                // `ahead` was drawn a tile ago: look at its dependency now, use the answer after this tile
                unsigned im, tl;
                ahead_seen = (unsigned)NY;
                if (ahead < total && !decode(ahead, im, tl)) ahead_seen = ld_relaxed(rows_done + im);
                next_ticket = atomicAdd(ticket_counter, 1u);
            }
        } else {
            if (tid == 0) next_ticket = atomicAdd(ticket_counter, 1u);
        }
        unsigned image, tile;
        const bool is_rows = decode(k, image, tile);

        cpx *__restrict__ out = static_cast<cpx *>(a.out) + (size_t)image * a.out_image;
#ifdef MUFFT_F2D_DIAG
        // diagnostic builds (tools/ab.sh): 1 = row tiles only, 2 = column tiles only (results are wrong)
        if ((MUFFT_F2D_DIAG == 1 && !is_rows) || (MUFFT_F2D_DIAG == 2 && is_rows)) {
            if (is_rows && tid == 0) { release_pending(); pending = image; release_pending(); }
            if (tid == 0) publish(next_ticket, (it + 1) & 1);
            continue;
        }
#endif
        if constexpr (OM != MCU_OUT_COMPLEX) {
          if (is_rows) {
            // real rows: NX complex = 2 NX floats in, NX + 1 (2 NX) bins out at pitch 2 NX
            const int slot = tid / TX, t = tid - slot * TX;
            const size_t row = (size_t)tile * ROWS + slot;
            const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + (size_t)image * a.in_image + row * NX;
            if (tid == 0 && image + 1 < a.batch && (reinterpret_cast<uintptr_t>(in + a.in_image) & 15) == 0)
                prefetch_l2_bulk(in + a.in_image, (unsigned)(ROWS * NX * sizeof(cpx)));
            cpx v[RX];
#pragma unroll
            for (int e = 0; e < RX; e++) v[e] = __ldcs(in + t + e * TX);
            if (tid == 0) release_pending();
            cpx *sm = smem + slot * engine_smem_elems(NX, VX);
            engine_run<NX, DIR, 1, TX, ROWS, 0, VX, -1, f2d_twm(FLAGS)>(v, sm, twx, t, slot);
            rows_store<NX, OM, TX, ROWS, VX>(v, out + row * PITCH, sm, static_cast<const cpx *>(a.rtw), t, slot, true);
            if (tid == 0) pending = image;
          } else {
            const int c = tid % COLS, t = tid / COLS;
            const bool active = tile * COLS + c < (unsigned)WIDTH;
            cpx *__restrict__ col = out + (size_t)tile * COLS + c;
            cpx v[RY];
#pragma unroll
            for (int e = 0; e < RY; e++) v[e] = active ? __ldcg(col + (size_t)(t + e * TY) * PITCH) : make_float2(0.0f, 0.0f);
            if (tid == 0) release_pending();
            engine_run<NY, DIR, COLS, MUFFT_F2D_THREADS, 1, 0, 0, -1, f2d_twm(FLAGS)>(v, smem + c, twy, t, 0);
            if (active) {
#pragma unroll
                for (int e = 0; e < RY; e++) __stcs(col + (size_t)(t + e * TY) * PITCH, v[e]);
            }
          }
        } else
        if (is_rows) {
            const int slot = tid / TX, t = tid - slot * TX;
            // pull the same rows of the next image into L2 (one bulk prefetch of the ticket's
            // contiguous REP*ROWS*NX*8 bytes): that ticket is about one wave of tickets ahead
            if (tid == 0 && image + 1 < a.batch) {
                const cpx *nxt = static_cast<const cpx *>(a.in) + (size_t)(image + 1) * a.in_image + (size_t)tile * (REP * ROWS) * NX;
                if ((reinterpret_cast<uintptr_t>(nxt) & 15) == 0)
                    prefetch_l2_bulk(nxt, (unsigned)(REP * ROWS * NX * sizeof(cpx)));
            }
#pragma unroll 1
            for (int rep = 0; rep < REP; rep++) {
                const size_t row = ((size_t)tile * REP + rep) * ROWS + slot;
                const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + (size_t)image * a.in_image + row * NX;
                cpx v[RX];
#pragma unroll
                for (int e = 0; e < RX; e++) v[e] = __ldcs(in + t + e * TX); // streamed: read once
                if (rep == 0 && tid == 0) release_pending(); // the previous row ticket, while this one's loads fly
                engine_run<NX, DIR, 1, TX, ROWS, 0, 0, -1, f2d_twm(FLAGS)>(v, smem + slot * engine_smem_elems(NX), twx, t, slot);
                cpx *__restrict__ o = out + row * NX;
#pragma unroll
                for (int e = 0; e < RX; e++) o[t + e * TX] = v[e];
                if (REP > 1) engine_sync<TX, ROWS>(slot); // the exchange area is reused by the next row
            }
            if (tid == 0) pending = image;
        } else {
            const int c = tid % COLS, t = tid / COLS;
            cpx *__restrict__ col = out + (size_t)tile * COLS + c;
            cpx v[RY];
#pragma unroll
            for (int e = 0; e < RY; e++) v[e] = __ldcg(col + (size_t)(t + e * TY) * NX); // written by other SMs: L2 only
            if (tid == 0) release_pending();
            engine_run<NY, DIR, COLS, MUFFT_F2D_THREADS, 1, 0, 0, -1, f2d_twm(FLAGS)>(v, smem + c, twy, t, 0);
#pragma unroll
            for (int e = 0; e < RY; e++) __stcs(col + (size_t)(t + e * TY) * NX, v[e]);
        }
        if constexpr ((FLAGS & F2D_EARLY) != 0) {
            if (tid == 0) {
                publish(ahead, (it + 1) & 1, ahead_seen);
                ahead = next_ticket;
            }
        } else {
            if (tid == 0) publish(next_ticket, (it + 1) & 1);
        }
    }
}

} // namespace mufft
