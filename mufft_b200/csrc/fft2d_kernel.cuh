// Whole batched 2D transforms in ONE persistent kernel.  CTAs draw tickets from a global counter; a
// ticket is a tile of rows or a tile of columns of one image.  Every image has a PRODUCER pass and a
// CONSUMER pass:
//
//   kind            producer tiles                                consumer tiles
//   C2C             rows: input -> output buffer                  columns of the output buffer, in place
//   R2C [FULL]      real rows + resolve -> output (pitch 2 NX)    the NX + 1 (2 NX) columns that carry data, in place
//   C2R             columns 0..NX of the input -> ring slot       rows of the ring slot: C2R resolve + inverse
//                   (plan-owned, pitch NX + 2)                    transform -> dense real output
//
// (the reference's order: rows then columns, and columns first for C2R - fft.c:852-989).  Tickets are ordered
//     P(0) .. P(lag-1), P(lag), C(0), P(lag+1), C(1), ... , C(B-1)
// so the consumer pass of an image runs `lag` images behind its producer pass, while the intermediate
// (8 MiB for 1024 x 1024) is still resident in the 126 MB L2: HBM sees the input once and the output once.
//
// A consumer ticket of image i waits (thread 0 spins on a relaxed load, with back-off, then an acquire
// fence) until all producer tiles of image i have signalled completion.  For C2R the intermediate lives in
// a ring of `ring` > lag image slots, and a producer ticket of image i also waits until the consumers of
// image i - ring are done with the slot.  Tickets are handed out in increasing order and every wait is on
// tickets that were handed out earlier, i.e. that are running or finished: no deadlock whatever the number
// of resident CTAs.
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

// Kernel flavour per shape, from the measurements of profiles/README.md (round-1 flavour table, round-2
// occupancy probe).  ncu had shown the twiddle loads of the plain kernel pulling as many sectors from L2 as
// a whole pass over the data (the acquire fence of the ticket protocol is CCTL.IVALL and empties the SM's
// L1), hence either
//   F2D_SMEM_TW  twiddle tables staged in shared memory once per CTA, or
//   F2D_DERIVE   only the log2(r) power-of-two twiddles of a radix-r butterfly are loaded, the rest derived.
// Sides of 256 / 512 points run with 16 values per thread and fit THREE CTAs per SM without spills
// (-7 .. -10 % against two); sides of 1024 / 2048 points need 104-127 registers: two CTAs per SM.
#define F2D_SMEM_TW 1
#define F2D_DERIVE 4
__host__ __device__ constexpr bool f2d_small(int nx, int ny) { return engine_regs(nx) <= 16 && engine_regs(ny) <= 16; }
__host__ __device__ constexpr int f2d_flavour(int nx, int ny)
{
#ifdef MUFFT_F2D_FLAVOUR
    return MUFFT_F2D_FLAVOUR;
#else
    return (nx == ny && nx == 1024) ? (F2D_SMEM_TW | F2D_DERIVE) : (nx == ny && nx == 512) ? F2D_SMEM_TW : F2D_DERIVE; // profiles/r02q_ab.txt
#endif
}
__host__ __device__ constexpr int f2d_min_blocks(int nx, int ny)
{
#ifdef MUFFT_F2D_MIN_BLOCKS
    return MUFFT_F2D_MIN_BLOCKS;
#else
    return f2d_small(nx, ny) ? 3 : 2;
#endif
}
__host__ __device__ constexpr int f2d_twm(int flags) { return ((flags & F2D_SMEM_TW) ? MUFFT_TW_SMEM : 0) | ((flags & F2D_DERIVE) ? MUFFT_TW_DERIVE : 0); }

enum { F2D_C2C = 0, F2D_R2C = 1, F2D_R2C_FULL = 2, F2D_C2R = 3 };
__host__ __device__ constexpr int f2d_out_mode(int kind) { return kind == F2D_R2C ? MCU_OUT_R2C : (kind == F2D_R2C_FULL ? MCU_OUT_R2C_FULL : MCU_OUT_COMPLEX); }
__host__ __device__ constexpr int f2d_in_mode(int kind) { return kind == F2D_C2R ? MCU_IN_C2R : MCU_IN_COMPLEX; }
// Engine shape of the row transforms: the mirrored-pass variants of the real transforms when they keep the
// tile shape of the complex kernel (same number of threads per row), else the general shape.
__host__ __device__ constexpr int f2d_rows_variant(int nx, int kind)
{
    const int v = rows_variant(nx, f2d_in_mode(kind), f2d_out_mode(kind));
    return engine_threads(nx, v) == engine_threads(nx) ? v : MUFFT_VARIANT_GENERAL;
}
__host__ __device__ constexpr int f2d_rows_per_tile(int nx) { return MUFFT_F2D_THREADS / engine_threads(nx); }
__host__ __device__ constexpr int f2d_cols_per_tile(int ny) { return MUFFT_F2D_THREADS / engine_threads(ny); }
__host__ __device__ constexpr bool f2d_supported(int nx, int ny)
{
    return engine_regs(nx) > 0 && engine_regs(ny) > 0 && engine_threads(nx) <= MUFFT_F2D_THREADS &&
           engine_threads(ny) * 4 <= MUFFT_F2D_THREADS && (engine_threads(nx) <= 32 || f2d_rows_per_tile(nx) <= 15) &&
           ny % f2d_rows_per_tile(nx) == 0 && nx % f2d_cols_per_tile(ny) == 0;
}
// Row pitch (complex elements) of the intermediate / output image a kind works on, columns that carry data.
__host__ __device__ constexpr int f2d_pitch(int nx, int kind) { return kind == F2D_C2C ? nx : (kind == F2D_C2R ? nx + 2 : 2 * nx); }
__host__ __device__ constexpr int f2d_width(int nx, int kind) { return kind == F2D_C2C ? nx : (kind == F2D_R2C_FULL ? 2 * nx : nx + 1); }
// (Several row groups per row ticket for short rows - 8 rows of 256 points are little work per ticket - were measured in
// round 2: 2 groups -4..+2 %, 4 groups -20 %: profiles/r02h_ab.txt.  One group per ticket it stays.)
__host__ __device__ constexpr unsigned f2d_row_tiles(int nx, int ny) { return ny / f2d_rows_per_tile(nx); }
__host__ __device__ constexpr unsigned f2d_col_tiles(int nx, int ny, int kind) { return (f2d_width(nx, kind) + f2d_cols_per_tile(ny) - 1) / f2d_cols_per_tile(ny); }
__host__ __device__ constexpr unsigned f2d_tickets_per_image(int nx, int ny, int kind) { return f2d_row_tiles(nx, ny) + f2d_col_tiles(nx, ny, kind); }
__host__ __device__ constexpr size_t f2d_xch_elems(int nx, int ny, int vx)
{
    size_t r = (size_t)f2d_rows_per_tile(nx) * engine_smem_elems(nx, vx);
    size_t c = (size_t)f2d_cols_per_tile(ny) * engine_smem_elems(ny);
    return r > c ? r : c;
}
// Twiddle tables of both axes in shared memory (one table when the sides are equal and the rows use the
// general engine shape), in cpx.
__host__ __device__ constexpr int f2d_tw_elems(int nx, int ny, int vx) { return engine_tw_total(nx, vx) + ((nx != ny || vx != 0) ? engine_tw_total(ny) : 0); }
__host__ __device__ constexpr size_t f2d_smem_bytes(int nx, int ny, int kind)
{
    const int vx = f2d_rows_variant(nx, kind);
    return f2d_xch_elems(nx, ny, vx) * sizeof(float2) + 16 + ((f2d_flavour(nx, ny) & F2D_SMEM_TW) ? (size_t)f2d_tw_elems(nx, ny, vx) * sizeof(float2) : 0);
}
// Complex elements of one ring slot (C2R intermediate).
__host__ __device__ constexpr size_t f2d_mid_elems(int nx, int ny) { return (size_t)f2d_pitch(nx, F2D_C2R) * ny; }

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
__device__ __forceinline__ void red_relaxed_add(unsigned *p, unsigned v)
{
#ifdef MUFFT_EMU
    __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
#else
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
// One-sided fences.  fence.acq_rel.gpu is MEMBAR.ALL.GPU + CCTL.IVALL in SASS: the second half throws
// away the SM's whole L1, which a RELEASE does not need, and the first half waits for this thread's
// outstanding writes, which an ACQUIRE does not need (measured: 1.279 -> 1.233 ms for 1024 x 1024 x 256).
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
// Pull `bytes` (a multiple of 16, 16-byte aligned) into L2 ahead of use.
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned bytes)
{
#ifdef MUFFT_EMU
    (void)p; (void)bytes;
#else
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#endif
}
// Reads that must come from L2: data written by other SMs during this launch.
struct cg_ptr {
    const cpx *p;
    __device__ __forceinline__ cpx operator[](long long i) const { return __ldcg(p + i); }
};
struct cs_ptr { // streamed input: read once
    const cpx *p;
    __device__ __forceinline__ cpx operator[](long long i) const { return __ldcs(p + i); }
};

template <int NX, int NY, int DIR, int KIND>
__global__ void __launch_bounds__(MUFFT_F2D_THREADS, f2d_min_blocks(NX, NY))
fft2d_fused_kernel(const mcu_2d_args a)
{
    constexpr int FLAGS = f2d_flavour(NX, NY);
    constexpr int TWM = f2d_twm(FLAGS);
    constexpr int VX = f2d_rows_variant(NX, KIND);
    constexpr int RX = engine_regs(NX, VX), TX = NX / RX, ROWS = f2d_rows_per_tile(NX);
    constexpr int RY = engine_regs(NY), TY = NY / RY, COLS = f2d_cols_per_tile(NY);
    constexpr int PITCH = f2d_pitch(NX, KIND);   // rows of the image the column tiles work on
    constexpr int WIDTH = f2d_width(NX, KIND);   // columns to transform
    constexpr bool COLS_FIRST = KIND == F2D_C2R;
    constexpr int REP = 1; // row groups per row ticket
    constexpr unsigned NR = f2d_row_tiles(NX, NY), NC = f2d_col_tiles(NX, NY, KIND);
    constexpr unsigned NP = COLS_FIRST ? NC : NR; // producer / consumer tiles per image
    constexpr unsigned NQ = COLS_FIRST ? NR : NC;
    static_assert(MUFFT_F2D_THREADS / TX == ROWS, "the real row variants keep the tile shape");
    static_assert(KIND == F2D_C2C || (KIND == F2D_C2R ? DIR > 0 : DIR < 0), "R2C is a forward, C2R an inverse transform");
    MUFFT_DYNAMIC_SMEM(smem_raw);
    MUFFT_STATIC_SMEM_U32(s_ticket, 2);
    MUFFT_STATIC_SMEM_U32(s_wait, 2);
    cpx *smem = reinterpret_cast<cpx *>(smem_raw);

    const int tid = threadIdx.x;
    const cpx *__restrict__ twx = static_cast<const cpx *>(a.tw_x);
    const cpx *__restrict__ twy = static_cast<const cpx *>(a.tw_y);
    if constexpr ((FLAGS & F2D_SMEM_TW) != 0) {
        cpx *tws = smem + f2d_xch_elems(NX, NY, VX) + 2; // behind the exchange area (+16 bytes); the first ticket barrier publishes it
        for (int i = tid; i < engine_tw_total(NX, VX); i += MUFFT_F2D_THREADS) tws[i] = __ldg(twx + i);
        if constexpr (NX != NY || VX != 0)
            for (int i = tid; i < engine_tw_total(NY); i += MUFFT_F2D_THREADS) tws[engine_tw_total(NX, VX) + i] = __ldg(twy + i);
        twx = tws;
        twy = (NX != NY || VX != 0) ? tws + engine_tw_total(NX, VX) : tws;
    }
    unsigned *ticket_counter = a.counters;
    unsigned *produced = a.counters + 1;            // [batch] producer tiles finished per image
    unsigned *consumed = a.counters + 1 + a.batch;  // [batch] consumer tiles finished per image (C2R: frees the ring slot)
    const unsigned long long total = (unsigned long long)a.batch * (NP + NQ);
    const unsigned lag = a.lag < a.batch ? a.lag : (unsigned)a.batch;
    const unsigned long long head = (unsigned long long)lag * NP;                    // producers only
    const unsigned long long body = (unsigned long long)(a.batch - lag) * (NP + NQ); // producers + lagged consumers

    // ticket -> (image, tile); true = producer tile
    auto decode = [&](unsigned k, unsigned &image, unsigned &tile) -> bool {
        if (k < head) {
            image = k / NP; tile = k % NP;
            return true;
        } else if (k - head < body) {
            const unsigned long long kk = k - head;
            const unsigned s = lag + (unsigned)(kk / (NP + NQ)), j = (unsigned)(kk % (NP + NQ));
            if (j < NP) { image = s; tile = j; return true; }
            image = s - lag; tile = j - NP;
            return false;
        }
        const unsigned long long kk = k - head - body;
        image = (unsigned)a.batch - lag + (unsigned)(kk / NQ); tile = (unsigned)(kk % NQ);
        return false;
    };
    // A finished tile is announced LATE: a release needs the tile's stores to have reached L2 (fence), and
    // waiting for that right after issuing them would put a full store round trip on the critical path of
    // every tile.  Thread 0 remembers the counter instead (`pending`) and releases it while the NEXT tile's
    // loads are in flight, when the stores have long landed.  (All threads' accesses precede the CTA
    // barrier at the top of the next iteration, thread 0's fence follows it: fence cumulativity covers them.)
    unsigned *pending = nullptr; // thread 0 only
    auto release_pending = [&]() {
        if (pending != nullptr) {
            fence_release_gpu();
            red_relaxed_add(pending, 1u);
            pending = nullptr;
        }
    };
    // What ticket k has to wait for: a consumer for the producers of its image; a C2R producer for the
    // consumers of the image that used its ring slot before.  nullptr = nothing.
    // `reads` = the ticket will read what the counter announces (acquire needed); a producer waiting for its
    // ring slot only overwrites it (write-after-read: the wait alone orders it, and the acquire fence would
    // throw away the SM's L1 - twiddle tables included - for nothing).
    auto dependency = [&](unsigned k, unsigned &need, bool &reads) -> const unsigned * {
        unsigned image, tile;
        if (!decode(k, image, tile)) { need = NP; reads = true; return produced + image; }
        reads = false;
        if (COLS_FIRST && image >= a.ring) { need = NQ; return consumed + (image - a.ring); }
        return nullptr;
    };
    auto wait_for = [&](const unsigned *counter, unsigned need) {
        unsigned ns = 32;
        while (ld_relaxed(counter) < need) {
            __nanosleep(ns);
            if (ns < 1024) ns *= 2;
        }
    };
    // Thread 0 publishes a ticket to the CTA once its dependency is met, so one CTA barrier per ticket is
    // enough.  The exception: the dependency is not met yet AND this CTA still holds back the announcement of
    // the tile it has just finished (it may be the very tile it would wait for, and it cannot be released
    // before the next CTA barrier has covered every thread's stores).  Then the ticket is published with the
    // wait flag set and the wait happens after that barrier.
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

    unsigned next_ticket = 0;
    if (tid == 0) publish(atomicAdd(ticket_counter, 1u), 0);
    for (int it = 0;; it++) {
        __syncthreads(); // ticket visible; the previous ticket's users of the exchange area are done
        const unsigned k = s_ticket[it & 1];
        if (k >= total) {
            if (tid == 0) release_pending();
            break;
        }
        if (s_wait[it & 1]) { // rare: see publish
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
        if (tid == 0) next_ticket = atomicAdd(ticket_counter, 1u); // its round trip to L2 overlaps this tile
        unsigned image, tile;
        const bool producer = decode(k, image, tile);
        const bool is_rows = producer != COLS_FIRST;

        if (is_rows) {
            const int slot = tid / TX, t = tid - slot * TX;
            cpx *sm = smem + slot * engine_smem_elems(NX, VX);
            if constexpr (KIND != F2D_C2R) {
                // pull the same rows of the next image into L2: that ticket is about one wave of tickets ahead
                if (tid == 0 && image + 1 < a.batch) {
                    const cpx *nxt = static_cast<const cpx *>(a.in) + (size_t)(image + 1) * a.in_image + (size_t)tile * (REP * ROWS) * NX;
                    if ((reinterpret_cast<uintptr_t>(nxt) & 15) == 0) prefetch_l2_bulk(nxt, (unsigned)(REP * ROWS * NX * sizeof(cpx)));
                }
            }
#pragma unroll 1
            for (int rep = 0; rep < REP; rep++) {
                const size_t row = ((size_t)tile * REP + rep) * ROWS + slot;
                cpx v[RX];
                if constexpr (KIND == F2D_C2R) {
                    // consumer: ring slot (written by other SMs: L2 reads) -> dense real rows of NX complex = 2 NX floats
                    const cpx *mid = static_cast<const cpx *>(a.mid) + (size_t)(image % a.ring) * f2d_mid_elems(NX, NY) + row * PITCH;
                    rows_load<NX, MCU_IN_C2R, VX>(v, cg_ptr{ mid }, static_cast<const cpx *>(nullptr), static_cast<const cpx *>(a.rtw), 1.0f, t);
                    if (rep == 0 && tid == 0) release_pending();
                    engine_run<NX, DIR, 1, TX, ROWS, 0, VX, -1, TWM>(v, sm, twx, t, slot);
                    cpx *__restrict__ o = static_cast<cpx *>(a.out) + (size_t)image * a.out_image + row * NX;
#pragma unroll
                    for (int e = 0; e < RX; e++) __stcs(o + t + e * TX, v[e]);
                } else {
                    // producer: input rows (streamed, read once) -> rows of the output buffer
                    const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + (size_t)image * a.in_image + row * NX;
#pragma unroll
                    for (int e = 0; e < RX; e++) v[e] = __ldcs(in + t + e * TX);
                    if (rep == 0 && tid == 0) release_pending(); // the previous tile, while this one's loads fly
                    engine_run<NX, DIR, 1, TX, ROWS, 0, VX, -1, TWM>(v, sm, twx, t, slot);
                    cpx *__restrict__ o = static_cast<cpx *>(a.out) + (size_t)image * a.out_image + row * PITCH;
                    rows_store<NX, f2d_out_mode(KIND), TX, ROWS, VX>(v, o, sm, static_cast<const cpx *>(a.rtw), t, slot, true);
                }
                if constexpr (REP > 1) engine_sync<TX, ROWS>(slot); // the exchange area is reused by the next row of this group
            }
            if (tid == 0) pending = (KIND == F2D_C2R ? consumed : produced) + image;
        } else {
            const int c = tid % COLS, t = tid / COLS;
            const bool active = WIDTH % COLS == 0 || tile * COLS + c < (unsigned)WIDTH;
            cpx v[RY];
            if constexpr (KIND == F2D_C2R) {
                // producer: columns 0..NX of the input spectrum (pitch 2 NX, streamed) -> ring slot
                const cpx *__restrict__ src = static_cast<const cpx *>(a.in) + (size_t)image * a.in_image + (size_t)tile * COLS + c;
#pragma unroll
                for (int e = 0; e < RY; e++) v[e] = active ? __ldcs(src + (size_t)(t + e * TY) * (2 * NX)) : make_float2(0.0f, 0.0f);
                if (tid == 0) release_pending();
                engine_run<NY, DIR, COLS, MUFFT_F2D_THREADS, 1, 0, 0, -1, TWM>(v, smem + c, twy, t, 0);
                cpx *__restrict__ dst = static_cast<cpx *>(a.mid) + (size_t)(image % a.ring) * f2d_mid_elems(NX, NY) + (size_t)tile * COLS + c;
                if (active) {
#pragma unroll
                    for (int e = 0; e < RY; e++) dst[(size_t)(t + e * TY) * PITCH] = v[e];
                }
                if (tid == 0) pending = produced + image;
            } else {
                // consumer: columns of the output buffer in place (written by other SMs: L2 reads)
                cpx *__restrict__ col = static_cast<cpx *>(a.out) + (size_t)image * a.out_image + (size_t)tile * COLS + c;
#pragma unroll
                for (int e = 0; e < RY; e++) v[e] = active ? __ldcg(col + (size_t)(t + e * TY) * PITCH) : make_float2(0.0f, 0.0f);
                if (tid == 0) release_pending();
                engine_run<NY, DIR, COLS, MUFFT_F2D_THREADS, 1, 0, 0, -1, TWM>(v, smem + c, twy, t, 0);
                if (active) {
#pragma unroll
                    for (int e = 0; e < RY; e++) __stcs(col + (size_t)(t + e * TY) * PITCH, v[e]);
                }
            }
        }
        if (tid == 0) publish(next_ticket, (it + 1) & 1);
    }
}

// (A pipelined variant - tickets drawn two ahead and published a barrier early, so that every warp issues the next tile's
// loads before it goes to the ticket barrier - was built and measured in round 2 as well: 1.303 ms against 1.213 ms for
// 1024 x 1024 x 256; ncu shows the load stalls (16 % -> 9 % of the samples) simply turning into barrier stalls (17 % -> 26 %):
// profiles/r02n_ab.txt, r02n_c2c2d1024_pipe_rejected.txt.  Gone too.)
// (A split variant - row groups and a column half on the two halves of a 512-thread CTA, static work sequences, no
// tickets and no CTA-wide barrier - was built and measured in round 2: 1.402 ms against 1.210 ms for 1024 x 1024 x 256,
// profiles/r02e_ab.txt and r02e_c2c2d1024_split_rejected.txt.  Desynchronising the two passes does not help when the
// SM still holds 16 warps: it is gone.)

} // namespace mufft
