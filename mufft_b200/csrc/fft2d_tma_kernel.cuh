// Fused batched 2D C2C transform, second generation: the ticket scheme of fft2d_kernel.cuh (row
// tiles, then column tiles of the same image a few images later, the intermediate staying in L2),
// with every tile's input brought into shared memory by the TMA unit while the PREVIOUS tile is
// still being transformed:
//   * a row tile (ROWS consecutive rows) is one contiguous bulk copy (cp.async.bulk);
//   * a column tile (COLS adjacent columns x NY rows) is NY/256 boxes of a 2D tensor map over the
//     output buffer (cp.async.bulk.tensor.2d): the strided gather that costs the LSU four
//     64-byte segments per warp instruction in fft2d_kernel.cuh is done by the copy engine.
// Loads therefore cost the warps one shared-memory read instead of a trip to L2 / HBM, and the
// latency of that trip overlaps the butterflies of the tile before.  ncu on the first-generation
// kernel (profiles/r01c_c2c2d1024_full.txt) shows why that matters: 16 warps per SM, 43 % issue
// utilisation, the top stalls are the barrier behind the loads and long-scoreboard waits.
//
// Shared memory per CTA: one landing buffer (a whole tile, 64 KiB for 1024-point sides) and the
// exchange area at HALF size - the engine exchanges real and imaginary planes one after the other
// (engine_run_split) - so that two CTAs still fit an SM (2 x 98 KiB).
//
// Per tile, thread 0 runs one step ahead of the CTA: while the CTA transforms tile k it has
// already drawn ticket k+1, checked its dependency (a column tile needs all rows of its image) and
// started the copy into the landing buffer, which every thread finished reading just before.
//
// Replaces, like fft2d_kernel.cuh, the reference's 2D driver (fft.c:852-989) and its vertical
// passes (kernel.c:455-738).
#pragma once

#include "abi.h"
#include "fft2d_kernel.cuh"
#include "fft_engine.cuh"
#include "tma.cuh"

namespace mufft {

__host__ __device__ constexpr int f2t_box_rows(int ny) { return ny < 256 ? ny : 256; } // tensor-map boxes are at most 256 rows
__host__ __device__ constexpr size_t f2t_land_bytes(int nx, int ny)
{
    size_t r = (size_t)f2d_rows_per_tile(nx) * nx, c = (size_t)f2d_cols_per_tile(ny) * ny;
    return (r > c ? r : c) * sizeof(float2);
}
__host__ __device__ constexpr size_t f2t_xch_floats(int nx, int ny)
{
    size_t r = (size_t)f2d_rows_per_tile(nx) * engine_smem_elems(nx), c = (size_t)f2d_cols_per_tile(ny) * engine_smem_elems(ny);
    return r > c ? r : c;
}
// [mbarrier, padded to 128 bytes][landing buffer][exchange area, floats][LITE: twiddle tables]
__host__ __device__ constexpr size_t f2t_smem_bytes(int nx, int ny, bool lite = false)
{
    return 128 + f2t_land_bytes(nx, ny) + ((f2t_xch_floats(nx, ny) + 1) & ~(size_t)1) * sizeof(float) + (lite ? (size_t)f2d_tw_elems(nx, ny) * sizeof(float2) : 0);
}
__host__ __device__ constexpr bool f2t_supported(int nx, int ny)
{
    return f2d_supported(nx, ny) && f2d_row_repeat(nx, ny) == 1 && ny % f2t_box_rows(ny) == 0 &&
           f2d_cols_per_tile(ny) * sizeof(float2) >= 16 && f2t_smem_bytes(nx, ny, true) <= 227 * 1024 &&
           (size_t)f2d_rows_per_tile(nx) * nx * sizeof(float2) < (1u << 20) && (size_t)f2d_cols_per_tile(ny) * ny * sizeof(float2) < (1u << 20);
}

// `tmap`: the OUTPUT buffer as a row-major [batch * NY][NX] array of 8-byte elements, box
// f2d_cols_per_tile(NY) x f2t_box_rows(NY).  Requirements checked by the launcher: input images
// 16-byte aligned, images tightly packed in the output.
// LITE: as in fft2d_fused_kernel - twiddle tables in shared memory, one-sided fences.
template <int NX, int NY, int DIR, bool LITE = false>
__global__ void __launch_bounds__(MUFFT_F2D_THREADS, MUFFT_F2D_MIN_BLOCKS)
fft2d_tma_kernel(const mcu_2d_args a, const MUFFT_GRID_CONSTANT tensor_map_t tmap)
{
    constexpr int RX = engine_regs(NX), TX = NX / RX, ROWS = f2d_rows_per_tile(NX);
    constexpr int RY = engine_regs(NY), TY = NY / RY, COLS = f2d_cols_per_tile(NY);
    constexpr unsigned NR = NY / ROWS; // row tickets per image
    constexpr unsigned NC = NX / COLS; // column tickets per image
    constexpr int BOXH = f2t_box_rows(NY);
    constexpr unsigned ROW_TILE_BYTES = ROWS * NX * sizeof(cpx), COL_TILE_BYTES = COLS * NY * sizeof(cpx);
    MUFFT_DYNAMIC_SMEM(smem_raw);
    MUFFT_STATIC_SMEM_U32(s_ticket, 2);
    MUFFT_STATIC_SMEM_U32(s_wait, 2);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    unsigned char *land_raw = smem_raw + 128;
    const cpx *land = reinterpret_cast<const cpx *>(land_raw);
    float *xch = reinterpret_cast<float *>(land_raw + f2t_land_bytes(NX, NY));

    const int tid = threadIdx.x;
    const cpx *__restrict__ twx = static_cast<const cpx *>(a.tw_x);
    const cpx *__restrict__ twy = static_cast<const cpx *>(a.tw_y);
    if constexpr (LITE) {
        cpx *tws = reinterpret_cast<cpx *>(xch + ((f2t_xch_floats(NX, NY) + 1) & ~(size_t)1));
        f2d_stage_twiddles<NX, NY>(tws, twx, twy, tid, MUFFT_F2D_THREADS);
        twx = tws;
        twy = NX != NY ? tws + engine_tw_total(NX) : tws;
    }
    unsigned *ticket_counter = a.counters;
    unsigned *rows_done = a.counters + 1;
    const unsigned long long total = (unsigned long long)a.batch * (NR + NC);
    const unsigned lag = a.lag < a.batch ? a.lag : (unsigned)a.batch;
    const unsigned long long head = (unsigned long long)lag * NR;
    const unsigned long long body = (unsigned long long)(a.batch - lag) * (NR + NC);

    // ticket -> (image, phase, tile); same order as fft2d_kernel.cuh
    auto decode = [&](unsigned k, unsigned &image, unsigned &tile) -> bool {
        if (k < head) {
            image = k / NR; tile = k % NR;
            return true;
        } else if (k - head < body) {
            const unsigned long long kk = k - head;
            const unsigned s = lag + (unsigned)(kk / (NR + NC)), j = (unsigned)(kk % (NR + NC));
            if (j < NR) { image = s; tile = j; return true; }
            image = s - lag; tile = j - NR;
            return false;
        }
        const unsigned long long kk = k - head - body;
        image = (unsigned)a.batch - lag + (unsigned)(kk / NC); tile = (unsigned)(kk % NC);
        return false;
    };
    // thread 0: rows of the tile stored in the previous iteration, announced after a CTA barrier has
    // covered every thread's stores (fence cumulativity), see fft2d_kernel.cuh
    unsigned pending = ~0u;
    auto release_pending = [&]() {
        if (pending != ~0u) {
            f2d_fence_release<LITE>();
            red_relaxed_add(rows_done + pending, (unsigned)ROWS);
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
    // thread 0: start the copy of ticket k's tile into the landing buffer (dependency already met)
    auto issue = [&](unsigned k) {
        unsigned image, tile;
        if (decode(k, image, tile)) {
            const cpx *src = static_cast<const cpx *>(a.in) + (size_t)image * a.in_image + (size_t)tile * ROWS * NX;
            tma_mbar_expect_tx(full, ROW_TILE_BYTES);
            tma_load_1d(land_raw, src, ROW_TILE_BYTES, full);
            // the same rows of the next image, into L2: that ticket is about one wave of tickets ahead
            if (image + 1 < a.batch) prefetch_l2_bulk(src + a.in_image, ROW_TILE_BYTES);
        } else {
            // the rows were written by other SMs through the generic proxy and acquired by this
            // thread (fence after the rows-done load); order them before the async-proxy reads
            tma_fence_generic_to_async();
            tma_mbar_expect_tx(full, COL_TILE_BYTES);
#pragma unroll
            for (int j = 0; j < NY / BOXH; j++)
                tma_load_2d(land_raw + (size_t)j * BOXH * COLS * sizeof(cpx), &tmap, (int)(tile * COLS), (int)(image * NY + j * BOXH), full);
        }
    };

    unsigned next_ticket = 0; // thread 0: drawn, not yet published
    if (tid == 0) {
        tma_mbar_init(full, 1);
        tma_fence_mbar_init();
        const unsigned k0 = atomicAdd(ticket_counter, 1u);
        if (k0 < total) {
            unsigned image, tile;
            if (!decode(k0, image, tile)) {
                wait_rows(image); // nothing of ours is pending: every wait is on tickets already running elsewhere
                f2d_fence_acquire<LITE>();
            }
            issue(k0);
        }
        s_ticket[0] = k0;
        s_wait[0] = 0;
        next_ticket = atomicAdd(ticket_counter, 1u);
    }
    uint32_t parity = 0;
    for (int it = 0;; it++) {
        __syncthreads(); // (A) ticket visible; every thread is done with the previous tile (exchange area, stores issued)
        const unsigned k = s_ticket[it & 1];
        if (k >= total) {
            if (tid == 0) release_pending();
            break;
        }
        unsigned image, tile;
        const bool is_rows = decode(k, image, tile);
        if (s_wait[it & 1] && tid == 0) {
            // rare: the rows this column tile needs were not complete when the ticket was published,
            // so its copy has not been started.  Everything this CTA stored is covered by (A).
            release_pending();
            wait_rows(image);
            f2d_fence_acquire<LITE>();
            issue(k);
        }
        tma_mbar_wait(full, parity);
        parity ^= 1;

        cpx *__restrict__ out = static_cast<cpx *>(a.out) + (size_t)image * a.out_image;
        if (is_rows) {
            const int slot = tid / TX, t = tid - slot * TX;
            cpx v[RX];
#pragma unroll
            for (int e = 0; e < RX; e++) v[e] = land[slot * NX + t + e * TX];
            tma_fence_reads_before_refill();
            __syncthreads(); // (B) the landing buffer has been read by everyone
            if (tid == 0) {
                release_pending(); // the row tile of the previous iteration
                const unsigned kn = next_ticket;
                unsigned wait = 0, image2, tile2;
                if (kn < total) {
                    if (!decode(kn, image2, tile2)) {
                        if (ld_relaxed(rows_done + image2) < (unsigned)NY) wait = 1;
                        else f2d_fence_acquire<LITE>();
                    }
                    if (!wait) issue(kn);
                }
                s_ticket[(it + 1) & 1] = kn;
                s_wait[(it + 1) & 1] = wait;
                next_ticket = atomicAdd(ticket_counter, 1u);
            }
            engine_run_split<NX, DIR, 1, TX, ROWS, 0, -1, LITE>(v, xch + slot * engine_smem_elems(NX), twx, t, slot);
            cpx *__restrict__ o = out + ((size_t)tile * ROWS + slot) * NX;
#pragma unroll
            for (int e = 0; e < RX; e++) o[t + e * TX] = v[e];
            if (tid == 0) pending = image;
        } else {
            const int c = tid % COLS, t = tid / COLS;
            cpx v[RY];
#pragma unroll
            for (int e = 0; e < RY; e++) v[e] = land[(t + e * TY) * COLS + c];
            tma_fence_reads_before_refill();
            __syncthreads(); // (B)
            if (tid == 0) {
                release_pending();
                const unsigned kn = next_ticket;
                unsigned wait = 0, image2, tile2;
                if (kn < total) {
                    if (!decode(kn, image2, tile2)) {
                        if (ld_relaxed(rows_done + image2) < (unsigned)NY) wait = 1;
                        else f2d_fence_acquire<LITE>();
                    }
                    if (!wait) issue(kn);
                }
                s_ticket[(it + 1) & 1] = kn;
                s_wait[(it + 1) & 1] = wait;
                next_ticket = atomicAdd(ticket_counter, 1u);
            }
            engine_run_split<NY, DIR, COLS, MUFFT_F2D_THREADS, 1, 0, -1, LITE>(v, xch + c, twy, t, 0);
            cpx *__restrict__ col = out + (size_t)tile * COLS + c;
#pragma unroll
            for (int e = 0; e < RY; e++) __stcs(col + (size_t)(t + e * TY) * NX, v[e]);
        }
    }
}

} // namespace mufft
