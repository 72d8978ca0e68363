// Batched 2D C2C transforms, warp-specialised: ONE persistent CTA per SM, tiles moved by the TMA unit.
//
// The ticket kernel of fft2d_kernel.cuh keeps a tile in REGISTERS from its loads to its stores: with 100 registers per
// thread an SM holds two CTAs, and ncu shows each of them computing a quarter of the time and waiting for its own loads,
// stores and barriers the rest (profiles/README.md R2.4).  Here the tile lives in SHARED memory from the moment the TMA
// unit is asked for it until the TMA unit has written it back, so loads, stores and butterflies of different tiles
// overlap inside one CTA:
//
//   loader   (1 thread)   draws tickets (same order as the ticket kernel: rows of image i, columns of image i - lag),
//                         waits for a column ticket's rows to be announced, waits for a free buffer, asks the TMA unit
//                         for the tile: 8 rows = 8 bulk copies of NX complex, 8 columns = boxes of 256 rows x 64 bytes
//                         through a tensor map of the output buffer; completion arrives on the buffer's `full` barrier
//   groups   (2 x 256)    wait for `full`, pull the tile into registers (natural order: element t + e*T in slot e),
//                         run the engine with the buffer itself as exchange area, put the result back in natural
//                         order, fence.proxy.async, arrive on `staged`
//   storers  (S threads)  one per buffer, each in a warp of its own (as lanes of ONE warp their polls and sleeps serialise:
//                         measured 1.49 ms instead of 1.06): wait for `staged`, hand the buffer to the TMA unit (bulk
//                         stores), wait until it has been READ (-> `freed`, the loader may refill it) and until the
//                         writes are COMPLETE (-> release + add on the image's counter: a finished row tile)
//
// S = 3 buffers of 66 KB for sides of 1024 points, 6 of 33 KB for 512 x 512.  Buffer of sequence number s (per CTA) is
// s % S, its group s % 2: while two tiles are being computed the other buffers are in flight from or to memory.  Every
// wait is on work with a smaller ticket number, as in the ticket kernel: no deadlock.  End of work: lcm(2, S) EXIT
// entries, so that every group and every storer sees the same number of them.
//
// L2 policies: the input is read once (evict_first), the row-transformed intermediate must survive in L2 until its
// columns are read `lag` images later (evict_last), the final store is streaming again (evict_first).  Without them
// ncu shows 6.3 GB of DRAM traffic for 4.29 GB of images (profiles/r02r_tma_nohints.txt); with them 4.31 GB.
//
// Measured (profiles/README.md R2.12): 1024 x 1024 x 256 in 1.056 ms against 1.176 ms for the ticket kernel.  With the
// groups doing NOTHING (MUFFT_TMA_DIAG=3: tiles are only moved in and out) the same launch takes 0.985 ms: the round trip
// of an intermediate of 4 x 8 MB through L2 is not free - the L2 keeps evict_last lines in a set-aside of ~24 MB, and
// tools/l2_ceiling_probe.cu shows a round trip costing 2-6 % of a copy below that size and 40 % above.
#pragma once

#include "fft2d_kernel.cuh"
#include "tma_util.cuh"

namespace mufft {

#define F2D_TMA_GROUPS 2
#define F2D_TMA_GROUP_THREADS 256
#define F2D_TMA_BOX_ROWS 256
#define F2D_TMA_MAX_STAGES 6
enum { F2D_TMA_ROWS = 0, F2D_TMA_COLS = 1, F2D_TMA_EXIT = 2 };

struct f2d_tma_params {
    mcu_2d_args a;
    alignas(64) CUtensorMap map; // output buffer as [batch * NY][2 * NX] floats, boxes of 16 floats x F2D_TMA_BOX_ROWS rows
};

// Sides whose transforms sit on 32 threads (one warp per row, 8 columns per group): 512 and 1024 points.
__host__ __device__ constexpr bool f2d_tma_supported(int nx, int ny)
{
    return engine_regs(nx) > 0 && engine_regs(ny) > 0 && nx / engine_regs(nx) == 32 && ny / engine_regs(ny) == 32 && ny % F2D_TMA_BOX_ROWS == 0;
}
// ... and where it is the faster kernel (profiles/r02aa_ab.txt, ms for 2 GiB of images, this kernel / ticket kernel):
// 1024 x 1024 1.056 / 1.175, 512 x 1024 1.261 / 1.333, 1024 x 512 1.276 / 1.381, but 512 x 512 1.487 / 1.339 (six buffers
// of 33 KB; four: 1.457) - the ticket kernel holds three CTAs per SM there.
__host__ __device__ constexpr bool f2d_tma_preferred(int nx, int ny)
{
#ifdef MUFFT_F2D_TMA_ALL
    return f2d_tma_supported(nx, ny);
#else
    return f2d_tma_supported(nx, ny) && (nx == 1024 || ny == 1024);
#endif
}
// Kinds: C2C, and R2C where the rows have the mirrored-pass engine shape (the resolve then happens in registers, on the way
// back into the buffer): 512 complex points per row, i.e. real images 1024 wide.
__host__ __device__ constexpr int f2d_tma_rows_variant(int nx, int kind) { return kind == F2D_R2C ? f2d_rows_variant(nx, kind) : 0; }
__host__ __device__ constexpr bool f2d_tma_kind_supported(int nx, int ny, int kind)
{
    if (!f2d_tma_supported(nx, ny)) return false;
    if (kind == F2D_C2C) return true;
    return kind == F2D_R2C && f2d_tma_rows_variant(nx, kind) == MUFFT_VARIANT_MIRROR_LAST && nx / engine_regs(nx, MUFFT_VARIANT_MIRROR_LAST) == 32;
}
// Row tiles of 64 KB whatever the row length, for C2C images: 8 rows of 1024 points, 16 rows (two per warp, one after the
// other) of 512.  Measured (profiles/r02au_ab_tma_64k_tiles.txt): 512 x 1024 1.258 -> 1.170 ms.  The same for the columns
// (16 columns of 512 points per tile, two per thread, 128-byte segments) LOSES, 1024 x 512 1.281 -> 1.372 ms, and so do 16-row
// tiles for R2C rows (0.672 -> 0.691 ms): columns and R2C rows stay at 8 per tile.  The code below is written for UY columns
// per thread all the same.
#ifndef MUFFT_F2D_TMA_UNITS
#define MUFFT_F2D_TMA_UNITS 2
#endif
__host__ __device__ constexpr int f2d_tma_units_x(int nx, int kind) { return nx <= 512 && kind == F2D_C2C ? MUFFT_F2D_TMA_UNITS : 1; }
__host__ __device__ constexpr int f2d_tma_units_y(int ny) { (void)ny; return 1; }
__host__ __device__ constexpr int f2d_tma_rows_per_tile(int nx, int kind) { return (F2D_TMA_GROUP_THREADS / 32) * f2d_tma_units_x(nx, kind); }
__host__ __device__ constexpr int f2d_tma_cols_per_tile(int ny) { return (F2D_TMA_GROUP_THREADS / 32) * f2d_tma_units_y(ny); }
__host__ __device__ constexpr unsigned f2d_tma_row_tiles(int nx, int ny, int kind) { return ny / f2d_tma_rows_per_tile(nx, kind); }
__host__ __device__ constexpr unsigned f2d_tma_col_tiles(int nx, int ny, int kind) { return (f2d_width(nx, kind) + f2d_tma_cols_per_tile(ny) - 1) / f2d_tma_cols_per_tile(ny); }
// A row's exchange area inside a buffer; its first nx (+ 1) elements hold the row in natural order (16-byte aligned for the bulk copies).
__host__ __device__ constexpr int f2d_tma_row_pitch(int nx, int kind = F2D_C2C)
{
    const int e = engine_smem_elems(nx, f2d_tma_rows_variant(nx, kind));
    return ((e > nx + 1 ? e : nx + 1) + 1) & ~1;
}
__host__ __device__ constexpr size_t f2d_tma_buf_bytes(int nx, int ny, int kind = F2D_C2C)
{
    size_t r = (size_t)f2d_tma_rows_per_tile(nx, kind) * f2d_tma_row_pitch(nx, kind), c = (size_t)f2d_tma_cols_per_tile(ny) * engine_smem_elems(ny);
    return ((r > c ? r : c) * sizeof(float2) + 127) / 128 * 128;
}
__host__ __device__ constexpr size_t f2d_tma_fixed_bytes(int nx, int ny, int stages, int kind = F2D_C2C)
{
    return (size_t)f2d_tw_elems(nx, ny, f2d_tma_rows_variant(nx, kind)) * sizeof(float2) + (F2D_TMA_GROUPS + 2) * stages * 8 + stages * 16;
}
// As many buffers as 220 KB hold, an even number or 3 (lcm with the two groups stays small), at most 6.
__host__ __device__ constexpr int f2d_tma_stages(int nx, int ny, int kind = F2D_C2C)
{
#ifdef MUFFT_F2D_TMA_STAGES
    return MUFFT_F2D_TMA_STAGES;
#else
    int s = F2D_TMA_MAX_STAGES;
    while (s > 2 && s * f2d_tma_buf_bytes(nx, ny, kind) + f2d_tma_fixed_bytes(nx, ny, s, kind) > 220 * 1024) s--;
    return s == 5 ? 4 : s;
#endif
}
__host__ __device__ constexpr int f2d_tma_lcm(int stages) { return stages % F2D_TMA_GROUPS == 0 ? stages : stages * F2D_TMA_GROUPS; }
__host__ __device__ constexpr size_t f2d_tma_smem_bytes(int nx, int ny, int kind = F2D_C2C)
{
    return f2d_tma_stages(nx, ny, kind) * f2d_tma_buf_bytes(nx, ny, kind) + f2d_tma_fixed_bytes(nx, ny, f2d_tma_stages(nx, ny, kind), kind);
}
// the two groups + a warp for the loader and one per storer
__host__ __device__ constexpr int f2d_tma_threads(int nx, int ny, int kind = F2D_C2C) { return F2D_TMA_GROUPS * F2D_TMA_GROUP_THREADS + 32 * (1 + f2d_tma_stages(nx, ny, kind)); }

// Timing experiments with WRONG results (never in the product build): 3 = the groups only pass the buffers on,
// 4 = that, and every tile is moved as 8 rows, 5 = every tile as 8 columns.
#ifndef MUFFT_TMA_DIAG
#define MUFFT_TMA_DIAG 0
#endif
// Tickets drawn this many tiles ahead of the load, the input rows of a row ticket pulled into L2 at once (0 / 1 / 2:
// 1.069 / 1.056 / 1.057 ms).
#ifndef MUFFT_TMA_AHEAD
#define MUFFT_TMA_AHEAD 2
#endif
// Nanoseconds between polls of the loader and the storers: their polling must not take issue slots from the compute
// warps on their scheduler (0 / 20 / 100 / 400: the same within 0.3 %).
#ifndef MUFFT_TMA_POLL_NS
#define MUFFT_TMA_POLL_NS 100
#endif
// L2 policies of the four streams of tiles: 0 = none, 1 = evict_first, 2 = evict_last.
#ifndef MUFFT_TMA_POLICY_IN
#define MUFFT_TMA_POLICY_IN 1
#endif
#ifndef MUFFT_TMA_POLICY_MID_ST
#define MUFFT_TMA_POLICY_MID_ST 2
#endif
#ifndef MUFFT_TMA_POLICY_MID_LD
#define MUFFT_TMA_POLICY_MID_LD 0
#endif
#ifndef MUFFT_TMA_POLICY_OUT
#define MUFFT_TMA_POLICY_OUT 1
#endif

__device__ __forceinline__ void group_barrier(int group)
{
#ifdef MUFFT_EMU
    emu::bar_sync(group + 1, F2D_TMA_GROUP_THREADS);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(F2D_TMA_GROUP_THREADS) : "memory");
#endif
}

// (A dynamic tile order - loaders take column tiles whenever an image is complete, rows at most n images ahead, to keep the
// live intermediate under the ~24 MB of L2 that evict_last lines can hold (tools/l2_ceiling_probe.cu) - was measured:
// 1.24-1.35 ms against 1.055 ms.  All SMs then read L2 at once and DRAM at once; the static order mixes them.
// profiles/r02al_ab_tma_dynamic_order_rejected.txt.)
// (A version whose groups stored their results from registers, handing the buffer back to the loader right after the
// gather of the exchange, was measured too: 1.184 ms against 1.056 ms - the stores stall the compute warps and a finished
// row can only be announced a tile late.  profiles/r02u_ab.txt.)
// KIND = F2D_R2C: NX complex points per row = 2 NX reals in, NX + 1 bins out at a pitch of 2 NX bins (the layout of
// fft2d_kernel.cuh); the columns that carry data are NX + 1: the tensor map is that wide, and the TMA unit fills what the
// last column tile reads beyond it with zeros and drops what it would write there.
template <int NX, int NY, int DIR, int TWM, int KIND>
__global__ void __launch_bounds__(f2d_tma_threads(NX, NY, KIND), 1)
fft2d_tma_kernel(const __grid_constant__ f2d_tma_params p)
{
    constexpr int F2D_TMA_THREADS = f2d_tma_threads(NX, NY, KIND);
    static_assert(f2d_tma_kind_supported(NX, NY, KIND), "shape / kind");
    constexpr int VX = f2d_tma_rows_variant(NX, KIND);
    constexpr int PITCH = f2d_pitch(NX, KIND), WIDTH = f2d_width(NX, KIND);
    constexpr int UX = f2d_tma_units_x(NX, KIND), UY = f2d_tma_units_y(NY); // rows per warp, columns per thread of a tile
    constexpr int RX = engine_regs(NX, VX), TX = NX / RX, ROWS = f2d_tma_rows_per_tile(NX, KIND), WARPS = F2D_TMA_GROUP_THREADS / 32;
    constexpr int RY = engine_regs(NY), TY = NY / RY, COLS = f2d_tma_cols_per_tile(NY), C8 = F2D_TMA_GROUP_THREADS / TY;
    constexpr unsigned NR = f2d_tma_row_tiles(NX, NY, KIND), NC = f2d_tma_col_tiles(NX, NY, KIND);
    constexpr int ROW_PITCH = f2d_tma_row_pitch(NX, KIND);
    constexpr size_t BUF = f2d_tma_buf_bytes(NX, NY, KIND);
    constexpr int F2D_TMA_STAGES = f2d_tma_stages(NX, NY, KIND), LCM = f2d_tma_lcm(F2D_TMA_STAGES);
    static_assert(TX == 32 && TY == 32, "one warp per row, 8 columns per group");
    constexpr unsigned ROW_TILE_BYTES = ROWS * NX * sizeof(cpx), COL_TILE_BYTES = COLS * NY * sizeof(cpx);
    constexpr int UTIL0 = F2D_TMA_GROUPS * F2D_TMA_GROUP_THREADS; // first thread of the loader's warp
    static_assert(F2D_TMA_STAGES >= 2 && F2D_TMA_STAGES < 32, "buffers");
    static_assert((size_t)ROWS * ROW_PITCH * sizeof(cpx) <= BUF && (ROW_PITCH * sizeof(cpx)) % 16 == 0, "row tile layout");
    static_assert((f2d_tw_elems(NX, NY, VX) * sizeof(cpx)) % 16 == 0, "meta alignment");
    // Used as it is: pointer arithmetic through integers would turn every LDS / STS below into a generic LD / ST (seen
    // in the first ncu capture).  The kernel has no static shared memory: offset 0, aligned for the tensor copies.
    MUFFT_DYNAMIC_SMEM(smem_raw);
    unsigned char *base = smem_raw;
    cpx *tws = reinterpret_cast<cpx *>(base + F2D_TMA_STAGES * BUF);
    uint4 *meta = reinterpret_cast<uint4 *>(tws + f2d_tw_elems(NX, NY, VX)); // x = type, y = image, z = tile
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(meta + F2D_TMA_STAGES);
    // full[b * GROUPS + g]: one barrier per (buffer, group) - a parity wait only works for a waiter that sees EVERY phase
    // of its barrier, and a group meets a given buffer every second time the buffer is used
    unsigned long long *full = bars, *staged = bars + F2D_TMA_GROUPS * F2D_TMA_STAGES, *freed = staged + F2D_TMA_STAGES;

    const mcu_2d_args &a = p.a;
    const int tid = threadIdx.x;
    const cpx *twx_g = static_cast<const cpx *>(a.tw_x), *twy_g = static_cast<const cpx *>(a.tw_y);
    for (int i = tid; i < engine_tw_total(NX, VX); i += F2D_TMA_THREADS) tws[i] = __ldg(twx_g + i);
    if constexpr (NX != NY || VX != 0)
        for (int i = tid; i < engine_tw_total(NY); i += F2D_TMA_THREADS) tws[engine_tw_total(NX, VX) + i] = __ldg(twy_g + i);
    const cpx *twx = tws, *twy = (NX != NY || VX != 0) ? tws + engine_tw_total(NX, VX) : tws;
    if (tid == 0) {
        for (int b = 0; b < F2D_TMA_STAGES; b++) {
            for (int g = 0; g < F2D_TMA_GROUPS; g++) mbar_init(full + b * F2D_TMA_GROUPS + g, 1);
            mbar_init(staged + b, F2D_TMA_GROUP_THREADS);
            mbar_init(freed + b, 1);
        }
        mbar_init_fence();
    }
    __syncthreads(); // the only CTA-wide barrier

    unsigned *ticket_counter = a.counters;
    unsigned *produced = a.counters + 1; // [batch] row tiles finished per image

    if (tid < UTIL0) {
        // ---- compute groups ----------------------------------------------------------------------------------------
        const int group = tid / F2D_TMA_GROUP_THREADS, gtid = tid % F2D_TMA_GROUP_THREADS;
        int exits = 0;
        for (unsigned s = group;; s += F2D_TMA_GROUPS) {
            const unsigned b = s % F2D_TMA_STAGES;
            [[maybe_unused]] const unsigned use = s / F2D_TMA_STAGES; // (debug message)
            cpx *buf = reinterpret_cast<cpx *>(base + b * BUF);
            MUFFT_TMA_WAIT(0, full + b * F2D_TMA_GROUPS + group, (s / LCM) & 1, "full", b, use);
            const unsigned type = meta[b].x;
            if (type == F2D_TMA_EXIT) {
                mbar_arrive(staged + b);
                if (++exits == LCM / F2D_TMA_GROUPS) break;
                continue;
            }
            if (MUFFT_TMA_DIAG >= 3) {
            } else if (type == F2D_TMA_ROWS) {
                const int t = gtid % TX;
#pragma unroll 1
                for (int u = 0; u < UX; u++) { // the warp's rows, one after the other: each in its own part of the buffer
                    const int slot = gtid / TX + u * WARPS;
                    cpx *sm = buf + slot * ROW_PITCH;
                    cpx v[RX];
#pragma unroll
                    for (int e = 0; e < RX; e++) v[e] = sm[t + e * TX];
                    __syncwarp(); // the row is in registers before the exchange overwrites it
                    engine_run<NX, DIR, 1, TX, ROWS, 0, VX, -1, TWM>(v, sm, twx, t, slot);
                    __syncwarp(); // every lane has gathered
                    if constexpr (KIND == F2D_R2C) {
                        // resolve in registers (mirrored last pass), bins 0 .. NX into the row's natural positions
                        rows_store_spectrum<NX, MCU_OUT_R2C, TX, ROWS, VX>(v, [sm](int k, cpx val) { sm[k] = val; }, sm, static_cast<const cpx *>(a.rtw), t, slot, true);
                    } else {
#pragma unroll
                        for (int e = 0; e < RX; e++) sm[t + e * TX] = v[e];
                    }
                }
            } else {
                const int c = gtid % C8, t = gtid / C8;
                cpx v[UY][RY]; // the thread's columns c, c + 8: all in registers before the first exchange overwrites the tile
#pragma unroll
                for (int u = 0; u < UY; u++)
#pragma unroll
                    for (int e = 0; e < RY; e++) v[u][e] = buf[(t + e * TY) * COLS + c + u * C8];
                group_barrier(group);
#pragma unroll
                for (int u = 0; u < UY; u++) // exchange areas of different columns are disjoint
                    engine_run<NY, DIR, COLS, F2D_TMA_GROUP_THREADS, F2D_TMA_GROUPS, 0, 0, -1, TWM>(v[u], buf + c + u * C8, twy, t, group);
                if (WIDTH % COLS != 0 && meta[b].z == NC - 1) {
                    // The last column tile of an R2C image carries WIDTH % COLS columns.  The TMA unit clips a store at the
                    // edge of the tensor map in units of 16 bytes, i.e. it would write one padding column with zeros (seen:
                    // column 513): these few columns go out from registers instead, and the storer skips the tile.
#pragma unroll
                    for (int u = 0; u < UY; u++)
                        if (c + u * C8 < WIDTH % COLS) {
                            cpx *o = static_cast<cpx *>(a.out) + (size_t)meta[b].y * a.out_image + (size_t)(NC - 1) * COLS + c + u * C8;
#pragma unroll
                            for (int e = 0; e < RY; e++) __stcs(o + (size_t)(t + e * TY) * PITCH, v[u][e]);
                        }
                } else {
                    group_barrier(group);
#pragma unroll
                    for (int u = 0; u < UY; u++)
#pragma unroll
                        for (int e = 0; e < RY; e++) buf[(t + e * TY) * COLS + c + u * C8] = v[u][e];
                }
            }
            fence_proxy_async_smem(); // this thread's writes, before the TMA unit reads the buffer
            mbar_arrive(staged + b);
        }
    } else if (tid == UTIL0) {
        // ---- loader ------------------------------------------------------------------------------------------------
        const unsigned long long total = (unsigned long long)a.batch * (NR + NC);
        const unsigned lag = a.lag < a.batch ? a.lag : (unsigned)a.batch;
        const unsigned long long head = (unsigned long long)lag * NR;
        const unsigned long long body = (unsigned long long)(a.batch - lag) * (NR + NC);
        const unsigned long long pol_in = l2_policy<MUFFT_TMA_POLICY_IN>(), pol_mid = l2_policy<MUFFT_TMA_POLICY_MID_LD>();
        auto decode = [&](unsigned k, unsigned &image, unsigned &tile) -> bool { // true = rows
            if (k < head) {
                image = k / NR; tile = k % NR;
                return true;
            } else if (k - head < body) {
                const unsigned long long kk = k - head;
                const unsigned i = lag + (unsigned)(kk / (NR + NC)), j = (unsigned)(kk % (NR + NC));
                if (j < NR) { image = i; tile = j; return true; }
                image = i - lag; tile = j - NR;
                return false;
            }
            const unsigned long long kk = k - head - body;
            image = (unsigned)a.batch - lag + (unsigned)(kk / NC); tile = (unsigned)(kk % NC);
            return false;
        };
        // Tickets are drawn MUFFT_TMA_AHEAD tiles ahead; the input rows of a row ticket are pulled into L2 at once (the
        // columns read the intermediate: in L2 already).
        auto draw = [&]() -> unsigned {
            const unsigned k = atomicAdd(ticket_counter, 1u);
#if MUFFT_TMA_AHEAD > 0
            if (k < total) {
                unsigned image, tile;
                if (decode(k, image, tile))
                    prefetch_l2_bulk(static_cast<const cpx *>(a.in) + (size_t)image * a.in_image + (size_t)tile * ROWS * NX, ROW_TILE_BYTES);
            }
#endif
            return k;
        };
        unsigned t0 = draw();
#if MUFFT_TMA_AHEAD > 1
        unsigned t1 = draw();
#endif
        int exits = 0;
        for (unsigned s = 0;; s++) {
            const unsigned b = s % F2D_TMA_STAGES, use = s / F2D_TMA_STAGES;
            unsigned char *buf = base + b * BUF;
            unsigned long long *fb = full + b * F2D_TMA_GROUPS + s % F2D_TMA_GROUPS; // the group that gets this sequence number
            const unsigned k = t0;
            if (k >= total) {
                MUFFT_TMA_WAIT(MUFFT_TMA_POLL_NS, freed + b, (use & 1) ^ 1, "free(exit)", b, use);
                meta[b] = make_uint4(F2D_TMA_EXIT, 0, 0, 0);
                mbar_arrive(fb);
                if (++exits == LCM) break;
                continue;
            }
#if MUFFT_TMA_AHEAD > 1
            t0 = t1;
            t1 = draw();
#else
            t0 = draw(); // its round trip overlaps the waits below
#endif
            unsigned image, tile;
#if MUFFT_TMA_DIAG == 4
            decode(k, image, tile);
            const bool rows = true;
#elif MUFFT_TMA_DIAG == 5
            decode(k, image, tile);
            const bool rows = false;
#else
            const bool rows = decode(k, image, tile);
#endif
            if (!rows && MUFFT_TMA_DIAG != 5) {
                // the rows of this image, written by the TMA units of other SMs
                unsigned ns = 32;
                const unsigned long long w0 = global_ns();
                while (ld_relaxed(produced + image) < NR) {
                    __nanosleep(ns);
                    if (ns < 1024) ns *= 2;
                    else if (global_ns() - w0 > MUFFT_TMA_WATCHDOG_NS) {
#ifdef MUFFT_TMA_DEBUG
                        printf("cta %d loader: timeout rows of image %u: %u of %u (ticket %u)\n", (int)blockIdx.x, image, ld_relaxed(produced + image), NR, k);
                        return;
#else
                        __trap();
#endif
                    }
                }
                fence_acquire_gpu();
                fence_proxy_async_all(); // generic-proxy view of the counter, async-proxy reads of the data
            }
            MUFFT_TMA_WAIT(MUFFT_TMA_POLL_NS, freed + b, (use & 1) ^ 1, "free", b, use);
            meta[b] = make_uint4(rows ? F2D_TMA_ROWS : F2D_TMA_COLS, image, tile, 0);
            mbar_arrive_expect_tx(fb, rows ? ROW_TILE_BYTES : COL_TILE_BYTES);
            if (rows) {
                const cpx *src = static_cast<const cpx *>(a.in) + (size_t)image * a.in_image + (size_t)tile * ROWS * NX;
#pragma unroll 1
                for (int r = 0; r < ROWS; r++)
                    bulk_load_1d<MUFFT_TMA_POLICY_IN>(buf + (size_t)r * ROW_PITCH * sizeof(cpx), src + (size_t)r * NX, NX * sizeof(cpx), fb, pol_in);
            } else {
#pragma unroll 1
                for (int q = 0; q < NY / F2D_TMA_BOX_ROWS; q++)
                    tensor_load_2d<MUFFT_TMA_POLICY_MID_LD>(buf + (size_t)q * F2D_TMA_BOX_ROWS * COLS * sizeof(cpx), &p.map, (int)(tile * COLS * 2),
                                                            (int)(image * NY + q * F2D_TMA_BOX_ROWS), fb, pol_mid);
            }
        }
    } else if (tid >= UTIL0 + 32 && (tid & 31) == 0) {
        // ---- storers: one per buffer (lane 0 of a warp of its own) --------------------------------------------------------
        const unsigned b = (tid - (UTIL0 + 32)) / 32;
        unsigned char *buf = base + b * BUF;
        const unsigned long long pol_mid = l2_policy<MUFFT_TMA_POLICY_MID_ST>(), pol_out = l2_policy<MUFFT_TMA_POLICY_OUT>();
        int exits = 0;
        for (unsigned use = 0;; use++) {
            MUFFT_TMA_WAIT(MUFFT_TMA_POLL_NS, staged + b, use & 1, "staged", b, use);
            const uint4 m = meta[b];
            if (m.x == F2D_TMA_EXIT) {
                mbar_arrive(freed + b);
                if (++exits == LCM / F2D_TMA_STAGES) break;
                continue;
            }
            if (WIDTH % COLS != 0 && m.x == F2D_TMA_COLS && m.z == NC - 1) { // stored by the group itself, from registers
                mbar_arrive(freed + b);
                continue;
            }
            if (m.x == F2D_TMA_ROWS) {
                cpx *dst = static_cast<cpx *>(a.out) + (size_t)m.y * a.out_image + (size_t)m.z * ROWS * PITCH;
#pragma unroll 1
                for (int r = 0; r < ROWS; r++)
                    bulk_store_1d<MUFFT_TMA_POLICY_MID_ST>(dst + (size_t)r * PITCH, buf + (size_t)r * ROW_PITCH * sizeof(cpx), NX * sizeof(cpx), pol_mid);
                if constexpr (KIND == F2D_R2C) {
                    // bin NX of every row: 8 bytes, not a bulk copy
#pragma unroll 1
                    for (int r = 0; r < ROWS; r++) dst[(size_t)r * PITCH + NX] = reinterpret_cast<const cpx *>(buf)[(size_t)r * ROW_PITCH + NX];
                }
            } else {
#pragma unroll 1
                for (int q = 0; q < NY / F2D_TMA_BOX_ROWS; q++)
                    tensor_store_2d<MUFFT_TMA_POLICY_OUT>(&p.map, (int)(m.z * COLS * 2), (int)(m.y * NY + q * F2D_TMA_BOX_ROWS),
                                                          buf + (size_t)q * F2D_TMA_BOX_ROWS * COLS * sizeof(cpx), pol_out);
            }
            bulk_commit();
            bulk_wait_read(); // the TMA unit has read the buffer: the loader may refill it
            mbar_arrive(freed + b);
            if (m.x == F2D_TMA_ROWS) {
                bulk_wait_all(); // the writes are complete
                fence_proxy_async_all();
                fence_release_gpu();
                red_relaxed_add(produced + m.y, 1u);
            }
        }
        bulk_wait_all();
    }
}

} // namespace mufft
