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

namespace mufft {

#define MUFFT_F2D_THREADS 256

__host__ __device__ constexpr int f2d_rows_per_tile(int nx) { return MUFFT_F2D_THREADS / engine_threads(nx); }
__host__ __device__ constexpr int f2d_cols_per_tile(int ny) { return MUFFT_F2D_THREADS / engine_threads(ny); }
__host__ __device__ constexpr bool f2d_supported(int nx, int ny)
{
    return engine_regs(nx) > 0 && engine_regs(ny) > 0 && engine_threads(nx) <= MUFFT_F2D_THREADS &&
           engine_threads(ny) * 4 <= MUFFT_F2D_THREADS && (engine_threads(nx) <= 32 || f2d_rows_per_tile(nx) <= 15) &&
           ny % f2d_rows_per_tile(nx) == 0 && nx % f2d_cols_per_tile(ny) == 0;
}
__host__ __device__ constexpr size_t f2d_smem_bytes(int nx, int ny)
{
    size_t r = (size_t)f2d_rows_per_tile(nx) * engine_smem_elems(nx);
    size_t c = (size_t)f2d_cols_per_tile(ny) * engine_smem_elems(ny);
    return (r > c ? r : c) * sizeof(float2) + 16;
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int NX, int NY, int DIR>
__global__ void __launch_bounds__(MUFFT_F2D_THREADS, 2) fft2d_fused_kernel(const mcu_2d_args a)
{
    constexpr int RX = engine_regs(NX), TX = NX / RX, ROWS = f2d_rows_per_tile(NX);
    constexpr int RY = engine_regs(NY), TY = NY / RY, COLS = f2d_cols_per_tile(NY);
    constexpr unsigned NR = NY / ROWS; // row tickets per image
    constexpr unsigned NC = NX / COLS; // column tickets per image
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned s_ticket;
    cpx *smem = reinterpret_cast<cpx *>(smem_raw);

    const int tid = threadIdx.x;
    const cpx *__restrict__ twx = static_cast<const cpx *>(a.tw_x);
    const cpx *__restrict__ twy = static_cast<const cpx *>(a.tw_y);
    unsigned *ticket_counter = a.counters;
    unsigned *rows_done = a.counters + 1;
    const unsigned long long total = (unsigned long long)a.batch * (NR + NC);

    for (;;) {
        __syncthreads(); // the previous ticket's readers of s_ticket and of the exchange area are done
        if (tid == 0) s_ticket = atomicAdd(ticket_counter, 1u);
        __syncthreads();
        const unsigned k = s_ticket;
        if (k >= total) break;

        // ticket -> (image, phase, tile)
        unsigned image, tile;
        bool is_rows;
        if (k < NR) {
            image = 0; tile = k; is_rows = true;
        } else {
            const unsigned kk = k - NR;
            const unsigned s = 1 + kk / (NR + NC), j = kk % (NR + NC);
            if (s < a.batch && j < NR) { image = s; tile = j; is_rows = true; }
            else if (s < a.batch) { image = s - 1; tile = j - NR; is_rows = false; }
            else { image = (unsigned)a.batch - 1; tile = j; is_rows = false; }
        }

        cpx *__restrict__ out = static_cast<cpx *>(a.out) + (size_t)image * a.out_image;
        if (is_rows) {
            const int slot = tid / TX, t = tid - slot * TX;
            const size_t row = (size_t)tile * ROWS + slot;
            const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + (size_t)image * a.in_image + row * NX;
            cpx v[RX];
#pragma unroll
            for (int e = 0; e < RX; e++) v[e] = __ldcs(in + t + e * TX); // streamed: read once
            engine_run<NX, DIR, 1, TX, ROWS>(v, smem + slot * engine_smem_elems(NX), twx, t, slot);
            cpx *__restrict__ o = out + row * NX;
#pragma unroll
            for (int e = 0; e < RX; e++) o[t + e * TX] = v[e];
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicAdd(rows_done + image, 1u);
        } else {
            if (tid == 0) {
                unsigned ns = 32;
                while (ld_acquire(rows_done + image) < NR) {
                    __nanosleep(ns);
                    if (ns < 1024) ns *= 2;
                }
            }
            __syncthreads();
            const int c = tid % COLS, t = tid / COLS;
            cpx *__restrict__ col = out + (size_t)tile * COLS + c;
            cpx v[RY];
#pragma unroll
            for (int e = 0; e < RY; e++) v[e] = __ldcg(col + (size_t)(t + e * TY) * NX); // written by other SMs: L2 only
            engine_run<NY, DIR, COLS, MUFFT_F2D_THREADS, 1>(v, smem + c, twy, t, 0);
#pragma unroll
            for (int e = 0; e < RY; e++) __stcs(col + (size_t)(t + e * TY) * NX, v[e]);
        }
    }
}

} // namespace mufft
