// Column transforms without a transpose: a CTA owns a tile of TC adjacent columns, reads them
// as TC*8-byte row segments, runs the whole column transform on chip (same engine as the row
// kernel, TC transforms side by side in shared memory) and writes the tile back.
//
// Replaces the reference's vertical pass kernels (kernel.c:455-738, x86/kernel.h:670-971), which
// sweep the whole image once per radix pass (fft.c:975-987): here the image is read once and
// written once for the entire column transform.
#pragma once

#include "abi.h"
#include "fft_engine.cuh"

namespace mufft {

// Columns per tile: as many as keep the CTA at <= 512 threads and its exchange area within
// shared memory, at most 16 (= 128-byte row segments).
__host__ __device__ constexpr int cols_tile(int n)
{
    int t = n / engine_regs(n);
    int tc = 512 / t;
    tc = tc > 16 ? 16 : (tc < 1 ? 1 : tc);
    while (tc > 1 && (size_t)tc * engine_smem_elems(n) * sizeof(float2) > 160 * 1024) tc /= 2;
    return tc;
}
__host__ __device__ constexpr int cols_cta_threads(int n) { return cols_tile(n) * (n / engine_regs(n)); }
__host__ __device__ constexpr size_t cols_smem_bytes(int n)
{
    return (size_t)cols_tile(n) * engine_smem_elems(n) * sizeof(float2);
}

// TWIDDLE: multiply element (image, row) by roots[(image * row) mod n_total] before the
// transform: the inter-stage twiddle when a long column transform is split in two (the "image"
// index is then the first-stage output index k1, see mufft_host.c: run_cols).
template <int N, int DIR, bool TWIDDLE>
__global__ void __launch_bounds__(cols_cta_threads(N)) fft_cols_kernel(const mcu_cols_args a)
{
    constexpr int R = engine_regs(N);
    constexpr int T = N / R;
    constexpr int TC = cols_tile(N);
    MUFFT_DYNAMIC_SMEM(smem_raw);

    const int tid = threadIdx.x;
    const int c = tid % TC;  // column inside the tile: fastest, so a warp reads whole row segments
    const int t = tid / TC;  // engine thread index inside the column transform
    const size_t col = (size_t)blockIdx.x * TC + c;
    const bool active = col < a.width;
    const cpx *__restrict__ tw = static_cast<const cpx *>(a.tw);
    const unsigned rows_valid = a.rows_valid ? a.rows_valid : (unsigned)N;
    cpx *sm = reinterpret_cast<cpx *>(smem_raw) + c;

    for (size_t image = blockIdx.y; image < a.batch; image += gridDim.y) {
        const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + image * a.in_image + col;
        cpx *__restrict__ out = static_cast<cpx *>(a.out) + image * a.out_image + col;
        cpx v[R];
#pragma unroll
        for (int e = 0; e < R; e++)
            v[e] = (active && (unsigned)(t + e * T) < rows_valid) ? in[(size_t)(t + e * T) * a.in_pitch] : make_float2(0.0f, 0.0f);
        if constexpr (TWIDDLE) {
            const cpx *__restrict__ roots = static_cast<const cpx *>(a.roots);
#pragma unroll
            for (int e = 0; e < R; e++)
                v[e] = cmul(v[e], __ldg(roots + (((unsigned)image * (unsigned)(t + e * T)) & a.root_mask)));
        }

        engine_run<N, DIR, TC, cols_cta_threads(N), 1>(v, sm, tw, t, 0);

        if constexpr (TC == 16) {
            // dead scratch (the intermediate of a two-sweep column transform): one lane per 128-byte row segment drops it
            // from L2, so that it is never written back to HBM.  Every lane's loads completed before the engine's exchanges.
            if (a.discard_input && c == 0) {
#ifndef MUFFT_EMU
#pragma unroll
                for (int e = 0; e < R; e++)
                    if ((unsigned)(t + e * T) < rows_valid) asm volatile("discard.global.L2 [%0], 128;" ::"l"(in + (size_t)(t + e * T) * a.in_pitch) : "memory");
#endif
            }
        }

        if (active) {
#pragma unroll
            for (int e = 0; e < R; e++) out[(size_t)(t + e * T) * a.out_pitch] = v[e];
        }
        if constexpr (engine_num_passes(N) > 1) __syncthreads(); // next image reuses the exchange area
    }
}

} // namespace mufft
