// Batched row transforms: one kernel does a whole N-point transform (N <= 8192) per group of
// N/R threads, read once from global memory and written once.
//
// Fused into the first pass (prologue), instead of the reference's separate sweeps:
//   * zero-padded upper half            (ref kernel.c:104-122,200-226,333-387 *_half_* kernels)
//   * C2R resolve                       (ref kernel.c:36-45, called at fft.c:828-831)
//   * spectral multiply with 1/N        (ref kernel.c:22-34, called at fft.c:806-808)
// Fused into the last pass (epilogue):
//   * R2C resolve, half or full spectrum (ref kernel.c:47-82, called at fft.c:846-849)
#pragma once

#include "abi.h"
#include "fft_engine.cuh"

namespace mufft {

// Threads of one CTA and transforms per CTA for row transforms of length n.
__host__ __device__ constexpr int rows_threads_per_transform(int n) { return n / engine_regs(n); }
__host__ __device__ constexpr int rows_per_cta(int n)
{
    return rows_threads_per_transform(n) >= 256 ? 1 : 256 / rows_threads_per_transform(n);
}
__host__ __device__ constexpr int rows_cta_threads(int n) { return rows_threads_per_transform(n) * rows_per_cta(n); }
// Exchange area per transform (also used by the R2C epilogue, so never zero).
__host__ __device__ constexpr int rows_smem_elems(int n) { return n + (n >> engine_pad_shift(n)) + 1; }
__host__ __device__ constexpr size_t rows_smem_bytes(int n) { return (size_t)rows_per_cta(n) * rows_smem_elems(n) * sizeof(float2); }

template <int N, int DIR>
__global__ void __launch_bounds__(rows_cta_threads(N)) fft_rows_kernel(const mcu_rows_args a)
{
    constexpr int R = engine_regs(N);
    constexpr int T = N / R;
    constexpr int TPC = rows_per_cta(N);
    constexpr int SH = engine_pad_shift(N);
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int slot = tid / T;
    const int t = tid - slot * T;
    size_t b = (size_t)blockIdx.x * TPC + slot;
    const bool active = b < a.batch;
    if (!active) b = a.batch - 1; // keep the thread in step with the barriers; it stores nothing

    cpx *sm = reinterpret_cast<cpx *>(smem_raw) + slot * rows_smem_elems(N);
    const cpx *__restrict__ tw = static_cast<const cpx *>(a.tw);
    const cpx *__restrict__ rtw = static_cast<const cpx *>(a.rtw);
    const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + b * a.in_stride;

    cpx v[R];

    // ---- prologue ---------------------------------------------------------------------------
    if (a.in_mode == MCU_IN_COMPLEX) {
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = in[t + e * T];
    } else if (a.in_mode == MCU_IN_HALF) {
        // Slots e >= R/2 are exactly the elements >= N/2: zero, and their memory is not touched.
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = e < R / 2 ? in[t + e * T] : make_float2(0.0f, 0.0f);
    } else if (a.in_mode == MCU_IN_MUL) {
        const cpx *__restrict__ in2 = static_cast<const cpx *>(a.in2) + b * a.in2_stride;
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = cscale(a.norm, cmul(in[t + e * T], in2[t + e * T]));
    } else {
        // C2R resolve, optionally on the product of two spectra.  Reads bins k and N-k.
        const bool mul = a.in_mode == MCU_IN_MUL_C2R;
        const cpx *__restrict__ in2 = mul ? static_cast<const cpx *>(a.in2) + b * a.in2_stride : in;
#pragma unroll
        for (int e = 0; e < R; e++) {
            const int k = t + e * T;
            cpx xa = in[k];
            cpx xb = in[N - k];
            if (mul) {
                xa = cscale(a.norm, cmul(xa, in2[k]));
                xb = cscale(a.norm, cmul(xb, in2[N - k]));
            }
            xb = cconj(xb);
            v[e] = cadd(cadd(xa, xb), cmul(csub(xa, xb), __ldg(rtw + k)));
        }
    }

    // ---- the transform ----------------------------------------------------------------------------
    engine_run<N, DIR, 1>(v, sm, tw, t);

    // ---- epilogue ---------------------------------------------------------------------------
    cpx *__restrict__ out = static_cast<cpx *>(a.out) + b * a.out_stride;
    if (a.out_mode == MCU_OUT_COMPLEX) {
        if (active) {
#pragma unroll
            for (int e = 0; e < R; e++) out[t + e * T] = v[e];
        }
    } else {
        // X[k] = 1/2 ((Z[k] + conj Z[N-k]) + rtw[k] (Z[k] - conj Z[N-k])); Z[N-k] lives in another
        // thread, so the half-length spectrum makes one more trip through shared memory.
        const bool full = a.out_mode == MCU_OUT_R2C_FULL;
        if constexpr (engine_num_passes(N) > 1) __syncthreads();
        engine_scatter_natural<N, 1>(v, sm, t);
        __syncthreads();
        if (active) {
#pragma unroll
            for (int e = 0; e < R; e++) {
                const int k = t + e * T;
                if (k == 0) {
                    out[0] = make_float2(v[e].x + v[e].y, 0.0f);
                    out[N] = make_float2(v[e].x - v[e].y, 0.0f);
                } else {
                    const cpx za = v[e];
                    const cpx zb = cconj(sm[pad_pos<SH>(N - k)]);
                    const cpx fe = cadd(za, zb);
                    const cpx fo = cmul(__ldg(rtw + k), csub(za, zb));
                    out[k] = cscale(0.5f, cadd(fe, fo));
                    if (full) out[k + N] = cscale(0.5f, csub(fe, fo));
                }
            }
        }
    }
}

} // namespace mufft
