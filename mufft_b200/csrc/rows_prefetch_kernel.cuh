// Batched row transforms of 2048, 8192 and 16384 points - C2C, the convolution stages (zero-padded load, spectral
// multiply), R2C and C2R: the row kernel's shapes (one CTA per transform) made PERSISTENT, with the next transform fetched
// by the TMA unit while this one finishes.
//
// fft_rows_kernel<16384, ...> is alone on its SM: it loads 128 KB, computes three passes, stores 128 KB, and nothing
// overlaps - 0.62 of the HBM roofline.  Here the exchange area doubles as the landing buffer of the next transform: it is
// idle from the gather of the last exchange to the scatter of the next transform's first one, which is when the last
// pass and the stores of this transform run.  Fraction of the HBM roofline, row kernel -> this kernel (profiles/README.md
// R2.12; r02ac / r02ad / r02af / r02ai / r02aq _ab.txt):
//   C2C      16384 points 0.62 -> 0.77     8192 points 0.83 -> 0.92     2048 points (CTAs of 64 threads) 0.85 -> 0.91
//   stereo convolution at 8192 points, both stages   0.71 -> 0.75
//   R2C      32768 reals 0.56 -> 0.65      16384 reals 0.74 -> 0.76     (mirrored last pass: the resolve stays in registers)
//   C2R      32768 reals 0.53 -> 0.62      16384 reals 0.69 -> 0.75     4096 reals 0.87 -> 0.90
// Not used where it loses: C2C at 4096 points (0.98 -> 0.94: three CTAs of 80 registers per SM already hide their loads),
// R2C of 8192 / 4096 reals (0.81 -> 0.80 / 0.93 -> 0.86).  A warp-specialised version for 8192 points (loader, two compute
// groups, storers, three buffers, as fft2d_tma_kernel.cuh) reached 0.88 and lost on the convolution stages: gone.
//
//   wait for the bulk copy of transform k     (mbarrier)
//   registers <- buffer, natural order; barrier
//   pass 0, exchange, pass 1, exchange        (the engine, in the buffer)
//   barrier: every thread has gathered -> thread 0 asks the TMA unit for transform k + gridDim.x into the buffer
//   pass 2, stores from registers (streaming)
//
// Static assignment, no extra warps, no storer: the stores are the row kernel's own.
#pragma once

#include "rows_kernel.cuh"
#include "tma_util.cuh"

namespace mufft {

#ifndef MUFFT_ROWS_PREFETCH_R2C_ALSO
#define MUFFT_ROWS_PREFETCH_R2C_ALSO 0 // (A/B builds: one more length for the R2C flavour)
#endif
// Complex output with the general engine shape, or an R2C output (half or full spectrum) with the mirrored-last-pass shape,
// whose resolve happens in registers on the way out: nothing after the last gather touches the buffer.
__host__ __device__ constexpr bool rows_prefetch_supported(int n, int in_mode, int out_mode)
{
    const int v = rows_variant(n, in_mode, out_mode);
    if (engine_regs(n, v) == 0 || rows_per_cta(n, v) != 1 || engine_num_passes(n, v) < 2) return false;
    if (out_mode == MCU_OUT_COMPLEX && in_mode == MCU_IN_C2R) // spectrum rows of n + 1 bins, any pitch
        return (n == 2048 || n == 8192 || n == 16384) && v == MUFFT_VARIANT_MIRROR_FIRST;
    if (out_mode == MCU_OUT_COMPLEX)
        return (n == 2048 || n == 8192 || n == 16384) && v == 0 && (in_mode == MCU_IN_COMPLEX || in_mode == MCU_IN_HALF || in_mode == MCU_IN_MUL);
    if (out_mode == MCU_OUT_R2C || out_mode == MCU_OUT_R2C_FULL)
        return (n == MUFFT_ROWS_PREFETCH_R2C_ALSO || n == 8192 || n == 16384) && v == MUFFT_VARIANT_MIRROR_LAST && (in_mode == MCU_IN_COMPLEX || in_mode == MCU_IN_HALF);
    return false;
}
__host__ __device__ constexpr size_t rows_prefetch_smem_bytes(int n, int v = 0) { return ((size_t)engine_smem_elems(n, v) * sizeof(float2) + 127) / 128 * 128 + 16; }

template <int N, int DIR, int IN_MODE, int OUT_MODE>
__global__ void __launch_bounds__(rows_cta_threads(N, rows_variant(N, IN_MODE, OUT_MODE)), engine_min_blocks(N, rows_variant(N, IN_MODE, OUT_MODE)))
fft_rows_prefetch_kernel(const mcu_rows_args a)
{
    static_assert(rows_prefetch_supported(N, IN_MODE, OUT_MODE), "shape / mode");
    constexpr int V = rows_variant(N, IN_MODE, OUT_MODE);
    constexpr int R = engine_regs(N, V), T = N / R, NP = engine_num_passes(N, V);
    constexpr int PK = rows_packed(N, IN_MODE, OUT_MODE);
    constexpr bool C2R = IN_MODE == MCU_IN_C2R;
    constexpr unsigned BYTES = (IN_MODE == MCU_IN_HALF ? N / 2 : N) * sizeof(cpx); // a zero-padded input is half as long
    MUFFT_DYNAMIC_SMEM(smem_raw);
    cpx *buf = reinterpret_cast<cpx *>(smem_raw);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + rows_prefetch_smem_bytes(N, V) - 16);
    const int t = threadIdx.x;
    const cpx *__restrict__ tw = static_cast<const cpx *>(a.tw);
    const unsigned long long pol = l2_policy<1>(); // read once
    // C2R: a spectrum row has N + 1 bins = 8 bytes more than a multiple of 16, so with a tight pitch every other row starts
    // 8 bytes off.  The bulk copy starts at the 16-byte boundary at or below the row (one bin of the previous row comes
    // along: `off` = 1) and ends at the last boundary inside it; bin N, which that may leave out, is fetched by thread 0
    // with an ordinary load when it asks for the copy and put in place when the copy has landed.
    cpx nyquist = make_float2(0.0f, 0.0f); // thread 0: bin N of the transform in flight
    auto row_off = [&](size_t k) -> unsigned {
        return C2R ? (unsigned)((reinterpret_cast<uintptr_t>(static_cast<const cpx *>(a.in) + k * a.in_stride) & 15) / sizeof(cpx)) : 0u;
    };
    auto ask = [&](size_t k) { // thread 0
        const cpx *row = static_cast<const cpx *>(a.in) + k * a.in_stride;
        if constexpr (C2R) {
            const unsigned off = row_off(k), bytes = ((off + N + 1) * (unsigned)sizeof(cpx)) & ~15u;
            mbar_arrive_expect_tx(full, bytes);
            bulk_load_1d<1>(buf, row - off, bytes, full, pol);
            nyquist = __ldg(row + N);
        } else {
            mbar_arrive_expect_tx(full, BYTES);
            bulk_load_1d<1>(buf, row, BYTES, full, pol);
        }
    };

    if (t == 0) {
        mbar_init(full, 1);
        mbar_init_fence();
        if (blockIdx.x < a.batch) ask(blockIdx.x);
    }
    __syncthreads();

    unsigned phase = 0;
    for (size_t k = blockIdx.x; k < a.batch; k += gridDim.x, phase ^= 1) {
        MUFFT_TMA_WAIT(0, full, phase, "full", 0, phase);
        const unsigned off = row_off(k);
        if constexpr (C2R) {
            if (t == 0) buf[off + N] = nyquist;
            __syncthreads();
        }
        cpx v[R];
        rows_load<N, IN_MODE, V>(v, static_cast<const cpx *>(buf + off), static_cast<const cpx *>(a.in2) + k * a.in2_stride, static_cast<const cpx *>(a.rtw), a.norm, t);
        __syncthreads(); // the transform is in registers before the exchange overwrites it
        static_for<NP - 1>([&](auto pc) {
            constexpr int PASS = decltype(pc)::value;
            engine_butterflies<N, DIR, PASS, V, PK>(v, tw, t);
            if constexpr (PASS > 0) __syncthreads(); // previous gather complete
            engine_scatter<N, PASS, 1, V>(v, buf, t);
            __syncthreads();
            engine_gather<N, 1, PASS + 1, V>(v, buf, t);
        });
        __syncthreads(); // every thread has gathered: the buffer is free
        if (t == 0 && k + gridDim.x < a.batch) {
            fence_proxy_async_smem();
            ask(k + gridDim.x);
        }
        engine_butterflies<N, DIR, NP - 1, V, PK>(v, tw, t);
        cpx *__restrict__ out = static_cast<cpx *>(a.out) + k * a.out_stride;
        if constexpr (OUT_MODE == MCU_OUT_COMPLEX) {
#pragma unroll
            for (int e = 0; e < R; e++) __stcs(out + t + e * T, v[e]);
        } else {
            rows_store<N, OUT_MODE, T, 1, V>(v, out, buf, static_cast<const cpx *>(a.rtw), t, 0, true); // mirrored last pass: registers only
        }
    }
}

} // namespace mufft
