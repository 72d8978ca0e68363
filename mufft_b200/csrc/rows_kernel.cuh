// Batched row transforms: one kernel does a whole N-point transform (N <= 8192) per group of
// N/R threads, read once from global memory and written once.
//
// Fused into the first pass (prologue), instead of the reference's separate sweeps:
//   * zero-padded upper half            (ref kernel.c:104-122,200-226,333-387 *_half_* kernels)
//   * C2R resolve                       (ref kernel.c:36-45, called at fft.c:828-831)
//   * spectral multiply with 1/N        (ref kernel.c:22-34, called at fft.c:806-808)
// Fused into the last pass (epilogue):
//   * R2C resolve, half or full spectrum (ref kernel.c:47-82, called at fft.c:846-849)
//
// Two kernels share the prologue / engine / epilogue code:
//   fft_rows_kernel      one group of transforms per CTA, inputs loaded straight into registers;
//   fft_rows_tma_kernel  persistent CTAs: the inputs of the NEXT groups are already on their way
//                        into shared-memory landing buffers (cp.async.bulk + mbarrier, issued by
//                        one thread) while the current group is being transformed, so enough
//                        bytes are in flight per SM even when one transform fills the CTA.
#pragma once

#include "abi.h"
#include "fft_engine.cuh"

namespace mufft {

// Threads of one CTA and transforms per CTA for row transforms of length n.
__host__ __device__ constexpr int rows_threads_per_transform(int n) { return n / engine_regs(n); }
__host__ __device__ constexpr int rows_per_cta(int n)
{
    return rows_threads_per_transform(n) >= engine_cta_threads(n) ? 1 : engine_cta_threads(n) / rows_threads_per_transform(n);
}
__host__ __device__ constexpr int rows_cta_threads(int n) { return rows_threads_per_transform(n) * rows_per_cta(n); }
// Exchange area per transform (also used by the R2C epilogue, so never zero).
__host__ __device__ constexpr int rows_smem_elems(int n) { return engine_smem_elems(n); }
__host__ __device__ constexpr size_t rows_smem_bytes(int n) { return (size_t)rows_per_cta(n) * rows_smem_elems(n) * sizeof(float2); }

// ---- prologue: registers <- input, element k = t + e*T in slot e ---------------------------------
// `in` may point to global memory or to a shared-memory landing buffer.
template <int N, int IN_MODE, typename Ptr>
__device__ __forceinline__ void rows_load(cpx (&v)[engine_regs(N)], Ptr in, const cpx *__restrict__ in2,
                                          const cpx *__restrict__ rtw, float norm, int t)
{
    constexpr int R = engine_regs(N);
    constexpr int T = N / R;
    if constexpr (IN_MODE == MCU_IN_COMPLEX) {
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = in[t + e * T];
    } else if constexpr (IN_MODE == MCU_IN_HALF) {
        // Slots e >= R/2 are exactly the elements >= N/2: zero, and their memory is not touched.
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = e < R / 2 ? in[t + e * T] : make_float2(0.0f, 0.0f);
    } else if constexpr (IN_MODE == MCU_IN_MUL) {
        cpx w[R];
#pragma unroll
        for (int e = 0; e < R; e++) w[e] = __ldg(in2 + t + e * T);
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = in[t + e * T];
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = cscale(norm, cmul(v[e], w[e]));
    } else {
        // C2R resolve, optionally on the product of two spectra.  Reads bins k and N-k.
        constexpr bool mul = IN_MODE == MCU_IN_MUL_C2R;
#pragma unroll
        for (int e = 0; e < R; e++) {
            const int k = t + e * T;
            cpx xa = in[k];
            cpx xb = in[N - k];
            if constexpr (mul) {
                xa = cscale(norm, cmul(xa, __ldg(in2 + k)));
                xb = cscale(norm, cmul(xb, __ldg(in2 + N - k)));
            }
            xb = cconj(xb);
            v[e] = cadd(cadd(xa, xb), cmul(csub(xa, xb), __ldg(rtw + k)));
        }
    }
}

// ---- epilogue: output <- registers -----------------------------------------------------------------
template <int N, int OUT_MODE, int T_SYNC, int SLOTS>
__device__ __forceinline__ void rows_store(cpx (&v)[engine_regs(N)], cpx *__restrict__ out, cpx *sm,
                                           const cpx *__restrict__ rtw, int t, int slot, bool active)
{
    constexpr int R = engine_regs(N);
    constexpr int T = N / R;
    constexpr int SH = engine_pad_shift(N);
    if constexpr (OUT_MODE == MCU_OUT_COMPLEX) {
        if (active) {
#pragma unroll
            for (int e = 0; e < R; e++) out[t + e * T] = v[e];
        }
    } else {
        // X[k] = 1/2 ((Z[k] + conj Z[N-k]) + rtw[k] (Z[k] - conj Z[N-k])); Z[N-k] lives in another
        // thread, so the half-length spectrum makes one more trip through shared memory.
        constexpr bool full = OUT_MODE == MCU_OUT_R2C_FULL;
        if constexpr (engine_num_passes(N) > 1) engine_sync<T_SYNC, SLOTS>(slot);
        engine_scatter_natural<N, 1>(v, sm, t);
        engine_sync<T_SYNC, SLOTS>(slot);
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
                    if constexpr (full) out[k + N] = cscale(0.5f, csub(fe, fo));
                }
            }
        }
    }
}

// ---- kernel 1: one group per CTA, direct loads ---------------------------------------------------
template <int N, int DIR, int IN_MODE, int OUT_MODE>
__global__ void __launch_bounds__(rows_cta_threads(N), engine_min_blocks(N)) fft_rows_kernel(const mcu_rows_args a)
{
    constexpr int R = engine_regs(N);
    constexpr int T = N / R;
    constexpr int TPC = rows_per_cta(N);
    extern __shared__ __align__(128) unsigned char smem_raw[];

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
    const cpx *__restrict__ in2 = static_cast<const cpx *>(a.in2) + b * a.in2_stride;

    cpx v[R];
    rows_load<N, IN_MODE>(v, in, in2, rtw, a.norm, t);
    engine_run<N, DIR, 1, T, TPC>(v, sm, tw, t, slot);
    cpx *__restrict__ out = static_cast<cpx *>(a.out) + b * a.out_stride;
    rows_store<N, OUT_MODE, T, TPC>(v, out, sm, rtw, t, slot, active);
}

// ---- kernel 2: persistent, inputs prefetched into shared memory by bulk async copies ------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (16-byte aligned addresses, size a multiple of 16), completion
// counted in bytes on `bar`.
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

#define MUFFT_TMA_STAGES 2
// Bytes a landing buffer receives per transform (a multiple of 16).
__host__ __device__ constexpr int rows_land_bytes(int n, int in_mode)
{
    return in_mode == MCU_IN_HALF ? n * 4 : (in_mode == MCU_IN_C2R || in_mode == MCU_IN_MUL_C2R) ? (n + 2) * 8 : n * 8;
}
__host__ __device__ constexpr size_t rows_tma_smem_bytes(int n, int in_mode)
{
    return 128 + (size_t)rows_per_cta(n) * (MUFFT_TMA_STAGES * (size_t)rows_land_bytes(n, in_mode) + rows_smem_elems(n) * sizeof(float2));
}
// The bulk-copy path needs whole 16-byte chunks.
__host__ __device__ constexpr bool rows_tma_supported(int n, int in_mode)
{
    return n >= 64 && rows_land_bytes(n, in_mode) % 16 == 0 && rows_tma_smem_bytes(n, in_mode) <= 227 * 1024 &&
           MUFFT_TMA_STAGES * rows_per_cta(n) * 8 <= 128;
}

template <int N, int DIR, int IN_MODE, int OUT_MODE>
__global__ void __launch_bounds__(rows_cta_threads(N), engine_min_blocks(N)) fft_rows_tma_kernel(const mcu_rows_args a)
{
    constexpr int R = engine_regs(N);
    constexpr int T = N / R;
    constexpr int TPC = rows_per_cta(N);
    constexpr int STAGES = MUFFT_TMA_STAGES;
    constexpr int LAND = rows_land_bytes(N, IN_MODE);
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int slot = tid / T;
    const int t = tid - slot * T;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw) + slot * STAGES;
    unsigned char *land = smem_raw + 128 + (size_t)slot * STAGES * LAND;
    cpx *sm = reinterpret_cast<cpx *>(smem_raw + 128 + (size_t)TPC * STAGES * LAND) + slot * rows_smem_elems(N);
    const cpx *__restrict__ tw = static_cast<const cpx *>(a.tw);
    const cpx *__restrict__ rtw = static_cast<const cpx *>(a.rtw);
    const size_t groups = (a.batch + TPC - 1) / TPC;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    engine_sync<T, TPC>(slot);

    auto issue = [&](int stage, size_t g) {
        const size_t b = g * TPC + slot;
        if (g < groups && b < a.batch) {
            mbar_expect_tx(bars + stage, LAND);
            bulk_load(land + stage * LAND, static_cast<const cpx *>(a.in) + b * a.in_stride, LAND, bars + stage);
        }
    };
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) issue(s, (size_t)blockIdx.x + (size_t)s * gridDim.x);
    }

    int stage = 0;
    uint32_t parity = 0;
    for (size_t g = blockIdx.x; g < groups; g += gridDim.x) {
        size_t b = g * TPC + slot;
        const bool active = b < a.batch;
        if (!active) b = a.batch - 1;
        const cpx *__restrict__ in2 = static_cast<const cpx *>(a.in2) + b * a.in2_stride;

        cpx v[R];
        if (active) mbar_wait(bars + stage, parity);
        rows_load<N, IN_MODE>(v, reinterpret_cast<const cpx *>(land + stage * LAND), in2, rtw, a.norm, t);
        // Order this thread's generic-proxy reads of the landing buffer before the async-proxy
        // (bulk copy) write that refills it: without the fence a queued LDS can still be reading
        // when the next copy lands (seen as rare wrong transforms in long persistent loops).
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        // every thread of the transform has its inputs in registers: the landing buffer can be
        // refilled, and the exchange area is free of the previous group's readers
        engine_sync<T, TPC>(slot);
        if (t == 0) issue(stage, g + (size_t)STAGES * gridDim.x);

        engine_run<N, DIR, 1, T, TPC>(v, sm, tw, t, slot);
        cpx *__restrict__ out = static_cast<cpx *>(a.out) + b * a.out_stride;
        rows_store<N, OUT_MODE, T, TPC>(v, out, sm, rtw, t, slot, active);

        if (++stage == STAGES) {
            stage = 0;
            parity ^= 1;
        }
    }
}

} // namespace mufft
