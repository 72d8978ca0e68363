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
// fft_rows_kernel: one group of transforms per CTA, inputs loaded straight into registers.  (A persistent
// variant fed by bulk async copies into shared-memory landing buffers was measured in round 1 and was
// 2-12 % slower at every length - the landing-buffer read is one more shared-memory access per element on
// the busiest pipe - so it is gone; profiles/README.md keeps the numbers.)
#pragma once

#include "abi.h"
#include "fft_engine.cuh"

namespace mufft {

// Engine shape variant used by a (length, input mode, output mode) combination: the real
// transforms take the mirrored-pass variants where the shape table has them (fft_engine.cuh).
__host__ __device__ constexpr int rows_variant(int n, int in_mode, int out_mode)
{
    if ((out_mode == MCU_OUT_R2C || out_mode == MCU_OUT_R2C_FULL || out_mode == MCU_OUT_R2C_PEERS) && engine_variant_exists(n, MUFFT_VARIANT_MIRROR_LAST))
        return MUFFT_VARIANT_MIRROR_LAST;
    if ((in_mode == MCU_IN_C2R || in_mode == MCU_IN_MUL_C2R || in_mode == MCU_IN_MULACC_C2R) && engine_variant_exists(n, MUFFT_VARIANT_MIRROR_FIRST))
        return MUFFT_VARIANT_MIRROR_FIRST;
    return MUFFT_VARIANT_GENERAL;
}

// Packed-FP32 choice per kernel (-1 = the shape's flag).  Measured on B200 (profiles/README.md, A/B
// "c8k_pk"): at N = 8192 the convolution stages (zero-padded load, spectral multiply) gain 2-7 %
// from packed complex adds while the plain C2C transform of the same shape loses 3 %.
__host__ __device__ constexpr int rows_packed(int n, int in_mode, int out_mode)
{
#ifdef MUFFT_PACKED
    return -1;
#else
    return (n == 8192 && out_mode == MCU_OUT_COMPLEX && (in_mode == MCU_IN_HALF || in_mode == MCU_IN_MUL)) ? 1 : -1;
#endif
}

// Threads of one CTA and transforms per CTA for row transforms of length n.
__host__ __device__ constexpr int rows_threads_per_transform(int n, int v = 0) { return n / engine_regs(n, v); }
__host__ __device__ constexpr int rows_per_cta(int n, int v = 0)
{
    return rows_threads_per_transform(n, v) >= engine_cta_threads(n, v) ? 1 : engine_cta_threads(n, v) / rows_threads_per_transform(n, v);
}
__host__ __device__ constexpr int rows_cta_threads(int n, int v = 0) { return rows_threads_per_transform(n, v) * rows_per_cta(n, v); }
// Exchange area per transform (also used by the R2C epilogue, so never zero).
__host__ __device__ constexpr int rows_smem_elems(int n, int v = 0) { return engine_smem_elems(n, v); }
__host__ __device__ constexpr size_t rows_smem_bytes(int n, int v = 0) { return (size_t)rows_per_cta(n, v) * rows_smem_elems(n, v) * sizeof(float2); }

// C2R resolve of one bin pair (ref kernel.c:36-45): given X[k], X[N-k] and rtw[k] returns z[k] and,
// through `zm`, z[N-k] (rtw[N-k] = conj rtw[k], so the second costs no extra twiddle).
__device__ __forceinline__ cpx c2r_pair(cpx xa, cpx xb, cpx w, cpx &zm)
{
    const cpx cb = cconj(xb);
    const cpx a = cadd(xa, cb);
    const cpx b = cmul(csub(xa, cb), w);
    zm = cconj(csub(a, b));
    return cadd(a, b);
}

// Partitioned convolution (MCU_IN_MULACC*): the spectral product is a sum over `parts` pairs of
// (block spectrum, filter-partition spectrum), `in_step` / `in2_step` complex elements apart.
struct rows_parts {
    unsigned parts;
    long long in_step, in2_step;
};

// sum_p in[p*in_step + k] * in2[p*in2_step + k]
template <typename Ptr>
__device__ __forceinline__ cpx mulacc_at(Ptr in, const cpx *__restrict__ in2, int k, const rows_parts &pp)
{
    cpx acc = make_float2(0.0f, 0.0f);
#pragma unroll 2
    for (unsigned p = 0; p < pp.parts; p++) {
        const cpx x = in[(long long)p * pp.in_step + k];
        const cpx w = __ldg(in2 + (long long)p * pp.in2_step + k);
        acc.x = fmaf(x.x, w.x, fmaf(-x.y, w.y, acc.x));
        acc.y = fmaf(x.x, w.y, fmaf(x.y, w.x, acc.y));
    }
    return acc;
}

// ---- prologue: registers <- input ------------------------------------------------------------------
// Variant 0 / 1: element k = t + e*T in slot e.  Variant 2: the slots of the mirrored first pass.
template <int N, int IN_MODE, int V, typename Ptr>
__device__ __forceinline__ void rows_load(cpx (&v)[engine_regs(N, V)], Ptr in, const cpx *__restrict__ in2,
                                          const cpx *__restrict__ rtw, float norm, int t, const rows_parts pp = { 1, 0, 0 })
{
    constexpr int R = engine_regs(N, V);
    constexpr int T = N / R;
    if constexpr (IN_MODE == MCU_IN_COMPLEX) {
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = in[t + e * T];
    } else if constexpr (IN_MODE == MCU_IN_HALF) {
        // Slots e >= R/2 are exactly the elements >= N/2: zero, and their memory is not touched.
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = e < R / 2 ? in[t + e * T] : make_float2(0.0f, 0.0f);
    } else if constexpr (IN_MODE == MCU_IN_MUL) {
        // the block spectrum (HBM) first, all loads in flight; the filter bins (L1 / L2 hits: one filter for the whole
        // batch) are then fetched as they are used - holding all R of them as well would need 2 R register pairs
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = in[t + e * T];
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = cscale(norm, cmul(v[e], __ldg(in2 + t + e * T)));
    } else if constexpr (IN_MODE == MCU_IN_MULACC) {
        // partition loop outermost: every partition is one coalesced sweep over the block spectrum
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = make_float2(0.0f, 0.0f);
#pragma unroll 1
        for (unsigned p = 0; p < pp.parts; p++) {
            const cpx *__restrict__ f = in2 + (long long)p * pp.in2_step;
            cpx w[R];
#pragma unroll
            for (int e = 0; e < R; e++) w[e] = __ldg(f + t + e * T);
#pragma unroll
            for (int e = 0; e < R; e++) {
                const cpx x = in[(long long)p * pp.in_step + t + e * T];
                v[e].x = fmaf(x.x, w[e].x, fmaf(-x.y, w[e].y, v[e].x));
                v[e].y = fmaf(x.x, w[e].y, fmaf(x.y, w[e].x, v[e].y));
            }
        }
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = cscale(norm, v[e]);
    } else if constexpr (V == MUFFT_VARIANT_MIRROR_FIRST) {
        // C2R resolve, optionally on the product of two spectra, feeding the mirrored first pass:
        // group u holds butterfly i0 (bins i0 + m*Q0), group u + H its image Q0 - i0 (bins
        // Q0 - i0 + m*Q0); bin N - (i0 + m*Q0) is input r-1-m of the image.
        constexpr bool mul = IN_MODE == MCU_IN_MUL_C2R;
        constexpr bool mulacc = IN_MODE == MCU_IN_MULACC_C2R;
        constexpr int r = engine_radix(N, 0, V);
        constexpr int NB = R / r;
        constexpr int H = NB / 2;
        constexpr int Q0 = N / r;
        auto fetch = [&](int k) {
            if constexpr (mulacc) {
                return cscale(norm, mulacc_at(in, in2, k, pp));
            } else {
                cpx x = in[k];
                if constexpr (mul) x = cscale(norm, cmul(x, __ldg(in2 + k)));
                return x;
            }
        };
#pragma unroll
        for (int u = 0; u < H; u++) {
            const int i0 = t + u * T;
            if (u == 0 && i0 == 0) {
                // butterflies 0 and Q0/2 are their own images: bin m*Q0 pairs with (r-m)*Q0 (and 0
                // with N), bin Q0/2 + m*Q0 with Q0/2 + (r-1-m)*Q0
                cpx a[r + 1], b[r];
#pragma unroll
                for (int m = 0; m <= r; m++) a[m] = fetch(m * Q0);
#pragma unroll
                for (int m = 0; m < r; m++) b[m] = fetch(Q0 / 2 + m * Q0);
#pragma unroll
                for (int m = 0; m <= r / 2; m++) {
                    cpx zm;
                    const cpx z = c2r_pair(a[m], a[r - m], __ldg(rtw + m * Q0), zm);
                    v[m * NB] = z;
                    if (m != 0 && m != r - m) v[(r - m) * NB] = zm;
                }
#pragma unroll
                for (int m = 0; m < r / 2; m++) {
                    cpx zm;
                    const cpx z = c2r_pair(b[m], b[r - 1 - m], __ldg(rtw + Q0 / 2 + m * Q0), zm);
                    v[H + m * NB] = z;
                    v[H + (r - 1 - m) * NB] = zm;
                }
            } else {
                cpx a[r], b[r];
#pragma unroll
                for (int m = 0; m < r; m++) a[m] = fetch(i0 + m * Q0);
#pragma unroll
                for (int m = 0; m < r; m++) b[m] = fetch(Q0 - i0 + m * Q0);
#pragma unroll
                for (int m = 0; m < r; m++) {
                    cpx zm;
                    v[u + m * NB] = c2r_pair(a[m], b[r - 1 - m], __ldg(rtw + i0 + m * Q0), zm);
                    v[u + H + (r - 1 - m) * NB] = zm;
                }
            }
        }
    } else {
        // C2R resolve, optionally on the product of two spectra.  Reads bins k and N-k.
        constexpr bool mul = IN_MODE == MCU_IN_MUL_C2R;
        constexpr bool mulacc = IN_MODE == MCU_IN_MULACC_C2R;
#pragma unroll
        for (int e = 0; e < R; e++) {
            const int k = t + e * T;
            cpx xa, xb;
            if constexpr (mulacc) {
                xa = cscale(norm, mulacc_at(in, in2, k, pp));
                xb = cscale(norm, mulacc_at(in, in2, N - k, pp));
            } else {
                xa = in[k];
                xb = in[N - k];
            }
            if constexpr (mul) {
                xa = cscale(norm, cmul(xa, __ldg(in2 + k)));
                xb = cscale(norm, cmul(xb, __ldg(in2 + N - k)));
            }
            xb = cconj(xb);
            v[e] = cadd(cadd(xa, xb), cmul(csub(xa, xb), __ldg(rtw + k)));
        }
    }
}

// R2C resolve of one bin pair (ref kernel.c:47-82): from Z[k], Z[N-k] and rtw[k], X[k] and X[N-k]
// (rtw[N-k] = conj rtw[k]); with FULL also X[k+N] and X[2N-k].
// Where a row transform's spectrum goes: a plain row, or (exchange step of a slab R2C transform)
// the column slabs of several devices, bin k to device min(k >> shift, last).
struct spectrum_row {
    cpx *__restrict__ out;
    __device__ __forceinline__ void operator()(int k, cpx v) const { out[k] = v; }
};
struct spectrum_peers {
    const mcu_peer_dest &d;
    size_t b; // row of this transform inside the slabs
    __device__ __forceinline__ void operator()(int k, cpx v) const
    {
        unsigned p = (unsigned)k >> d.shift;
        p = p > d.last ? d.last : p;
        const size_t pitch = p == d.last ? d.pitch_last : d.pitch;
        static_cast<cpx *>(d.base[p])[(d.row0 + b) * pitch + ((unsigned)k - (p << d.shift))] = v;
    }
};

template <bool FULL, typename Store>
__device__ __forceinline__ void r2c_pair_store(const Store &store, int k, int n, cpx za, cpx zb, cpx w)
{
    const cpx cb = cconj(zb);
    const cpx fe = cadd(za, cb);
    const cpx fo = cmul(w, csub(za, cb));
    const cpx xk = cscale(0.5f, cadd(fe, fo));
    const cpx xn = cscale(0.5f, csub(fe, fo)); // = conj X[N-k] = X[k+N]
    store(k, xk);
    if (k != n - k) store(n - k, cconj(xn));
    if constexpr (FULL) {
        store(k + n, xn);
        if (k != n - k) store(2 * n - k, cconj(xk));
    }
}

// Exchange-step epilogue (MCU_OUT_PEERS): element x of transform b goes to the column slab of
// device x >> shift.  Consecutive lanes hold consecutive x, so every warp store is a contiguous
// run inside one device's slab (local memory or NVLink peer memory).
template <int N, int V>
__device__ __forceinline__ void rows_store_peers(const cpx (&v)[engine_regs(N, V)], const mcu_peer_dest &d, size_t b, int t)
{
    constexpr int R = engine_regs(N, V);
    constexpr int T = N / R;
    const size_t row = (d.row0 + b) * d.pitch;
    const unsigned mask = (1u << d.shift) - 1u;
#pragma unroll
    for (int e = 0; e < R; e++) {
        const unsigned x = (unsigned)(t + e * T);
        cpx *dst = static_cast<cpx *>(d.base[x >> d.shift]) + row + (x & mask);
        *dst = v[e];
    }
}

// ---- epilogue: output <- registers -----------------------------------------------------------------
// R2C epilogues (half or full spectrum) writing through `store(k, value)`: spectrum_row for a plain
// output row, spectrum_peers for the exchange step of a slab transform.
template <int N, int OUT_MODE, int T_SYNC, int SLOTS, int V, typename Store>
__device__ __forceinline__ void rows_store_spectrum(cpx (&v)[engine_regs(N, V)], const Store &store, cpx *sm,
                                                    const cpx *__restrict__ rtw, int t, int slot, bool active)
{
    constexpr int R = engine_regs(N, V);
    constexpr int T = N / R;
    constexpr int SH = engine_pad_shift(N, V);
    if constexpr (V == MUFFT_VARIANT_MIRROR_LAST) {
        // The mirrored last pass left Z[i0 + q*Q0] in slot u + q*NB and Z[Q0 - i0 + q*Q0] in slot
        // u + H + q*NB; the partner of bin i0 + q*Q0 is output r-1-q of the image butterfly.
        constexpr bool full = OUT_MODE == MCU_OUT_R2C_FULL;
        constexpr int r = engine_radix(N, engine_num_passes(N, V) - 1, V);
        constexpr int NB = R / r;
        constexpr int H = NB / 2;
        constexpr int Q0 = N / r;
        if (!active) return;
#pragma unroll
        for (int u = 0; u < H; u++) {
            const int i0 = t + u * T;
            if (u == 0 && i0 == 0) {
                // butterflies 0 and Q0/2 are their own images
                store(0, make_float2(v[0].x + v[0].y, 0.0f));
                store(N, make_float2(v[0].x - v[0].y, 0.0f));
                if constexpr (full) {
                    // bin 2N - 0 does not exist; X[N] is written above
                }
#pragma unroll
                for (int q = 1; q <= r / 2; q++)
                    r2c_pair_store<full>(store, q * Q0, N, v[q * NB], v[(r - q) * NB], __ldg(rtw + q * Q0));
#pragma unroll
                for (int q = 0; q < r / 2; q++)
                    r2c_pair_store<full>(store, Q0 / 2 + q * Q0, N, v[H + q * NB], v[H + (r - 1 - q) * NB], __ldg(rtw + Q0 / 2 + q * Q0));
            } else {
#pragma unroll
                for (int q = 0; q < r; q++)
                    r2c_pair_store<full>(store, i0 + q * Q0, N, v[u + q * NB], v[u + H + (r - 1 - q) * NB], __ldg(rtw + i0 + q * Q0));
            }
        }
    } else {
        // X[k] = 1/2 ((Z[k] + conj Z[N-k]) + rtw[k] (Z[k] - conj Z[N-k])); Z[N-k] lives in another
        // thread, so the half-length spectrum makes one more trip through shared memory.
        constexpr bool full = OUT_MODE == MCU_OUT_R2C_FULL;
        if constexpr (engine_num_passes(N, V) > 1) engine_sync<T_SYNC, SLOTS>(slot);
        engine_scatter_natural<N, 1, V>(v, sm, t);
        engine_sync<T_SYNC, SLOTS>(slot);
        if (active) {
#pragma unroll
            for (int e = 0; e < R; e++) {
                const int k = t + e * T;
                if (k == 0) {
                    store(0, make_float2(v[e].x + v[e].y, 0.0f));
                    store(N, make_float2(v[e].x - v[e].y, 0.0f));
                } else {
                    const cpx za = v[e];
                    const cpx zb = cconj(sm[pad_pos<SH>(N - k)]);
                    const cpx fe = cadd(za, zb);
                    const cpx fo = cmul(__ldg(rtw + k), csub(za, zb));
                    store(k, cscale(0.5f, cadd(fe, fo)));
                    if constexpr (full) store(k + N, cscale(0.5f, csub(fe, fo)));
                }
            }
        }
    }
}

template <int N, int OUT_MODE, int T_SYNC, int SLOTS, int V>
__device__ __forceinline__ void rows_store(cpx (&v)[engine_regs(N, V)], cpx *__restrict__ out, cpx *sm,
                                           const cpx *__restrict__ rtw, int t, int slot, bool active)
{
    if constexpr (OUT_MODE == MCU_OUT_COMPLEX) {
        constexpr int R = engine_regs(N, V);
        constexpr int T = N / R;
        if (active) {
#pragma unroll
            for (int e = 0; e < R; e++) out[t + e * T] = v[e];
        }
    } else {
        rows_store_spectrum<N, OUT_MODE, T_SYNC, SLOTS, V>(v, spectrum_row{ out }, sm, rtw, t, slot, active);
    }
}

// ---- the kernel: one group per CTA, direct loads ---------------------------------------------------
template <int N, int DIR, int IN_MODE, int OUT_MODE>
__global__ void __launch_bounds__(rows_cta_threads(N, rows_variant(N, IN_MODE, OUT_MODE)), engine_min_blocks(N, rows_variant(N, IN_MODE, OUT_MODE)))
fft_rows_kernel(const mcu_rows_args a)
{
    constexpr int V = rows_variant(N, IN_MODE, OUT_MODE);
    constexpr int R = engine_regs(N, V);
    constexpr int T = N / R;
    constexpr int TPC = rows_per_cta(N, V);
    MUFFT_DYNAMIC_SMEM(smem_raw);

    const int tid = threadIdx.x;
    const int slot = tid / T;
    const int t = tid - slot * T;
    size_t b = (size_t)blockIdx.x * TPC + slot;
    const bool active = b < a.batch;
    if (!active) b = a.batch - 1; // keep the thread in step with the barriers; it stores nothing

    cpx *sm = reinterpret_cast<cpx *>(smem_raw) + slot * rows_smem_elems(N, V);
    const cpx *__restrict__ tw = static_cast<const cpx *>(a.tw);
    const cpx *__restrict__ rtw = static_cast<const cpx *>(a.rtw);
    const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + b * a.in_stride;
    const cpx *__restrict__ in2 = static_cast<const cpx *>(a.in2) + b * a.in2_stride;

    cpx v[R];
    if constexpr (IN_MODE == MCU_IN_MULACC || IN_MODE == MCU_IN_MULACC_C2R)
        rows_load<N, IN_MODE, V>(v, in, static_cast<const cpx *>(a.in2), rtw, a.norm, t, rows_parts{ a.parts, a.in_part_stride, a.in2_part_stride });
    else
        rows_load<N, IN_MODE, V>(v, in, in2, rtw, a.norm, t);
    engine_run<N, DIR, 1, T, TPC, 0, V, rows_packed(N, IN_MODE, OUT_MODE)>(v, sm, tw, t, slot);
    if constexpr (OUT_MODE == MCU_OUT_PEERS) {
        if (active) rows_store_peers<N, V>(v, a.peers, b, t);
    } else if constexpr (OUT_MODE == MCU_OUT_R2C_PEERS) {
        rows_store_spectrum<N, MCU_OUT_R2C, T, TPC, V>(v, spectrum_peers{ a.peers, b }, sm, rtw, t, slot, active);
    } else {
        cpx *__restrict__ out = static_cast<cpx *>(a.out) + b * a.out_stride;
        rows_store<N, OUT_MODE, T, TPC, V>(v, out, sm, rtw, t, slot, active);
    }
}

} // namespace mufft
