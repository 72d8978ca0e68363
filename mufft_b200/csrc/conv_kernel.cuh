// Whole FFT convolution of a block in ONE kernel: load -> forward transform -> x filter spectrum
// (with the 1/N) -> inverse transform -> store.  The block's spectrum never leaves the SM, so a
// stereo block moves 4N (zero padded) or 8N bytes in and 8N out instead of the 28N bytes of the
// reference's call sequence
//     mufft_execute_conv_input(plan, 0, block0, in)   forward FFT            (ref fft.c:799-802)
//     mufft_execute_conv_output(plan, out, block0, f) multiply + inverse FFT (ref fft.c:804-809)
// The filter spectrum `f` is what mufft_execute_conv_input(plan, 1, f, taps) produced, unchanged:
// FULL_R2C spectrum of N bins for stereo (ref fft.c:575-577), N/2+1 bins for mono (fft.c:589-590).
//
// The engine keeps element t + e*T in register slot e before AND after a transform
// (fft_engine.cuh), so the inverse runs straight on the registers the forward left:
//   stereo  y[k] = norm * X[k] * F[k]                                   (ref kernel.c:22-34)
//   mono    Z = FFT_n(reals as n = N/2 complex); bins k and n-k are resolved to X[k], X[n-k]
//           (ref kernel.c:47-63), multiplied by F, and folded back to the half-length spectrum
//           z[k] (ref kernel.c:36-45), all in the registers of the thread that owns k; only
//           Z[n-k] is fetched from the exchange area.
#pragma once

#include "abi.h"
#include "fft_engine.cuh"
#include "rows_kernel.cuh"

namespace mufft {

template <int N, int IN_MODE, bool MONO>
__global__ void __launch_bounds__(rows_cta_threads(N, 0), engine_min_blocks(N, 0))
fft_conv_fused_kernel(const mcu_conv_args a)
{
    constexpr int R = engine_regs(N, 0);
    constexpr int T = N / R;
    constexpr int TPC = rows_per_cta(N, 0);
    constexpr int SH = engine_pad_shift(N, 0);
    constexpr int PK = N == 8192 ? 1 : -1; // packed complex adds: measured 2.79 -> 2.60 ms on the stereo N = 8192 workload
    MUFFT_DYNAMIC_SMEM(smem_raw);

    const int tid = threadIdx.x;
    const int slot = tid / T;
    const int t = tid - slot * T;
    size_t b = (size_t)blockIdx.x * TPC + slot;
    const bool active = b < a.batch;
    if (!active) b = a.batch - 1;

    cpx *sm = reinterpret_cast<cpx *>(smem_raw) + slot * rows_smem_elems(N, 0);
    const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + b * a.in_stride;
    const cpx *__restrict__ filt = static_cast<const cpx *>(a.filter) + b * a.filter_stride;

    cpx v[R];
    rows_load<N, IN_MODE, 0>(v, in, nullptr, nullptr, 1.0f, t);
    engine_run<N, -1, 1, T, TPC, 0, 0, PK>(v, sm, static_cast<const cpx *>(a.tw_fwd), t, slot);

    if constexpr (!MONO) {
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = cscale(a.norm, cmul(v[e], __ldg(filt + t + e * T)));
        if constexpr (engine_num_passes(N, 0) > 1) engine_sync<T, TPC>(slot); // forward's last gather vs inverse's first scatter
    } else {
        const cpx *__restrict__ rf = static_cast<const cpx *>(a.rtw_fwd);
        const cpx *__restrict__ ri = static_cast<const cpx *>(a.rtw_inv);
        if constexpr (engine_num_passes(N, 0) > 1) engine_sync<T, TPC>(slot);
        engine_scatter_natural<N, 1, 0>(v, sm, t);
        engine_sync<T, TPC>(slot);
#pragma unroll
        for (int e = 0; e < R; e++) {
            const int k = t + e * T;
            cpx ya, yb; // Y[k], conj Y[n-k]
            if (k == 0) {
                const cpx f0 = __ldg(filt), fn = __ldg(filt + N);
                const float x0 = v[e].x + v[e].y, xn = v[e].x - v[e].y;
                ya = cscale(a.norm * x0, f0);
                yb = cconj(cscale(a.norm * xn, fn));
            } else {
                const cpx za = v[e];
                const cpx cb = cconj(sm[pad_pos<SH>(N - k)]);
                const cpx fe = cadd(za, cb);
                const cpx fo = cmul(__ldg(rf + k), csub(za, cb));
                const cpx xk = cscale(0.5f, cadd(fe, fo));  // X[k]
                const cpx xn = cscale(0.5f, csub(fe, fo));  // conj X[n-k]
                ya = cscale(a.norm, cmul(xk, __ldg(filt + k)));
                yb = cscale(a.norm, cmul(xn, cconj(__ldg(filt + N - k)))); // conj(X[n-k] F[n-k])
            }
            v[e] = cadd(cadd(ya, yb), cmul(csub(ya, yb), __ldg(ri + k)));
        }
        engine_sync<T, TPC>(slot); // every partner read is done before the inverse scatters
    }

    engine_run<N, 1, 1, T, TPC, 0, 0, PK>(v, sm, static_cast<const cpx *>(a.tw_inv), t, slot);
    if (active) {
        cpx *__restrict__ out = static_cast<cpx *>(a.out) + b * a.out_stride;
#pragma unroll
        for (int e = 0; e < R; e++) out[t + e * T] = v[e];
    }
}

// (The stereo kernel made persistent, with the next block's input fetched by the TMA unit into the exchange area while
// the last inverse pass and the stores run - the scheme of rows_prefetch_kernel.cuh - was measured in round 2: 0.677 ms
// against 0.666 ms per 16384 blocks of 8192 points, 0.605 against 0.578 at 4096 points.  This kernel waits for issue
// slots, not for its loads.  profiles/r02ae_ab_conv_prefetch_rejected.txt.)

} // namespace mufft
