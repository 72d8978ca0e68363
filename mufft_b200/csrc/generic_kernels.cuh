// Kernels used beyond the engine limits and by mufft_get_convolve_func():
//   * stockham_pass_kernel   one radix-8/4/2 Stockham pass, any length, 1D or vertical
//                            (the role of ref kernel.c:84-453 and 455-738)
//   * resolve_kernel         stand-alone R2C / C2R resolve and spectral multiply (ref kernel.c:22-82)
// In a header so that tests/emu can run them on the CPU emulation as well.
#pragma once

#include "abi.h"
#include "fft_engine.cuh"

namespace mufft {

// ---- generic Stockham pass ----------------------------------------------------------------
template <int RADIX, int DIR>
__global__ void __launch_bounds__(256) stockham_pass_kernel(const mcu_pass_args a)
{
    const size_t q = a.n / RADIX;
    const size_t total = q * a.width;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const size_t x = idx % a.width;
    const size_t i = idx / a.width;
    const size_t k = i & (size_t)(a.p - 1);
    const size_t base = (i - k) * RADIX + k;
    const size_t step = a.n / ((size_t)RADIX * a.p);
    const cpx *__restrict__ roots = static_cast<const cpx *>(a.roots);

    for (size_t b = blockIdx.y; b < a.batch; b += gridDim.y) {
        const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + b * a.in_batch + x;
        cpx *__restrict__ out = static_cast<cpx *>(a.out) + b * a.out_batch + x;
        cpx v[RADIX];
#pragma unroll
        for (int m = 0; m < RADIX; m++)
            v[m] = (a.half && m >= RADIX / 2) ? make_float2(0.0f, 0.0f) : in[(i + m * q) * a.in_stride];
        if (a.p > 1) {
#pragma unroll
            for (int m = 1; m < RADIX; m++) v[m] = cmul(v[m], __ldg(roots + k * m * step));
        }
        dft_regs<RADIX, DIR>(v);
#pragma unroll
        for (int r = 0; r < RADIX; r++) out[(base + (size_t)r * a.p) * a.out_stride] = v[r];
    }
}

// ---- stand-alone resolve / multiply -----------------------------------------------------
__global__ void __launch_bounds__(256) resolve_kernel(const mcu_resolve_args a)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n) return;
    const unsigned n = a.n;
    const cpx *__restrict__ rtw = static_cast<const cpx *>(a.rtw);
    for (size_t b = blockIdx.y; b < a.batch; b += gridDim.y) {
        const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + b * a.in_stride;
        const cpx *__restrict__ in2 = a.in2 ? static_cast<const cpx *>(a.in2) + b * a.in2_stride : in;
        cpx *__restrict__ out = static_cast<cpx *>(a.out) + b * a.out_stride;
        if (a.mode == MCU_RES_MUL) {
            out[k] = cscale(a.norm, cmul(in[k], in2[k]));
        } else if (a.mode == MCU_RES_C2R || a.mode == MCU_RES_MUL_C2R) {
            cpx xa = in[k], xb = in[n - k];
            if (a.mode == MCU_RES_MUL_C2R) {
                xa = cscale(a.norm, cmul(xa, in2[k]));
                xb = cscale(a.norm, cmul(xb, in2[n - k]));
            }
            xb = cconj(xb);
            out[k] = cadd(cadd(xa, xb), cmul(csub(xa, xb), __ldg(rtw + k)));
        } else {
            const cpx za = in[k];
            if (k == 0) {
                out[0] = make_float2(za.x + za.y, 0.0f);
                out[n] = make_float2(za.x - za.y, 0.0f);
            } else {
                const cpx zb = cconj(in[n - k]);
                const cpx fe = cadd(za, zb);
                const cpx fo = cmul(__ldg(rtw + k), csub(za, zb));
                out[k] = cscale(0.5f, cadd(fe, fo));
                if (a.mode == MCU_RES_R2C_FULL) out[k + n] = cscale(0.5f, csub(fe, fo));
            }
        }
    }
}

} // namespace mufft
