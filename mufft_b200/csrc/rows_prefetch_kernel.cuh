// Batched row transforms of 8192 and 16384 points (C2C, and the convolution stages: zero-padded load, spectral multiply):
// the row kernel's shapes - one CTA per transform, 32 values per thread, 66 / 132 KB of exchange area, two / one CTA per
// SM - made PERSISTENT, with the next transform fetched by the TMA unit while this one finishes.
//
// fft_rows_kernel<16384, ...> is alone on its SM: it loads 128 KB, computes three passes, stores 128 KB, and nothing
// overlaps - 0.62 of the HBM roofline.  Here the exchange area doubles as the landing buffer of the next transform: it is
// idle from the gather of the last exchange to the scatter of the next transform's first one, which is when the last
// pass and the stores of this transform run.  Measured (profiles/r02ac_ab.txt, r02ad_ab.txt): 16384 points 0.62 -> 0.77 of
// the HBM roofline, 8192 points 0.83 -> 0.92, the two stages of the stereo convolution at 8192 points 0.71 -> 0.735.
// (A warp-specialised version for 8192 points - loader, two compute groups, storers, three buffers, as
// fft2d_tma_kernel.cuh - reached 0.88 and lost on the convolution stages: profiles/r02aa_ab.txt.  Gone.)
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

__host__ __device__ constexpr bool rows_prefetch_supported(int n, int in_mode, int out_mode)
{
    return (n == 8192 || n == 16384) && engine_regs(n) > 0 && rows_per_cta(n, 0) == 1 && engine_num_passes(n, 0) >= 2 && out_mode == MCU_OUT_COMPLEX &&
           (in_mode == MCU_IN_COMPLEX || in_mode == MCU_IN_HALF || in_mode == MCU_IN_MUL);
}
__host__ __device__ constexpr size_t rows_prefetch_smem_bytes(int n) { return ((size_t)engine_smem_elems(n, 0) * sizeof(float2) + 127) / 128 * 128 + 16; }

#ifndef MUFFT_EMU
template <int N, int DIR, int IN_MODE>
__global__ void __launch_bounds__(rows_cta_threads(N, 0), engine_min_blocks(N, 0))
fft_rows_prefetch_kernel(const mcu_rows_args a)
{
    static_assert(rows_prefetch_supported(N, IN_MODE, MCU_OUT_COMPLEX), "shape / mode");
    constexpr int R = engine_regs(N, 0), T = N / R, NP = engine_num_passes(N, 0);
    constexpr int PK = rows_packed(N, IN_MODE, MCU_OUT_COMPLEX);
    constexpr unsigned BYTES = (IN_MODE == MCU_IN_HALF ? N / 2 : N) * sizeof(cpx); // a zero-padded input is half as long
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx *buf = reinterpret_cast<cpx *>(smem_raw);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + rows_prefetch_smem_bytes(N) - 16);
    const int t = threadIdx.x;
    const cpx *__restrict__ tw = static_cast<const cpx *>(a.tw);
    const unsigned long long pol = l2_policy<1>(); // read once

    if (t == 0) {
        mbar_init(full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async_all();
        if (blockIdx.x < a.batch) {
            mbar_arrive_expect_tx(full, BYTES);
            bulk_load_1d<1>(buf, static_cast<const cpx *>(a.in) + (size_t)blockIdx.x * a.in_stride, BYTES, full, pol);
        }
    }
    __syncthreads();

    unsigned phase = 0;
    for (size_t k = blockIdx.x; k < a.batch; k += gridDim.x, phase ^= 1) {
        MUFFT_TMA_WAIT(0, full, phase, "full", 0, phase);
        cpx v[R];
        rows_load<N, IN_MODE, 0>(v, static_cast<const cpx *>(buf), static_cast<const cpx *>(a.in2) + k * a.in2_stride, static_cast<const cpx *>(nullptr), a.norm, t);
        __syncthreads(); // the transform is in registers before the exchange overwrites it
        static_for<NP - 1>([&](auto pc) {
            constexpr int PASS = decltype(pc)::value;
            engine_butterflies<N, DIR, PASS, 0, PK>(v, tw, t);
            if constexpr (PASS > 0) __syncthreads(); // previous gather complete
            engine_scatter<N, PASS, 1>(v, buf, t);
            __syncthreads();
            engine_gather<N, 1, PASS + 1>(v, buf, t);
        });
        __syncthreads(); // every thread has gathered: the buffer is free
        if (t == 0 && k + gridDim.x < a.batch) {
            fence_proxy_async_smem();
            mbar_arrive_expect_tx(full, BYTES);
            bulk_load_1d<1>(buf, static_cast<const cpx *>(a.in) + (k + gridDim.x) * a.in_stride, BYTES, full, pol);
        }
        engine_butterflies<N, DIR, NP - 1, 0, PK>(v, tw, t);
        cpx *__restrict__ out = static_cast<cpx *>(a.out) + k * a.out_stride;
#pragma unroll
        for (int e = 0; e < R; e++) __stcs(out + t + e * T, v[e]);
    }
}
#endif // MUFFT_EMU

} // namespace mufft
