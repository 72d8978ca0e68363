// Batched C2C row transforms of 8192 points, warp-specialised: ONE persistent CTA per SM, transforms moved by the TMA unit
// (see fft2d_tma_kernel.cuh).  fft_rows_kernel<8192, ...> holds a transform in registers from its loads to its stores (32
// values x 256 threads at 128 registers, two CTAs per SM, each waiting for its own loads): 0.83 of the HBM roofline; this
// kernel 0.88 (profiles/r02aa_ab.txt).  The convolution stages (zero-padded load, spectral multiply) were measured too and
// stay on the row kernel: 0.66 - 0.675 here against 0.71 - the filter spectrum no longer fits the L1 that the three
// buffers leave.  Twiddles: power-of-two entries loaded, the rest derived (the full table: 0.79).
//
//   loader   (1 thread)   transform k = blockIdx.x + s * gridDim.x for the CTA's s-th buffer use: waits for a free buffer,
//                         one bulk copy of the transform (half of it for a zero-padded input) -> `full`
//   groups   (2 x 256)    wait for `full`, pull the transform into registers in natural order (the load stage of the row
//                         kernel - zero padding, spectral multiply - reading shared memory instead of global), run the
//                         engine with the buffer as exchange area, put the result back, fence.proxy.async -> `staged`
//   storers  (3 threads)  one per buffer, a warp each: bulk store, wait until the buffer has been read -> `freed`
//
// No tickets and no dependencies: the assignment is static.  End of work: six EXIT entries.
#pragma once

#include "rows_kernel.cuh"
#include "tma_util.cuh"

namespace mufft {

#define ROWS_TMA_GROUPS 2
#define ROWS_TMA_STAGES 3
#define ROWS_TMA_THREADS(n) (ROWS_TMA_GROUPS * rows_threads_per_transform(n) + 32 * (1 + ROWS_TMA_STAGES))
enum { ROWS_TMA_WORK = 0, ROWS_TMA_EXIT = 2 };

__host__ __device__ constexpr bool rows_tma_supported(int n, int in_mode, int out_mode)
{
    return n == 8192 && engine_regs(n) > 0 && rows_threads_per_transform(n) == 256 && out_mode == MCU_OUT_COMPLEX &&
           in_mode == MCU_IN_COMPLEX;
}
__host__ __device__ constexpr size_t rows_tma_buf_bytes(int n) { return ((size_t)engine_smem_elems(n, 0) * sizeof(float2) + 127) / 128 * 128; }
__host__ __device__ constexpr size_t rows_tma_smem_bytes(int n) { return ROWS_TMA_STAGES * rows_tma_buf_bytes(n) + (ROWS_TMA_GROUPS + 2) * ROWS_TMA_STAGES * 8 + ROWS_TMA_STAGES * 16; }

#ifndef MUFFT_EMU
#ifndef MUFFT_ROWS_TMA_POLL_NS
#define MUFFT_ROWS_TMA_POLL_NS 100
#endif
#ifndef MUFFT_ROWS_TMA_TWM
#define MUFFT_ROWS_TMA_TWM MUFFT_TW_DERIVE // twiddle mode of the engine passes (-1 = the shape's own)
#endif

template <int N, int DIR, int IN_MODE>
__global__ void __launch_bounds__(ROWS_TMA_THREADS(N), 1)
fft_rows_tma_kernel(const mcu_rows_args a)
{
    static_assert(rows_tma_supported(N, IN_MODE, MCU_OUT_COMPLEX), "shape / mode");
    constexpr int R = engine_regs(N, 0), T = N / R;
    constexpr size_t BUF = rows_tma_buf_bytes(N);
    constexpr unsigned IN_BYTES = (IN_MODE == MCU_IN_HALF ? N / 2 : N) * sizeof(cpx), OUT_BYTES = N * sizeof(cpx);
    constexpr int UTIL0 = ROWS_TMA_GROUPS * T;
    constexpr int PK = rows_packed(N, IN_MODE, MCU_OUT_COMPLEX);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *base = smem_raw;
    uint4 *meta = reinterpret_cast<uint4 *>(base + ROWS_TMA_STAGES * BUF); // x = type, (y, z) = transform index
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(meta + ROWS_TMA_STAGES);
    unsigned long long *full = bars, *staged = bars + ROWS_TMA_GROUPS * ROWS_TMA_STAGES, *freed = staged + ROWS_TMA_STAGES;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int b = 0; b < ROWS_TMA_STAGES; b++) {
            for (int g = 0; g < ROWS_TMA_GROUPS; g++) mbar_init(full + b * ROWS_TMA_GROUPS + g, 1);
            mbar_init(staged + b, T);
            mbar_init(freed + b, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async_all();
    }
    __syncthreads(); // the only CTA-wide barrier

    if (tid < UTIL0) {
        const int group = tid / T, t = tid % T;
        const cpx *__restrict__ tw = static_cast<const cpx *>(a.tw);
        int exits = 0;
        for (unsigned s = group;; s += ROWS_TMA_GROUPS) {
            const unsigned b = s % ROWS_TMA_STAGES;
            [[maybe_unused]] const unsigned use = s / ROWS_TMA_STAGES;
            cpx *buf = reinterpret_cast<cpx *>(base + b * BUF);
            MUFFT_TMA_WAIT(0, full + b * ROWS_TMA_GROUPS + group, (s / (ROWS_TMA_GROUPS * ROWS_TMA_STAGES)) & 1, "full", b, use);
            const uint4 m = meta[b];
            if (m.x == ROWS_TMA_EXIT) {
                mbar_arrive(staged + b);
                if (++exits == ROWS_TMA_STAGES) break;
                continue;
            }
            const size_t k = ((size_t)m.z << 32) | m.y;
            cpx v[R];
            rows_load<N, IN_MODE, 0>(v, static_cast<const cpx *>(buf), static_cast<const cpx *>(a.in2) + k * a.in2_stride, static_cast<const cpx *>(nullptr), a.norm, t);
            engine_sync<T, ROWS_TMA_GROUPS>(group); // the transform is in registers before the exchange overwrites it
            engine_run<N, DIR, 1, T, ROWS_TMA_GROUPS, 0, 0, PK, MUFFT_ROWS_TMA_TWM>(v, buf, tw, t, group);
            engine_sync<T, ROWS_TMA_GROUPS>(group); // every thread has gathered
#pragma unroll
            for (int e = 0; e < R; e++) buf[t + e * T] = v[e];
            fence_proxy_async_smem();
            mbar_arrive(staged + b);
        }
    } else if (tid == UTIL0) {
        const unsigned long long pol = l2_policy<1>(); // read once
        int exits = 0;
        for (unsigned s = 0;; s++) {
            const unsigned b = s % ROWS_TMA_STAGES, use = s / ROWS_TMA_STAGES;
            unsigned long long *fb = full + b * ROWS_TMA_GROUPS + s % ROWS_TMA_GROUPS;
            const size_t k = (size_t)blockIdx.x + (size_t)s * gridDim.x;
            MUFFT_TMA_WAIT(MUFFT_ROWS_TMA_POLL_NS, freed + b, (use & 1) ^ 1, "free", b, use);
            if (k >= a.batch) {
                meta[b] = make_uint4(ROWS_TMA_EXIT, 0, 0, 0);
                mbar_arrive(fb);
                if (++exits == ROWS_TMA_GROUPS * ROWS_TMA_STAGES) break;
                continue;
            }
            meta[b] = make_uint4(ROWS_TMA_WORK, (unsigned)k, (unsigned)(k >> 32), 0);
            mbar_arrive_expect_tx(fb, IN_BYTES);
            bulk_load_1d<1>(base + b * BUF, static_cast<const cpx *>(a.in) + k * a.in_stride, IN_BYTES, fb, pol);
        }
    } else if (tid >= UTIL0 + 32 && (tid & 31) == 0) {
        const unsigned b = (tid - (UTIL0 + 32)) / 32;
        const unsigned long long pol = l2_policy<1>(); // streaming
        int exits = 0;
        for (unsigned use = 0;; use++) {
            MUFFT_TMA_WAIT(MUFFT_ROWS_TMA_POLL_NS, staged + b, use & 1, "staged", b, use);
            const uint4 m = meta[b];
            if (m.x == ROWS_TMA_EXIT) {
                mbar_arrive(freed + b);
                if (++exits == ROWS_TMA_GROUPS) break;
                continue;
            }
            const size_t k = ((size_t)m.z << 32) | m.y;
            bulk_store_1d<1>(static_cast<cpx *>(a.out) + k * a.out_stride, base + b * BUF, OUT_BYTES, pol);
            bulk_commit();
            bulk_wait_read();
            mbar_arrive(freed + b);
        }
        bulk_wait_all();
    }
}
#endif // MUFFT_EMU

} // namespace mufft
