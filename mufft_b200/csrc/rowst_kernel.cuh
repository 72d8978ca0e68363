// Second kernel of the two-kernel ("four-step") transform for lengths beyond one CTA's shared
// memory, N = N1 * N2:
//   kernel A = cols_kernel.cuh on the [N1][N2] view of the input: N2 column transforms of length N1
//              (Y[k1][n2] = sum_n1 x[n1*N2 + n2] w_N1^(n1 k1)), written to scratch in the same layout;
//   kernel B = this file: for a tile of 16 consecutive rows k1, multiply by the inter-stage twiddle
//              w_N^(k1 n2), run the N2-point row transforms on chip and store X[k1 + N1*k2].
// The output index has stride N1 along k2, so the last engine pass switches the thread mapping
// from "row fastest in t" to "row index fastest across lanes": each warp then stores 16
// consecutive k1 (128 bytes) per k2 instead of 8-byte pieces N1 apart.  The switch is free: the
// exchange between passes goes through shared memory anyway.
//
// Replaces, for N > 8192, the reference's log8(N) passes over the whole array
// (kernel.c:395-453 driven by fft.c:811-850) by two passes.
#pragma once

#include "abi.h"
#include "fft_engine.cuh"

namespace mufft {

#define MUFFT_ROWST_TILE 16
__host__ __device__ constexpr int rowst_cta_threads(int n2) { return MUFFT_ROWST_TILE * engine_threads(n2); }
__host__ __device__ constexpr size_t rowst_smem_bytes(int n2) { return (size_t)MUFFT_ROWST_TILE * engine_smem_elems(n2) * sizeof(float2); }

template <int N2, int DIR>
__global__ void __launch_bounds__(rowst_cta_threads(N2)) fft_rowst_kernel(const mcu_rowst_args a)
{
    constexpr int R = engine_regs(N2);
    constexpr int T = N2 / R;
    constexpr int TR = MUFFT_ROWST_TILE;
    constexpr int NP = engine_num_passes(N2);
    constexpr int SH = engine_pad_shift(N2);
    constexpr int SE = engine_smem_elems(N2);
    static_assert(NP >= 2, "the mapping switch needs an exchange");
    MUFFT_DYNAMIC_SMEM(smem_raw);
    cpx *smem = reinterpret_cast<cpx *>(smem_raw);

    const int tid = threadIdx.x;
    const int slot = tid / T, t = tid - slot * T;   // passes 0 .. NP-2: consecutive lanes walk along a row
    const int slot2 = tid % TR, t2 = tid / TR;      // last pass: consecutive lanes walk across the 16 rows
    const unsigned n1 = a.n1;
    const unsigned mask = a.n - 1;                  // N is a power of two
    const unsigned k1 = blockIdx.x * TR + slot;
    const cpx *__restrict__ tw = static_cast<const cpx *>(a.tw);
    const cpx *__restrict__ roots = static_cast<const cpx *>(a.roots);

    // inter-stage twiddle of this thread's elements: w_N^(k1*(t + e*T)) = w_N^(k1*t) * w_N^(k1*T*e);
    // the second factor is uniform over the threads of a row
    const cpx wa = __ldg(roots + ((k1 * (unsigned)t) & mask));

    for (size_t b = blockIdx.y; b < a.batch; b += gridDim.y) {
        const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + b * a.in_batch + (size_t)k1 * N2;
        cpx v[R];
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = in[t + e * T];
#pragma unroll
        for (int e = 0; e < R; e++) {
            const cpx wb = __ldg(roots + ((k1 * (unsigned)(T * e)) & mask));
            v[e] = cmul(v[e], cmul(wa, wb));
        }

        cpx *sm = smem + slot * SE;
        static_for<NP - 1>([&](auto pc) {
            constexpr int PASS = decltype(pc)::value;
            engine_butterflies<N2, DIR, PASS>(v, tw, t);
            if constexpr (PASS > 0) engine_sync<T, TR>(slot);
            engine_scatter<N2, PASS, 1>(v, sm, t);
            if constexpr (PASS + 1 < NP - 1) {
                engine_sync<T, TR>(slot);
                engine_gather<N2, 1>(v, sm, t);
            }
        });
        __syncthreads();
        // switch mapping: this thread now owns elements t2 + e*T of row slot2
        const cpx *sm2 = smem + slot2 * SE;
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = sm2[pad_pos<SH>(t2 + e * T)];
        engine_butterflies<N2, DIR, NP - 1>(v, tw, t2);
        if (a.to_peers) {
            // exchange step of a slab 2D transform: X[x] goes to the column slab of device x >> shift
            const size_t row = (a.peers.row0 + b) * a.peers.pitch;
            const unsigned cmask = (1u << a.peers.shift) - 1u;
#pragma unroll
            for (int e = 0; e < R; e++) {
                const unsigned x = (unsigned)(t2 + e * T) * n1 + blockIdx.x * TR + slot2;
                static_cast<cpx *>(a.peers.base[x >> a.peers.shift])[row + (x & cmask)] = v[e];
            }
        } else {
            cpx *__restrict__ out = static_cast<cpx *>(a.out) + b * a.out_batch + (size_t)blockIdx.x * TR + slot2;
#pragma unroll
            for (int e = 0; e < R; e++) out[(size_t)(t2 + e * T) * n1] = v[e];
        }
        __syncthreads(); // the exchange area is reused by the next transform of the batch
    }
}

} // namespace mufft
