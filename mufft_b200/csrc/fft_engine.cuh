// In-CTA Stockham FFT engine for sm_100a.
//
// One transform of length N is spread over T = N/R threads; every thread keeps R complex
// values in registers.  Register slot e of thread t holds element t + e*T, for the input, for
// every intermediate and for the output alike, so global loads and stores are one coalesced
// 8-byte access per slot.  A pass of radix r at butterfly span p lets each thread run R/r
// butterflies entirely in registers; between passes the values are exchanged through padded
// shared memory (scatter to the Stockham output position, gather back at t + e*T).
//
// What this replaces in the reference: the chain of per-pass kernels of kernel.c:84-453 and
// x86/kernel.h:216-667, each of which streams the whole array through memory
// (fft.c:811-850).  Here a transform is read once and written once.
//
// tools/smem_conflicts.py holds a numpy model of exactly these index maps (checked against
// numpy.fft) and the bank-conflict count that picked the padding.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mufft {

typedef float2 cpx;

__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cpx cmul(cpx a, cpx b)
{
    return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cpx cconj(cpx a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ cpx cscale(float s, cpx a) { return make_float2(a.x * s, a.y * s); }

// ---- compile-time schedule ----------------------------------------------------------------
// R (values per thread) and the radix list for a transform length.  Radix-R passes first, the
// left-over small radix last: the first pass (p == 1) needs no twiddles, so it should be the
// widest one.
__host__ __device__ constexpr int engine_regs(int n)
{
    return n >= 256 ? 16 : (n >= 64 ? 8 : (n >= 16 ? 4 : 2));
}
__host__ __device__ constexpr int engine_num_passes(int n)
{
    int r = engine_regs(n), rem = n, c = 0;
    while (rem >= r) { rem /= r; c++; }
    return c + (rem > 1 ? 1 : 0);
}
__host__ __device__ constexpr int engine_radix(int n, int pass)
{
    int r = engine_regs(n), rem = n;
    for (int j = 0; j < pass; j++) rem /= r;
    return rem >= r ? r : rem;
}
__host__ __device__ constexpr int engine_span(int n, int pass)
{
    int p = 1;
    for (int j = 0; j < pass; j++) p *= engine_radix(n, j);
    return p;
}
// Offset (in cpx) of a pass's twiddle block inside the engine table; block layout is
// [m-1][k], m = 1..r-1, k < p, value exp(dir*2*pi*i*k*m/(r*p)).  Pass 0 has no block.
__host__ __device__ constexpr int engine_tw_offset(int n, int pass)
{
    int off = 0;
    for (int j = 1; j < pass; j++) off += (engine_radix(n, j) - 1) * engine_span(n, j);
    return off;
}
__host__ __device__ constexpr int engine_tw_total(int n) { return engine_tw_offset(n, engine_num_passes(n)); }
__host__ __device__ constexpr int engine_pad_shift(int n) { return engine_regs(n) >= 16 ? 4 : 3; }
// Shared-memory footprint of one transform's exchange area, in cpx.
__host__ __device__ constexpr int engine_smem_elems(int n)
{
    return engine_num_passes(n) > 1 ? n + (n >> engine_pad_shift(n)) : 0;
}

// ---- small DFTs in registers -----------------------------------------------------------------
// v * exp(DIR*2*pi*i*J/16) for a compile-time J in [0, 8).
template <int DIR, int J>
__device__ __forceinline__ cpx mul_w16(cpx v)
{
    constexpr float H = 0.70710678118654752440f;
    constexpr float C1 = 0.92387953251128675613f; // cos(pi/8)
    constexpr float S1 = 0.38268343236508977173f; // sin(pi/8)
    constexpr float D = (float)DIR;
    if constexpr (J == 0) return v;
    else if constexpr (J == 4) return make_float2(-D * v.y, D * v.x);
    else if constexpr (J == 2) return make_float2(H * (v.x - D * v.y), H * (v.y + D * v.x));
    else if constexpr (J == 6) return make_float2(-H * (v.x + D * v.y), H * (D * v.x - v.y));
    else {
        constexpr float c = (J == 1) ? C1 : (J == 3) ? S1 : (J == 5) ? -S1 : -C1;
        constexpr float s = D * ((J == 1 || J == 7) ? S1 : C1);
        return make_float2(fmaf(v.x, c, -(v.y * s)), fmaf(v.x, s, v.y * c));
    }
}

// One radix-2 Stockham stage inside the register file: span P, transform length R.
template <int R, int P, int DIR, int I>
struct RegStage {
    static __device__ __forceinline__ void run(cpx (&out)[R], const cpx (&in)[R])
    {
        constexpr int k = I & (P - 1);
        constexpr int J = k * (8 / P); // exp(DIR*2*pi*i*k/(2P)) on the 16th-root grid (P <= 8)
        const cpx a = in[I];
        const cpx b = mul_w16<DIR, J>(in[I + R / 2]);
        out[(I - k) * 2 + k] = cadd(a, b);
        out[(I - k) * 2 + k + P] = csub(a, b);
        if constexpr (I + 1 < R / 2) RegStage<R, P, DIR, I + 1>::run(out, in);
    }
};

// Natural-order in, natural-order out DFT of R in {2,4,8,16} register values.
template <int R, int DIR>
__device__ __forceinline__ void dft_regs(cpx (&v)[R])
{
    cpx w[R];
    if constexpr (R == 2) {
        RegStage<2, 1, DIR, 0>::run(w, v);
#pragma unroll
        for (int i = 0; i < 2; i++) v[i] = w[i];
    } else if constexpr (R == 4) {
        RegStage<4, 1, DIR, 0>::run(w, v);
        RegStage<4, 2, DIR, 0>::run(v, w);
    } else if constexpr (R == 8) {
        RegStage<8, 1, DIR, 0>::run(w, v);
        RegStage<8, 2, DIR, 0>::run(v, w);
        RegStage<8, 4, DIR, 0>::run(w, v);
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = w[i];
    } else {
        static_assert(R == 16, "register DFT sizes are 2, 4, 8, 16");
        RegStage<16, 1, DIR, 0>::run(w, v);
        RegStage<16, 2, DIR, 0>::run(v, w);
        RegStage<16, 4, DIR, 0>::run(w, v);
        RegStage<16, 8, DIR, 0>::run(v, w);
    }
}

// ---- one engine pass -----------------------------------------------------------------------------
// Padded exchange position.
template <int SHIFT>
__device__ __forceinline__ int pad_pos(int pos) { return pos + (pos >> SHIFT); }

// Butterflies of pass PASS for thread t, in place: inputs v[u + m*NB] (m < r) become outputs
// v[u + q*NB] (q < r).  `tw` points at the engine twiddle table of this transform length.
template <int N, int DIR, int PASS>
__device__ __forceinline__ void engine_butterflies(cpx (&v)[engine_regs(N)], const cpx *__restrict__ tw, int t)
{
    constexpr int R = engine_regs(N);
    constexpr int T = N / R;
    constexpr int r = engine_radix(N, PASS);
    constexpr int p = engine_span(N, PASS);
    constexpr int NB = R / r;
#pragma unroll
    for (int u = 0; u < NB; u++) {
        cpx x[r];
#pragma unroll
        for (int m = 0; m < r; m++) x[m] = v[u + m * NB];
        if constexpr (PASS > 0) {
            const int k = (t + u * T) & (p - 1);
            const cpx *w = tw + engine_tw_offset(N, PASS) + k;
#pragma unroll
            for (int m = 1; m < r; m++) x[m] = cmul(x[m], __ldg(w + (m - 1) * p));
        }
        dft_regs<r, DIR>(x);
#pragma unroll
        for (int q = 0; q < r; q++) v[u + q * NB] = x[q];
    }
}

// Scatter the outputs of pass PASS to shared memory.  `sm` is this transform's exchange area;
// CS is the element stride between consecutive positions (1 for row transforms, the number of
// columns in the tile for column transforms).
template <int N, int PASS, int CS>
__device__ __forceinline__ void engine_scatter(const cpx (&v)[engine_regs(N)], cpx *sm, int t)
{
    constexpr int R = engine_regs(N);
    constexpr int T = N / R;
    constexpr int r = engine_radix(N, PASS);
    constexpr int p = engine_span(N, PASS);
    constexpr int NB = R / r;
    constexpr int SH = engine_pad_shift(N);
#pragma unroll
    for (int u = 0; u < NB; u++) {
        const int i = t + u * T;
        const int k = i & (p - 1);
        const int base = (i - k) * r + k;
#pragma unroll
        for (int q = 0; q < r; q++) sm[pad_pos<SH>(base + q * p) * CS] = v[u + q * NB];
    }
}

// Gather element t + e*T into slot e.
template <int N, int CS>
__device__ __forceinline__ void engine_gather(cpx (&v)[engine_regs(N)], const cpx *sm, int t)
{
    constexpr int R = engine_regs(N);
    constexpr int T = N / R;
    constexpr int SH = engine_pad_shift(N);
#pragma unroll
    for (int e = 0; e < R; e++) v[e] = sm[pad_pos<SH>(t + e * T) * CS];
}

// Natural-order scatter (slot e -> element t + e*T); used by the R2C epilogue, which needs
// Z[k] and Z[N-k] from different threads.
template <int N, int CS>
__device__ __forceinline__ void engine_scatter_natural(const cpx (&v)[engine_regs(N)], cpx *sm, int t)
{
    constexpr int R = engine_regs(N);
    constexpr int T = N / R;
    constexpr int SH = engine_pad_shift(N);
#pragma unroll
    for (int e = 0; e < R; e++) sm[pad_pos<SH>(t + e * T) * CS] = v[e];
}

// All passes of an N-point transform.  Every thread of the CTA must call this (it contains
// __syncthreads()); `sm` may be anything when the transform has a single pass.
template <int N, int DIR, int CS, int PASS = 0>
__device__ __forceinline__ void engine_run(cpx (&v)[engine_regs(N)], cpx *sm, const cpx *__restrict__ tw, int t)
{
    engine_butterflies<N, DIR, PASS>(v, tw, t);
    if constexpr (PASS + 1 < engine_num_passes(N)) {
        if constexpr (PASS > 0) __syncthreads(); // previous gather must be complete (WAR)
        engine_scatter<N, PASS, CS>(v, sm, t);
        __syncthreads();
        engine_gather<N, CS>(v, sm, t);
        engine_run<N, DIR, CS, PASS + 1>(v, sm, tw, t);
    }
}

} // namespace mufft
