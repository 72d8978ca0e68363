// In-CTA Stockham FFT engine for sm_100a.
//
// One transform of length N is spread over T = N/R threads; every thread keeps R complex
// values in registers.  Register slot e of thread t holds element t + e*T, for the input, for
// every intermediate and for the output alike, so global loads and stores are one coalesced
// 8-byte access per slot.  A pass of radix r at butterfly span p lets each thread run R/r
// butterflies entirely in registers; between passes the values are exchanged through padded
// shared memory (scatter to the Stockham output position, gather back at t + e*T).
//
// The shape table below picks R and the radix list per length.  The load/store pipe (LSU) is
// the busiest unit of this kernel family on B200 (ncu: profiles/r01_rows4096_c2c_v0.txt), and
// every exchange costs 16 bytes of shared-memory traffic per element, so shapes minimise the
// number of passes: up to 64 values per thread and radix-64 butterflies (4096 = 64 x 64 is two
// passes, one exchange).
//
// What this replaces in the reference: the chain of per-pass kernels of kernel.c:84-453 and
// x86/kernel.h:216-667, each of which streams the whole array through memory
// (fft.c:811-850).  Here a transform is read once and written once.
//
// tools/smem_conflicts.py holds a numpy model of exactly these index maps (checked against
// numpy.fft) and the bank-conflict count that picked the padding.
#pragma once

#ifdef MUFFT_EMU
#include "cuda_emu.h" // tests/emu: CPU emulation of the execution model, test builds only
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <type_traits>
#include <utility>

// Shared-memory declarations, spelled through macros so that tests/emu can give every emulated
// CTA its own storage; in the CUDA build they are the plain declarations.
#ifdef MUFFT_EMU
#define MUFFT_DYNAMIC_SMEM(name) unsigned char *name = emu::dyn_smem()
#define MUFFT_STATIC_SMEM_U32(name, count) unsigned *name = emu::static_smem<unsigned>(count)
#else
#define MUFFT_DYNAMIC_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#define MUFFT_STATIC_SMEM_U32(name, count) __shared__ unsigned name[count]
#endif

namespace mufft {

typedef float2 cpx;

__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return make_float2(a.x - b.x, a.y - b.y); }
// Blackwell packed FP32 (add.f32x2 -> one FADD2 per complex add instead of two FADD): a complex
// value is a natural (re, im) register pair.  PK = false keeps the scalar form; which is faster
// depends on the register pressure of the kernel (pairs must sit in aligned registers), so it
// is a per-shape choice (engine_shape_t::packed, measured in profiles/).
template <bool PK>
__device__ __forceinline__ cpx caddp(cpx a, cpx b)
{
    if constexpr (PK) return __fadd2_rn(a, b);
    else return cadd(a, b);
}
template <bool PK>
__device__ __forceinline__ cpx csubp(cpx a, cpx b)
{
    if constexpr (PK) return __fadd2_rn(a, make_float2(-b.x, -b.y));
    else return csub(a, b);
}
__device__ __forceinline__ cpx cmul(cpx a, cpx b)
{
    return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ cpx cconj(cpx a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ cpx cscale(float s, cpx a) { return make_float2(a.x * s, a.y * s); }

// Compile-time loop: f(std::integral_constant<int, I>) for I in [0, N).
template <int N, typename F, int... I>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>)
{
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, typename F>
__device__ __forceinline__ void static_for(F &&f)
{
    static_for_impl<N>(static_cast<F &&>(f), std::make_integer_sequence<int, N>{});
}

// ---- compile-time schedule ----------------------------------------------------------------
#define MUFFT_ENGINE_MAX_PASSES 4
#ifndef MUFFT_ROWS_CTA_THREADS
#define MUFFT_ROWS_CTA_THREADS 256
#endif
#ifndef MUFFT_ROWS_MIN_BLOCKS
#define MUFFT_ROWS_MIN_BLOCKS 1
#endif
#ifndef MUFFT_SHAPE_R1
#define MUFFT_SHAPE_R1 1
#endif
#ifndef MUFFT_SHAPE_R2
#define MUFFT_SHAPE_R2 1
#endif
#ifndef MUFFT_SHAPE_R3
#define MUFFT_SHAPE_R3 1
#endif
struct engine_shape_t {
    int regs;   // R: complex values per thread
    int passes;
    int radix[MUFFT_ENGINE_MAX_PASSES];
    int cta_threads; // row kernels: threads per CTA (several transforms share a CTA when N/R is smaller)
    int min_blocks;  // row kernels: __launch_bounds__ minimum CTAs per SM
    int pad_shift;   // one padding element per 2^pad_shift elements of the exchange area (0 = derive from radix[0])
    int packed;      // 1 = complex adds as packed FP32x2 instructions (FADD2), see caddp
    int tw;          // twiddle mode of the kernels that do not choose one themselves: 0 = every twiddle from the
                     // table, MUFFT_TW_DERIVE = power-of-two entries loaded, the rest derived (see engine_butterflies)
};

// Shape variants.  Variant 0 is the general one.  Variants 1 and 2 exist for the real transforms:
// in ONE pass thread t owns butterfly i together with its mirror image q0 - i (q0 = N/r butterflies
// per pass), so that the bins k and N-k of the half-length spectrum sit in the same thread:
//   variant 1 (R2C): the LAST pass is mirrored; the resolve X[k] <- Z[k], Z[N-k] (ref kernel.c:47-82)
//                    happens in registers on the way out, with no extra trip through shared memory;
//   variant 2 (C2R): the FIRST pass is mirrored; X[k] and X[N-k] are loaded once, by the thread that
//                    needs both (ref kernel.c:36-45).
// The mirrored pass needs an even number (>= 2) of butterflies per thread, R / radix >= 2.
// tools/mirror_model.py is the numpy model of these index maps and of their bank conflicts.
#define MUFFT_VARIANT_GENERAL 0
#define MUFFT_VARIANT_MIRROR_LAST 1
#define MUFFT_VARIANT_MIRROR_FIRST 2
#ifndef MUFFT_SHAPE_V
#define MUFFT_SHAPE_V 0
#endif
#ifndef MUFFT_SHAPE_PAD
#define MUFFT_SHAPE_PAD 0
#endif
#ifndef MUFFT_SHAPE_PACKED
#define MUFFT_SHAPE_PACKED 1
#endif
#ifndef MUFFT_SHAPE_TW
#define MUFFT_SHAPE_TW 0
#endif
// -DMUFFT_TW_ALL=<mode> forces the twiddle mode of every shape (A/B measurements)
// -DMUFFT_PACKED=<0|1> forces the packed flag of every shape (A/B measurements)

// Widest radix first in variant 0: the first pass (p == 1) needs no twiddles.  Every radix divides R.
// cta_threads / min_blocks are __launch_bounds__ of the row kernels: min_blocks is really a
// register cap (65536 / (cta_threads * min_blocks) per thread) and decides how many CTAs share
// an SM; the entries are the best of the sweeps in profiles/.
// -DMUFFT_SHAPE_N=<n> [-DMUFFT_SHAPE_V=<variant>] -DMUFFT_SHAPE_R=<R> -DMUFFT_SHAPE_R0=<r0> [-DMUFFT_SHAPE_R1=<r1> ...]
// overrides one entry (used by tools/ab.sh to measure alternatives on the GPU).
__host__ __device__ constexpr engine_shape_t engine_shape(int n, int variant = 0)
{
#ifdef MUFFT_SHAPE_N
    if (n == MUFFT_SHAPE_N && variant == MUFFT_SHAPE_V) {
        constexpr int r[] = { MUFFT_SHAPE_R0, MUFFT_SHAPE_R1, MUFFT_SHAPE_R2, MUFFT_SHAPE_R3 };
        int passes = 0, prod = 1;
        while (prod < MUFFT_SHAPE_N) prod *= r[passes++];
        return { MUFFT_SHAPE_R, passes, { r[0], r[1], r[2], r[3] }, MUFFT_ROWS_CTA_THREADS, MUFFT_ROWS_MIN_BLOCKS, MUFFT_SHAPE_PAD, MUFFT_SHAPE_PACKED, MUFFT_SHAPE_TW };
    }
#endif
    if (variant == MUFFT_VARIANT_MIRROR_LAST) {
        switch (n) {
        case 512: return { 16, 3, { 16, 16, 2, 1 }, 256, 1, 4, 1 };
        case 2048: return { 32, 3, { 32, 32, 2, 1 }, 64, 1, 5, 1 };
        case 4096: return { 32, 3, { 16, 16, 16, 1 }, 128, 4, 4, 1 };
        case 8192: return { 32, 3, { 32, 16, 16, 1 }, 256, 2, 5, 0 };
        case 16384: return { 32, 3, { 32, 32, 16, 1 }, 512, 1, 5, 0 };
        default: return { 0, 0, { 1, 1, 1, 1 }, 256, 1, 0, 1 };
        }
    }
    if (variant == MUFFT_VARIANT_MIRROR_FIRST) {
        switch (n) {
        case 512: return { 16, 3, { 2, 16, 16, 1 }, 256, 1, 4, 1 };
        case 2048: return { 16, 3, { 8, 16, 16, 1 }, 128, 6, 4, 1 };
        case 4096: return { 32, 3, { 16, 16, 16, 1 }, 128, 4, 4, 1 };
        case 8192: return { 32, 3, { 16, 16, 32, 1 }, 256, 2, 4, 1 };
        case 16384: return { 32, 3, { 16, 32, 32, 1 }, 512, 1, 4, 1 };
        default: return { 0, 0, { 1, 1, 1, 1 }, 256, 1, 0, 1 };
        }
    }
    switch (n) {
    case 2: return { 2, 1, { 2, 1, 1, 1 } , 256, 1, 0, 1 };
    case 4: return { 4, 1, { 4, 1, 1, 1 } , 256, 1, 0, 1 };
    case 8: return { 4, 2, { 4, 2, 1, 1 } , 256, 1, 0, 1 };
    case 16: return { 4, 2, { 4, 4, 1, 1 } , 256, 1, 0, 1 };
    case 32: return { 4, 3, { 4, 4, 2, 1 } , 256, 1, 0, 1 };
    case 64: return { 8, 2, { 8, 8, 1, 1 } , 256, 1, 0, 1 };
    case 128: return { 8, 3, { 8, 8, 2, 1 } , 256, 1, 0, 1 };
    case 256: return { 16, 2, { 16, 16, 1, 1 } , 256, 1, 0, 1 };
    case 512: return { 16, 3, { 16, 16, 2, 1 } , 256, 1, 0, 1 };
    case 1024: return { 32, 2, { 32, 32, 1, 1 }, 128, 1, 0, 1 }; // packed: 1D 1.03 -> 1.05, fused 2D 1.210 -> 1.191 ms (profiles/r02p_ab.txt)
    case 2048: return { 32, 3, { 32, 32, 2, 1 }, 64, 1, 0, 1 };
    case 4096: return { 16, 3, { 16, 16, 16, 1 }, 256, 3, 0, 1 };
    case 8192: return { 32, 3, { 32, 16, 16, 1 }, 256, 2, 0, 0 };
    // One CTA of 512 threads per transform, the whole register file of an SM (128 registers per
    // thread) and 132 KiB of exchange area: one CTA per SM (measured 1.8x faster than the two-kernel split).
    case 16384: return { 32, 3, { 32, 32, 16, 1 }, 512, 1, 0, 0 };
    default: return { 0, 0, { 1, 1, 1, 1 }, 256, 1, 0, 1 };
    }
}
__host__ __device__ constexpr int engine_regs(int n, int v = 0) { return engine_shape(n, v).regs; }
__host__ __device__ constexpr bool engine_packed(int n, int v = 0)
{
#ifdef MUFFT_PACKED
    return MUFFT_PACKED != 0;
#else
    return engine_shape(n, v).packed != 0;
#endif
}
__host__ __device__ constexpr int engine_tw_mode(int n, int v = 0)
{
#ifdef MUFFT_TW_ALL
    return MUFFT_TW_ALL;
#else
    return engine_shape(n, v).tw;
#endif
}
__host__ __device__ constexpr int engine_num_passes(int n, int v = 0) { return engine_shape(n, v).passes; }
__host__ __device__ constexpr int engine_radix(int n, int pass, int v = 0) { return engine_shape(n, v).radix[pass]; }
__host__ __device__ constexpr int engine_threads(int n, int v = 0) { return n / engine_regs(n, v); }
__host__ __device__ constexpr int engine_cta_threads(int n, int v = 0) { return engine_shape(n, v).cta_threads; }
__host__ __device__ constexpr int engine_min_blocks(int n, int v = 0) { return engine_shape(n, v).min_blocks; }
__host__ __device__ constexpr int engine_span(int n, int pass, int v = 0)
{
    int p = 1;
    for (int j = 0; j < pass; j++) p *= engine_radix(n, j, v);
    return p;
}
// Is pass `pass` of variant v the mirrored one?
__host__ __device__ constexpr bool engine_pass_mirrored(int n, int pass, int v)
{
    return (v == MUFFT_VARIANT_MIRROR_LAST && pass == engine_num_passes(n, v) - 1) ||
           (v == MUFFT_VARIANT_MIRROR_FIRST && pass == 0);
}
// Does variant v exist for length n?
__host__ __device__ constexpr bool engine_variant_exists(int n, int v)
{
    if (engine_regs(n, v) == 0) return false;
    if (v == MUFFT_VARIANT_GENERAL) return true;
    const int pass = v == MUFFT_VARIANT_MIRROR_LAST ? engine_num_passes(n, v) - 1 : 0;
    const int nb = engine_regs(n, v) / engine_radix(n, pass, v);
    return engine_num_passes(n, v) >= 2 && nb >= 2 && nb % 2 == 0;
}
// Offset (in cpx) of a pass's twiddle block inside the engine table; block layout is
// [m-1][k], m = 1..r-1, k < p, value exp(dir*2*pi*i*k*m/(r*p)).  Pass 0 has no block.
__host__ __device__ constexpr int engine_tw_offset(int n, int pass, int v = 0)
{
    int off = 0;
    for (int j = 1; j < pass; j++) off += (engine_radix(n, j, v) - 1) * engine_span(n, j, v);
    return off;
}
__host__ __device__ constexpr int engine_tw_total(int n, int v = 0) { return engine_tw_offset(n, engine_num_passes(n, v), v); }
// One padding element after every 2^shift elements.  The first scatter writes with an element
// stride of radix[0] across lanes, so the default pad period is radix[0] (at least 8): lane l then
// hits bank pair (radix[0] + 1) * l, distinct for 16 consecutive lanes.
__host__ __device__ constexpr int engine_pad_shift(int n, int v = 0)
{
    if (engine_shape(n, v).pad_shift != 0) return engine_shape(n, v).pad_shift;
    int r = engine_radix(n, 0, v), s = 0;
    while ((1 << s) < r) s++;
    return s < 3 ? 3 : s;
}
// Shared-memory footprint of one transform's exchange area, in cpx.
__host__ __device__ constexpr int engine_smem_elems(int n, int v = 0) { return n + (n >> engine_pad_shift(n, v)) + 1; }

// ---- small DFTs in registers -----------------------------------------------------------------
// cos(2*pi*j/64), j = 0..16
constexpr float kCos64[17] = {
    1.0f, 0.99518472667219693f, 0.98078528040323043f, 0.95694033573220882f, 0.92387953251128674f,
    0.88192126434835505f, 0.83146961230254524f, 0.77301045336273699f, 0.70710678118654757f,
    0.63439328416364549f, 0.55557023301960229f, 0.47139673682599781f, 0.38268343236508984f,
    0.29028467725446233f, 0.19509032201612833f, 0.09801714032956077f, 0.0f };
__host__ __device__ constexpr float cos64(int j)
{
    j &= 63;
    return j <= 16 ? kCos64[j] : j <= 32 ? -kCos64[32 - j] : j <= 48 ? -kCos64[j - 32] : kCos64[64 - j];
}
__host__ __device__ constexpr float sin64(int j) { return cos64(j - 16); }

// v * exp(DIR*2*pi*i*J/64) for a compile-time J.
template <int DIR, int J64>
__device__ __forceinline__ cpx mul_w64(cpx v)
{
    constexpr int J = ((J64 % 64) + 64) % 64;
    constexpr float D = (float)DIR;
    constexpr float H = 0.70710678118654752440f;
    if constexpr (J == 0) return v;
    else if constexpr (J == 16) return make_float2(-D * v.y, D * v.x);
    else if constexpr (J == 32) return make_float2(-v.x, -v.y);
    else if constexpr (J == 48) return make_float2(D * v.y, -D * v.x);
    else if constexpr (J == 8) return make_float2(H * (v.x - D * v.y), H * (v.y + D * v.x));
    else if constexpr (J == 24) return make_float2(-H * (v.x + D * v.y), H * (D * v.x - v.y));
    else if constexpr (J == 40) return make_float2(-H * (v.x - D * v.y), -H * (v.y + D * v.x));
    else if constexpr (J == 56) return make_float2(H * (v.x + D * v.y), -H * (D * v.x - v.y));
    else {
        constexpr float c = cos64(J);
        constexpr float s = D * sin64(J);
        return make_float2(fmaf(v.x, c, -(v.y * s)), fmaf(v.x, s, v.y * c));
    }
}

// Natural-order in, natural-order out DFT of R in {2,4,8,16,32,64} register values:
// X[k] = sum_n x[n] exp(DIR*2*pi*i*n*k/R).  Composite sizes use the Cooley-Tukey split R = A*B
// with n = n1 + A*n2, k = B*k1 + k2: B-point DFTs over the stride-A subsequences, constant
// twiddles exp(DIR*2*pi*i*n1*k2/R), then A-point DFTs across n1.
template <int R, int DIR, bool PK>
struct RegDft;

template <int DIR, bool PK>
struct RegDft<2, DIR, PK> {
    static __device__ __forceinline__ void run(cpx (&x)[2])
    {
        const cpx a = x[0], b = x[1];
        x[0] = caddp<PK>(a, b);
        x[1] = csubp<PK>(a, b);
    }
};

template <int DIR, bool PK>
struct RegDft<4, DIR, PK> {
    static __device__ __forceinline__ void run(cpx (&x)[4])
    {
        const cpx t0 = caddp<PK>(x[0], x[2]), t1 = csubp<PK>(x[0], x[2]);
        const cpx t2 = caddp<PK>(x[1], x[3]), t3 = mul_w64<DIR, 16>(csubp<PK>(x[1], x[3]));
        x[0] = caddp<PK>(t0, t2);
        x[1] = caddp<PK>(t1, t3);
        x[2] = csubp<PK>(t0, t2);
        x[3] = csubp<PK>(t1, t3);
    }
};

__host__ __device__ constexpr int dft_split_a(int r) { return r == 8 ? 2 : r == 16 ? 4 : r == 32 ? 4 : 8; }

template <int R, int DIR, bool PK>
struct RegDft {
    static __device__ __forceinline__ void run(cpx (&x)[R])
    {
        constexpr int A = dft_split_a(R), B = R / A;
        static_assert(A * B == R && R <= 64, "register DFT sizes are 2..64");
        static_for<A>([&](auto n1c) {
            constexpr int n1 = decltype(n1c)::value;
            cpx s[B];
            static_for<B>([&](auto n2c) { constexpr int n2 = decltype(n2c)::value; s[n2] = x[n1 + A * n2]; });
            RegDft<B, DIR, PK>::run(s);
            static_for<B>([&](auto k2c) {
                constexpr int k2 = decltype(k2c)::value;
                x[n1 + A * k2] = mul_w64<DIR, n1 * k2 * (64 / R)>(s[k2]);
            });
        });
        static_for<B>([&](auto k2c) {
            constexpr int k2 = decltype(k2c)::value;
            cpx s[A];
            static_for<A>([&](auto n1c) { constexpr int n1 = decltype(n1c)::value; s[n1] = x[n1 + A * k2]; });
            RegDft<A, DIR, PK>::run(s);
            static_for<A>([&](auto k1c) { constexpr int k1 = decltype(k1c)::value; x[n1_slot<A, B>(k1, k2)] = s[k1]; });
        });
        // The second stage left X[B*k1 + k2] in slot k1 + A*k2; put it in natural order.
        if constexpr (A != 1 && B != 1) {
            cpx y[R];
            static_for<R>([&](auto ic) { constexpr int i = decltype(ic)::value; y[i] = x[i]; });
            static_for<A>([&](auto k1c) {
                constexpr int k1 = decltype(k1c)::value;
                static_for<B>([&](auto k2c) { constexpr int k2 = decltype(k2c)::value; x[B * k1 + k2] = y[k1 + A * k2]; });
            });
        }
    }
    template <int A_, int B_>
    static __host__ __device__ constexpr int n1_slot(int k1, int k2) { return k1 + A_ * k2; }
};

template <int R, int DIR, bool PK = false>
__device__ __forceinline__ void dft_regs(cpx (&v)[R]) { RegDft<R, DIR, PK>::run(v); }

// ---- synchronisation of the threads of one transform ---------------------------------------
// T threads share a transform.  T <= 32: they sit in one warp.  Otherwise a named barrier
// (id 1 + slot) when the CTA holds several transforms, or the CTA barrier when it holds one.
template <int T, int SLOTS>
__device__ __forceinline__ void engine_sync(int slot)
{
    if constexpr (T <= 32) {
        __syncwarp();
    } else if constexpr (SLOTS == 1) {
        __syncthreads();
    } else {
#ifdef MUFFT_EMU
        emu::bar_sync(slot + 1, T);
#else
        asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(T) : "memory");
#endif
    }
}

// ---- one engine pass -----------------------------------------------------------------------------
// Padded exchange position.
template <int SHIFT>
__device__ __forceinline__ int pad_pos(int pos) { return pos + (pos >> SHIFT); }

// Butterfly of pass PASS that thread t runs in register group u (u < NB = R / r).  In general
// t + u*T.  In a mirrored pass groups u < NB/2 hold i0 = t + u*T and group u + NB/2 holds its
// mirror image q0 - i0 (q0 = N/r butterflies per pass); butterflies 0 and q0/2 are their own
// images and share thread 0.
template <int N, int PASS, int V>
__device__ __forceinline__ int engine_bfly(int t, int u)
{
    constexpr int R = engine_regs(N, V);
    constexpr int T = N / R;
    if constexpr (engine_pass_mirrored(N, PASS, V)) {
        constexpr int r = engine_radix(N, PASS, V);
        constexpr int H = R / r / 2;
        constexpr int Q0 = N / r;
        if (u < H) return t + u * T;
        const int i0 = t + (u - H) * T;
        return i0 == 0 ? Q0 / 2 : Q0 - i0;
    } else {
        return t + u * T;
    }
}

// Butterflies of pass PASS for thread t, in place: inputs v[u + m*NB] (m < r) become outputs
// v[u + q*NB] (q < r).  `tw` points at the engine twiddle table of this transform length.
// PK: -1 = the shape's packed flag, 0 / 1 = the caller's choice (kernels whose prologue or epilogue
// changes the register pressure measure differently from the plain transform of the same shape).
// Twiddle handling of a pass (TWM, a bit mask):
//   MUFFT_TW_SMEM    `tw` points into SHARED memory (plain loads) instead of global memory (read-only
//                    cache loads);
//   MUFFT_TW_DERIVE  radix >= 8: only the log2(r) power-of-two entries w^(k 2^j) are loaded, the other
//                    r-1-log2(r) twiddles are products of them (at most log2(r)-1 multiplications deep).
//                    26 loads less and 26 complex multiplications more per radix-32 butterfly: for
//                    kernels bound by the load/store pipe, not for those bound by instruction issue.
#define MUFFT_TW_SMEM 1
#define MUFFT_TW_DERIVE 2
__host__ __device__ constexpr int ilog2(int r) { int s = 0; while ((1 << s) < r) s++; return s; }

// x[IDX + j] *= cur * (product of b[l] over the set bits l <= LEVEL of j), for every j below 2^(LEVEL+1):
// a depth-first walk over the bits of the twiddle exponent, one live product per level.
template <int R, int LG, int LEVEL, int IDX, bool ONE>
struct TwTree {
    static __device__ __forceinline__ void run(cpx (&x)[R], const cpx (&b)[LG], cpx cur)
    {
        if constexpr (LEVEL < 0) {
            if constexpr (!ONE) x[IDX] = cmul(x[IDX], cur);
        } else {
            TwTree<R, LG, LEVEL - 1, IDX, ONE>::run(x, b, cur);
            if constexpr (ONE) TwTree<R, LG, LEVEL - 1, IDX + (1 << LEVEL), false>::run(x, b, b[LEVEL]);
            else TwTree<R, LG, LEVEL - 1, IDX + (1 << LEVEL), false>::run(x, b, cmul(cur, b[LEVEL]));
        }
    }
};

template <int N, int DIR, int PASS, int V = 0, int PK = -1, int TWM_ = -1>
__device__ __forceinline__ void engine_butterflies(cpx (&v)[engine_regs(N, V)], const cpx *__restrict__ tw, int t)
{
    constexpr int R = engine_regs(N, V);
    constexpr int r = engine_radix(N, PASS, V);
    constexpr int p = engine_span(N, PASS, V);
    constexpr int NB = R / r;
    constexpr int TWM = TWM_ < 0 ? engine_tw_mode(N, V) : TWM_; // -1 = the shape's own choice
    static_assert(NB * r == R, "every pass radix must divide the values per thread");
#pragma unroll
    for (int u = 0; u < NB; u++) {
        cpx x[r];
#pragma unroll
        for (int m = 0; m < r; m++) x[m] = v[u + m * NB];
        if constexpr (PASS > 0) {
            const int k = engine_bfly<N, PASS, V>(t, u) & (p - 1);
            const cpx *w = tw + engine_tw_offset(N, PASS, V) + k;
            if constexpr ((TWM & MUFFT_TW_DERIVE) != 0 && r >= 8) {
                constexpr int LG = ilog2(r);
                cpx b[LG];
#pragma unroll
                for (int j = 0; j < LG; j++) {
                    if constexpr ((TWM & MUFFT_TW_SMEM) != 0) b[j] = w[((1 << j) - 1) * p];
                    else b[j] = __ldg(w + ((1 << j) - 1) * p);
                }
                TwTree<r, LG, LG - 1, 0, true>::run(x, b, make_float2(1.0f, 0.0f));
            } else {
#pragma unroll
                for (int m = 1; m < r; m++) {
                    if constexpr ((TWM & MUFFT_TW_SMEM) != 0) x[m] = cmul(x[m], w[(m - 1) * p]);
                    else x[m] = cmul(x[m], __ldg(w + (m - 1) * p));
                }
            }
        }
        dft_regs<r, DIR, (PK < 0 ? engine_packed(N, V) : PK != 0)>(x);
#pragma unroll
        for (int q = 0; q < r; q++) v[u + q * NB] = x[q];
    }
}

// Scatter the outputs of pass PASS to shared memory.  `sm` is this transform's exchange area;
// CS is the element stride between consecutive positions (1 for row transforms, the number of
// columns in the tile for column transforms).
template <int N, int PASS, int CS, int V = 0>
__device__ __forceinline__ void engine_scatter(const cpx (&v)[engine_regs(N, V)], cpx *sm, int t)
{
    constexpr int R = engine_regs(N, V);
    constexpr int r = engine_radix(N, PASS, V);
    constexpr int p = engine_span(N, PASS, V);
    constexpr int NB = R / r;
    constexpr int SH = engine_pad_shift(N, V);
#pragma unroll
    for (int u = 0; u < NB; u++) {
        const int i = engine_bfly<N, PASS, V>(t, u);
        const int k = i & (p - 1);
        const int base = (i - k) * r + k;
#pragma unroll
        for (int q = 0; q < r; q++) sm[pad_pos<SH>(base + q * p) * CS] = v[u + q * NB];
    }
}

// Gather the inputs of pass PASS: element i + m*(N/r) of butterfly i into slot u + m*NB.  For an
// ordinary pass that is element t + e*T into slot e.
template <int N, int CS, int PASS = 1, int V = 0>
__device__ __forceinline__ void engine_gather(cpx (&v)[engine_regs(N, V)], const cpx *sm, int t)
{
    constexpr int R = engine_regs(N, V);
    constexpr int T = N / R;
    constexpr int SH = engine_pad_shift(N, V);
    if constexpr (engine_pass_mirrored(N, PASS, V)) {
        constexpr int r = engine_radix(N, PASS, V);
        constexpr int NB = R / r;
        constexpr int Q0 = N / r;
#pragma unroll
        for (int u = 0; u < NB; u++) {
            const int i = engine_bfly<N, PASS, V>(t, u);
#pragma unroll
            for (int m = 0; m < r; m++) v[u + m * NB] = sm[pad_pos<SH>(i + m * Q0) * CS];
        }
    } else {
#pragma unroll
        for (int e = 0; e < R; e++) v[e] = sm[pad_pos<SH>(t + e * T) * CS];
    }
}

// Natural-order scatter (slot e -> element t + e*T); used by the R2C epilogue of variant 0,
// which needs Z[k] and Z[N-k] from different threads.
template <int N, int CS, int V = 0>
__device__ __forceinline__ void engine_scatter_natural(const cpx (&v)[engine_regs(N, V)], cpx *sm, int t)
{
    constexpr int R = engine_regs(N, V);
    constexpr int T = N / R;
    constexpr int SH = engine_pad_shift(N, V);
#pragma unroll
    for (int e = 0; e < R; e++) sm[pad_pos<SH>(t + e * T) * CS] = v[e];
}

// All passes of an N-point transform.  Every thread of the transform must call this (it
// synchronises them); SYNC_T / SLOTS say how (see engine_sync): the row kernel passes the
// threads of one transform, the column kernel passes the whole CTA.
template <int N, int DIR, int CS, int SYNC_T, int SLOTS, int PASS = 0, int V = 0, int PK = -1, int TWM = -1>
__device__ __forceinline__ void engine_run(cpx (&v)[engine_regs(N, V)], cpx *sm, const cpx *__restrict__ tw, int t, int slot)
{
    engine_butterflies<N, DIR, PASS, V, PK, TWM>(v, tw, t);
    if constexpr (PASS + 1 < engine_num_passes(N, V)) {
        if constexpr (PASS > 0) engine_sync<SYNC_T, SLOTS>(slot); // previous gather must be complete (WAR)
        engine_scatter<N, PASS, CS, V>(v, sm, t);
        engine_sync<SYNC_T, SLOTS>(slot);
        engine_gather<N, CS, PASS + 1, V>(v, sm, t);
        engine_run<N, DIR, CS, SYNC_T, SLOTS, PASS + 1, V, PK, TWM>(v, sm, tw, t, slot);
    }
}

} // namespace mufft
