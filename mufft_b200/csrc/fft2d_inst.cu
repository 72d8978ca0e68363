// One translation unit per image shape: compiled with -DF2D_NX=<nx> -DF2D_NY=<ny>.  Exposes
// f2d_launch_<nx>x<ny>(args, stream) (every kind, both directions of C2C) to dispatch.cu.
#include "fft2d_kernel.cuh"
#include "launch_util.cuh"
#ifndef MUFFT_F2D_TMA
#define MUFFT_F2D_TMA 1 // C2C images of the shapes the warp-specialised TMA kernel takes (1024 x 1024) go to it (fft2d_tma_kernel.cuh): 1.056 vs 1.176 ms
#endif
#if MUFFT_F2D_TMA
#include "fft2d_tma_kernel.cuh"
#endif

#if !defined(F2D_NX) || !defined(F2D_NY)
#error "compile with -DF2D_NX=<nx> -DF2D_NY=<ny>"
#endif

#define MUFFT_CAT4_(a, b, c, d) a##b##c##d
#define MUFFT_CAT4(a, b, c, d) MUFFT_CAT4_(a, b, c, d)

namespace mufft {

// lag L: the last producer ticket of image i and its first consumer ticket are L*(NP+NQ)-NP tickets
// apart; keep that above ~1.5x the tickets in flight (one per CTA).  Measured for 1024 x 1024 x 256 on 148
// SMs (lag 1 / 2 / 3 / 4 / 6: 1.395 / 1.247 / 1.196 / 1.217 / 1.382 ms): the rule's 3 is the optimum.
static unsigned default_lag(unsigned grid, unsigned per_image) { return 1 + (3 * grid / 2 + per_image - 1) / per_image; }

// Shapes the fused kernel cannot take with the configured CTA size report "unsupported" to the
// host layer (which then uses the row kernel + column kernel path) instead of failing the build.
#if MUFFT_F2D_TMA
#ifndef MUFFT_F2D_TMA_TWM
#define MUFFT_F2D_TMA_TWM (MUFFT_TW_SMEM | MUFFT_TW_DERIVE)
#endif
// cuTensorMapEncodeTiled through the runtime's driver entry point query: no link against libcuda.
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
[[maybe_unused]] static encode_tiled_fn encode_tiled()
{
    static encode_tiled_fn fn = nullptr;
    if (fn == nullptr) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult res;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &res) == cudaSuccess && res == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_fn>(ptr);
        else
            (void)cudaGetLastError();
    }
    return fn;
}
// true = launched (or failed: *err); false = this call is not for the TMA kernel (alignment, pitch), use the ticket kernel
template <int DIR, int KIND>
static bool launch_tma(const mcu_2d_args &a, cudaStream_t stream, cudaError_t *err)
{
  if constexpr (!f2d_tma_preferred(F2D_NX, F2D_NY) || !f2d_tma_kind_supported(F2D_NX, F2D_NY, KIND)) {
    (void)a; (void)stream; (void)err;
    return false;
  } else {
    constexpr int PITCH = f2d_pitch(F2D_NX, KIND), WIDTH = f2d_width(F2D_NX, KIND);
    if (a.out_image != (size_t)PITCH * F2D_NY || (a.in_image & 1) != 0 || (reinterpret_cast<uintptr_t>(a.out) & 127) != 0 ||
        (reinterpret_cast<uintptr_t>(a.in) & 15) != 0 || a.batch * (size_t)F2D_NY >= 0x7fffffffull || a.batch < 8)
        return false; // (up to 4 images the ticket kernel is the faster one, 34 vs 38 us: profiles/r02y_ab.txt)
    encode_tiled_fn encode = encode_tiled();
    if (encode == nullptr) return false;
    static launch_cfg cfg;
    auto kernel = fft2d_tma_kernel<F2D_NX, F2D_NY, DIR, MUFFT_F2D_TMA_TWM, KIND>;
    constexpr size_t smem = f2d_tma_smem_bytes(F2D_NX, F2D_NY, KIND);
    constexpr int F2D_TMA_THREADS = f2d_tma_threads(F2D_NX, F2D_NY, KIND);
    int ctas = 0;
    if ((*err = launch_prepare(cfg, kernel, smem, F2D_TMA_THREADS, &ctas)) != cudaSuccess) return true;
    f2d_tma_params p;
    p.a = a;
    const cuuint64_t dims[2] = { 2ull * WIDTH, (cuuint64_t)a.batch * F2D_NY }; // the columns that carry data
    const cuuint64_t strides[1] = { (cuuint64_t)PITCH * sizeof(float2) };
    const cuuint32_t box[2] = { 2u * f2d_tma_cols_per_tile(F2D_NY), F2D_TMA_BOX_ROWS }; // the columns of a tile x 256 rows
    const cuuint32_t estr[2] = { 1, 1 };
    if (encode(&p.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        *err = cudaErrorInvalidValue;
        return true;
    }
    constexpr unsigned per_image = f2d_tma_row_tiles(F2D_NX, F2D_NY, KIND) + f2d_tma_col_tiles(F2D_NX, F2D_NY, KIND);
    const unsigned long long tickets = (unsigned long long)a.batch * per_image;
    if (tickets + (unsigned long long)ctas >= 0xffffffffull) { *err = cudaErrorInvalidConfiguration; return true; }
    const unsigned grid = (unsigned)ctas;
#ifdef MUFFT_F2D_TMA_LAG
    if (p.a.lag == 0) p.a.lag = MUFFT_F2D_TMA_LAG;
#endif
    // two tiles per CTA are being computed, the others travel: measured optimum (profiles/r02as_ab_tma_lag.txt) 3 images for C2C
    // 1024^2 (1.045 against 1.058 ms at 4), 4 for R2C 1024^2 and 512 x 1024 (3: 5 % / 3 % slower)
    if (p.a.lag == 0) p.a.lag = default_lag(2 * grid, per_image);
    kernel<<<grid, F2D_TMA_THREADS, smem, stream>>>(p);
    *err = cudaGetLastError();
    return true;
  }
}
#endif

template <int DIR, int KIND>
static cudaError_t launch_kind(const mcu_2d_args &a, cudaStream_t stream, unsigned *lag_only)
{
  if constexpr (!f2d_supported(F2D_NX, F2D_NY)) {
    (void)a; (void)stream; (void)lag_only;
    return cudaErrorInvalidValue;
  } else {
    static launch_cfg cfg;
    auto kernel = fft2d_fused_kernel<F2D_NX, F2D_NY, DIR, KIND>;
    constexpr size_t smem = f2d_smem_bytes(F2D_NX, F2D_NY, KIND);
    constexpr unsigned per_image = f2d_tickets_per_image(F2D_NX, F2D_NY, KIND);
    int ctas = 0;
    cudaError_t err = launch_prepare(cfg, kernel, smem, MUFFT_F2D_THREADS, &ctas);
    if (err != cudaSuccess) return err;
    if (lag_only != nullptr) {
        *lag_only = default_lag((unsigned)ctas, per_image);
        return cudaSuccess;
    }
    if (a.batch == 0) return cudaSuccess;
#if MUFFT_F2D_TMA
    if constexpr (KIND == F2D_C2C || KIND == F2D_R2C) {
        cudaError_t terr = cudaSuccess;
        if (launch_tma<DIR, KIND>(a, stream, &terr)) return terr;
    }
#endif
    const unsigned long long tickets = (unsigned long long)a.batch * per_image;
    if (tickets + (unsigned long long)ctas >= 0xffffffffull) return cudaErrorInvalidConfiguration; // every CTA draws one ticket past the end
    const unsigned grid = tickets < (unsigned long long)ctas ? (unsigned)tickets : (unsigned)ctas;
    mcu_2d_args b = a;
    if (b.lag == 0) b.lag = default_lag(grid, per_image);
    if (KIND == F2D_C2R) {
        // the ring must hold more images than the consumers run behind (or simply every image of the batch)
        const size_t lag_eff = b.lag < b.batch ? b.lag : b.batch;
        if (b.mid == nullptr || b.ring == 0 || !(b.ring > lag_eff || b.ring >= b.batch)) return cudaErrorInvalidValue;
    }
    kernel<<<grid, MUFFT_F2D_THREADS, smem, stream>>>(b);
    return cudaGetLastError();
  }
}

template <typename F>
static cudaError_t dispatch_kind(const mcu_2d_args &a, F &&f)
{
    switch (a.kind) {
    case MCU_2D_C2C: return a.direction < 0 ? f(std::integral_constant<int, -1>{}, std::integral_constant<int, F2D_C2C>{})
                                            : f(std::integral_constant<int, 1>{}, std::integral_constant<int, F2D_C2C>{});
    case MCU_2D_R2C: return a.direction < 0 ? f(std::integral_constant<int, -1>{}, std::integral_constant<int, F2D_R2C>{}) : cudaErrorInvalidValue;
    case MCU_2D_R2C_FULL: return a.direction < 0 ? f(std::integral_constant<int, -1>{}, std::integral_constant<int, F2D_R2C_FULL>{}) : cudaErrorInvalidValue;
    case MCU_2D_C2R: return a.direction > 0 ? f(std::integral_constant<int, 1>{}, std::integral_constant<int, F2D_C2R>{}) : cudaErrorInvalidValue;
    default: return cudaErrorInvalidValue;
    }
}

bool MUFFT_CAT4(f2d_supported_, F2D_NX, x, F2D_NY)() { return f2d_supported(F2D_NX, F2D_NY); }
int MUFFT_CAT4(f2d_rows_variant_, F2D_NX, x, F2D_NY)(int kind) { return f2d_rows_variant(F2D_NX, kind); }

unsigned MUFFT_CAT4(f2d_lag_, F2D_NX, x, F2D_NY)(const mcu_2d_args &a)
{
    unsigned lag = 0;
    const cudaError_t err = dispatch_kind(a, [&](auto dir, auto kind) { return launch_kind<decltype(dir)::value, decltype(kind)::value>(a, nullptr, &lag); });
    if (err != cudaSuccess) (void)cudaGetLastError();
    return err == cudaSuccess ? lag : 0;
}

cudaError_t MUFFT_CAT4(f2d_launch_, F2D_NX, x, F2D_NY)(const mcu_2d_args &a, cudaStream_t stream)
{
    return dispatch_kind(a, [&](auto dir, auto kind) { return launch_kind<decltype(dir)::value, decltype(kind)::value>(a, stream, nullptr); });
}

} // namespace mufft
