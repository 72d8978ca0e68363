// One translation unit per image shape: compiled with -DF2D_NX=<nx> -DF2D_NY=<ny>.  Exposes
// f2d_launch_<nx>x<ny>(args, stream) (both directions) to dispatch.cu.
#include "fft2d_kernel.cuh"
#include "fft2d_tma_kernel.cuh"

#if !defined(F2D_NX) || !defined(F2D_NY)
#error "compile with -DF2D_NX=<nx> -DF2D_NY=<ny>"
#endif

#define MUFFT_CAT4_(a, b, c, d) a##b##c##d
#define MUFFT_CAT4(a, b, c, d) MUFFT_CAT4_(a, b, c, d)

namespace mufft {

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link against libcuda).
typedef CUresult (*tensor_map_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                         const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tensor_map_encode_fn tensor_map_encoder()
{
    static tensor_map_encode_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<tensor_map_encode_fn>(p);
    }();
    return fn;
}

// TMA-fed variant (fft2d_tma_kernel.cuh).  Returns cudaErrorNotSupported when shape, alignment or
// driver rule it out; the caller then takes the direct-load kernel.
template <int DIR, bool LITE>
static cudaError_t launch_tma(const mcu_2d_args &a, cudaStream_t stream)
{
  if constexpr (!f2t_supported(F2D_NX, F2D_NY)) {
    (void)a; (void)stream;
    return cudaErrorNotSupported;
  } else {
    constexpr size_t image = (size_t)F2D_NX * F2D_NY;
    if (a.out_image != image || (uintptr_t)a.out % 16 != 0 || (uintptr_t)a.in % 16 != 0 || (a.in_image * sizeof(float2)) % 16 != 0 ||
        (unsigned long long)a.batch * F2D_NY > 0xffffffffull)
        return cudaErrorNotSupported;
    tensor_map_encode_fn encode = tensor_map_encoder();
    if (encode == nullptr) return cudaErrorNotSupported;

    static int ctas = 0; // per instantiation
    auto kernel = fft2d_tma_kernel<F2D_NX, F2D_NY, DIR, LITE>;
    constexpr size_t smem = f2t_smem_bytes(F2D_NX, F2D_NY, LITE);
    if (ctas == 0) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        int dev = 0, sms = 0, per_sm = 0;
        if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
        if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
        if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, MUFFT_F2D_THREADS, smem)) != cudaSuccess) return err;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        ctas = sms * per_sm;
    }
    if (a.batch == 0) return cudaSuccess;
    // the output buffer as [batch * NY][NX] 8-byte elements, boxes of COLS columns x up to 256 rows
    CUtensorMap tmap;
    const cuuint64_t dims[2] = { F2D_NX, (cuuint64_t)a.batch * F2D_NY };
    const cuuint64_t strides[1] = { (cuuint64_t)F2D_NX * sizeof(float2) };
    const cuuint32_t box[2] = { (cuuint32_t)f2d_cols_per_tile(F2D_NY), (cuuint32_t)f2t_box_rows(F2D_NY) };
    const cuuint32_t elem_strides[2] = { 1, 1 };
    if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, a.out, dims, strides, box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorNotSupported;

    constexpr unsigned per_image = F2D_NY / f2d_rows_per_tile(F2D_NX) + F2D_NX / f2d_cols_per_tile(F2D_NY);
    const unsigned long long tickets = (unsigned long long)a.batch * per_image;
    // every CTA draws one ticket beyond the one it is working on, and two past the end
    if (tickets + 2ull * (unsigned long long)ctas >= 0xffffffffull) return cudaErrorInvalidConfiguration;
    const unsigned grid = tickets < (unsigned long long)ctas ? (unsigned)tickets : (unsigned)ctas;
    mcu_2d_args b = a;
    // tickets in flight: one being transformed and one being copied per CTA
    if (b.lag == 0) b.lag = 1 + (3 * grid + per_image - 1) / per_image;
    kernel<<<grid, MUFFT_F2D_THREADS, smem, stream>>>(b, tmap);
    return cudaGetLastError();
  }
}


// Shapes the fused kernel cannot take with the configured CTA size report "unsupported" to the
// host layer (which then uses the row kernel + column kernel path) instead of failing the build.
template <int DIR, int FLAGS>
static cudaError_t launch_dir(const mcu_2d_args &a, cudaStream_t stream)
{
  if constexpr (!f2d_supported(F2D_NX, F2D_NY)) {
    (void)a; (void)stream;
    return cudaErrorInvalidValue;
  } else {
    if (a.tma) {
        const cudaError_t err = launch_tma<DIR, FLAGS == F2D_LITE>(a, stream);
        if (err != cudaErrorNotSupported) return err;
    }
    static int ctas = 0; // per instantiation
    auto kernel = fft2d_fused_kernel<F2D_NX, F2D_NY, DIR, FLAGS>;
    constexpr size_t smem = f2d_smem_bytes(F2D_NX, F2D_NY, (FLAGS & F2D_SMEM_TW) != 0);
    if (ctas == 0) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        int dev = 0, sms = 0, per_sm = 0;
        if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
        if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
        if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, MUFFT_F2D_THREADS, smem)) != cudaSuccess) return err;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        ctas = sms * per_sm;
    }
    if (a.batch == 0) return cudaSuccess;
    const unsigned long long tickets = (unsigned long long)a.batch *
        (F2D_NY / (f2d_rows_per_tile(F2D_NX) * f2d_row_repeat(F2D_NX, F2D_NY)) + F2D_NX / f2d_cols_per_tile(F2D_NY));
    if (tickets + 2ull * (unsigned long long)ctas >= 0xffffffffull) return cudaErrorInvalidConfiguration; // F2D_EARLY draws two ahead
    const unsigned grid = tickets < (unsigned long long)ctas ? (unsigned)tickets : (unsigned)ctas;
    // lag L: the last row ticket of image i and its first column ticket are L*(NR+NC)-NR tickets
    // apart; keep that above ~1.5x the tickets in flight (one per CTA)
    constexpr unsigned per_image = F2D_NY / (f2d_rows_per_tile(F2D_NX) * f2d_row_repeat(F2D_NX, F2D_NY)) + F2D_NX / f2d_cols_per_tile(F2D_NY);
    mcu_2d_args b = a;
    if (b.lag == 0) b.lag = 1 + (3 * grid / 2 + per_image - 1) / per_image;
    kernel<<<grid, MUFFT_F2D_THREADS, smem, stream>>>(b);
    return cudaGetLastError();
  }
}

// Real images (forward R2C): one flavour, the one measured best for most complex shapes.
template <int OM>
static cudaError_t launch_real(const mcu_2d_args &a, cudaStream_t stream)
{
  if constexpr (!f2d_real_supported(F2D_NX, F2D_NY, OM)) {
    (void)a; (void)stream;
    return cudaErrorInvalidValue;
  } else {
    constexpr int FLAGS = F2D_ONE_SIDED | F2D_DERIVE;
    constexpr int VX = f2d_rows_variant(F2D_NX, OM);
    constexpr unsigned width = OM == MCU_OUT_R2C ? F2D_NX + 1 : 2 * F2D_NX;
    constexpr unsigned per_image = F2D_NY / f2d_rows_per_tile(F2D_NX) + (width + f2d_cols_per_tile(F2D_NY) - 1) / f2d_cols_per_tile(F2D_NY);
    static int ctas = 0; // per instantiation
    auto kernel = fft2d_fused_kernel<F2D_NX, F2D_NY, -1, FLAGS, OM>;
    constexpr size_t smem = f2d_smem_bytes(F2D_NX, F2D_NY, false, VX);
    if (ctas == 0) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        int dev = 0, sms = 0, per_sm = 0;
        if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
        if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
        if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, MUFFT_F2D_THREADS, smem)) != cudaSuccess) return err;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        ctas = sms * per_sm;
    }
    if (a.batch == 0) return cudaSuccess;
    const unsigned long long tickets = (unsigned long long)a.batch * per_image;
    if (tickets + (unsigned long long)ctas >= 0xffffffffull) return cudaErrorInvalidConfiguration;
    const unsigned grid = tickets < (unsigned long long)ctas ? (unsigned)tickets : (unsigned)ctas;
    mcu_2d_args b = a;
    if (b.lag == 0) b.lag = 1 + (3 * grid / 2 + per_image - 1) / per_image;
    kernel<<<grid, MUFFT_F2D_THREADS, smem, stream>>>(b);
    return cudaGetLastError();
  }
}

bool MUFFT_CAT4(f2d_supported_, F2D_NX, x, F2D_NY)() { return f2d_supported(F2D_NX, F2D_NY); }
bool MUFFT_CAT4(f2d_real_supported_, F2D_NX, x, F2D_NY)(int out_mode)
{
    return out_mode == MCU_OUT_R2C ? f2d_real_supported(F2D_NX, F2D_NY, MCU_OUT_R2C)
                                   : (out_mode == MCU_OUT_R2C_FULL ? f2d_real_supported(F2D_NX, F2D_NY, MCU_OUT_R2C_FULL) : false);
}

cudaError_t MUFFT_CAT4(f2d_launch_, F2D_NX, x, F2D_NY)(const mcu_2d_args &a, cudaStream_t stream)
{
    if (a.out_mode == MCU_OUT_R2C) return a.direction < 0 ? launch_real<MCU_OUT_R2C>(a, stream) : cudaErrorInvalidValue;
    if (a.out_mode == MCU_OUT_R2C_FULL) return a.direction < 0 ? launch_real<MCU_OUT_R2C_FULL>(a, stream) : cudaErrorInvalidValue;
    // Measured on B200 over all 16 shapes (profiles/README.md, round-1 flavour table): one-sided fences
    // + derived twiddles is the best or within 1.5 % of the best flavour everywhere except 1024 x 1024
    // (BASELINE configs[3]), where the twiddle tables in shared memory win: 1.279 ms plain, 1.194 / 1.209.
    const int flags = a.flags >= 0 ? a.flags : (F2D_NX == 1024 && F2D_NY == 1024 ? F2D_LITE : (F2D_ONE_SIDED | F2D_DERIVE));
    switch (flags) {
#define MUFFT_F2D_FLAVOUR(f) case f: return a.direction < 0 ? launch_dir<-1, f>(a, stream) : launch_dir<1, f>(a, stream);
        MUFFT_F2D_FLAVOUR(0)
        MUFFT_F2D_FLAVOUR(F2D_ONE_SIDED)
        MUFFT_F2D_FLAVOUR(F2D_LITE)
        MUFFT_F2D_FLAVOUR(F2D_ONE_SIDED | F2D_DERIVE)
        MUFFT_F2D_FLAVOUR(F2D_LITE | F2D_DERIVE)
        MUFFT_F2D_FLAVOUR(F2D_LITE | F2D_EARLY)
        MUFFT_F2D_FLAVOUR(F2D_ONE_SIDED | F2D_DERIVE | F2D_EARLY)
        MUFFT_F2D_FLAVOUR(F2D_ONE_SIDED | F2D_DERIVE | F2D_OCC3)
        MUFFT_F2D_FLAVOUR(F2D_ONE_SIDED | F2D_OCC3)
        MUFFT_F2D_FLAVOUR(F2D_LITE | F2D_OCC3)
        MUFFT_F2D_FLAVOUR(F2D_ONE_SIDED | F2D_DERIVE | F2D_OCC3 | F2D_INTERLEAVE)
        MUFFT_F2D_FLAVOUR(F2D_LITE | F2D_INTERLEAVE)
        MUFFT_F2D_FLAVOUR(F2D_ONE_SIDED | F2D_DERIVE | F2D_INTERLEAVE)
#undef MUFFT_F2D_FLAVOUR
    default: return cudaErrorInvalidValue;
    }
}

} // namespace mufft
