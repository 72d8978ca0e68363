// One translation unit per image shape: compiled with -DF2D_NX=<nx> -DF2D_NY=<ny>.  Exposes
// f2d_launch_<nx>x<ny>(args, stream) (both directions) to dispatch.cu.
#include "fft2d_kernel.cuh"

#if !defined(F2D_NX) || !defined(F2D_NY)
#error "compile with -DF2D_NX=<nx> -DF2D_NY=<ny>"
#endif

#define MUFFT_CAT4_(a, b, c, d) a##b##c##d
#define MUFFT_CAT4(a, b, c, d) MUFFT_CAT4_(a, b, c, d)

namespace mufft {


// Shapes the fused kernel cannot take with the configured CTA size report "unsupported" to the
// host layer (which then uses the row kernel + column kernel path) instead of failing the build.
template <int DIR>
static cudaError_t launch_dir(const mcu_2d_args &a, cudaStream_t stream)
{
  if constexpr (!f2d_supported(F2D_NX, F2D_NY)) {
    (void)a; (void)stream;
    return cudaErrorInvalidValue;
  } else {
    static int ctas = 0; // per direction
    auto kernel = fft2d_fused_kernel<F2D_NX, F2D_NY, DIR>;
    constexpr size_t smem = f2d_smem_bytes(F2D_NX, F2D_NY);
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
    if (tickets + (unsigned long long)ctas >= 0xffffffffull) return cudaErrorInvalidConfiguration;
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

bool MUFFT_CAT4(f2d_supported_, F2D_NX, x, F2D_NY)() { return f2d_supported(F2D_NX, F2D_NY); }

cudaError_t MUFFT_CAT4(f2d_launch_, F2D_NX, x, F2D_NY)(const mcu_2d_args &a, cudaStream_t stream)
{
    return a.direction < 0 ? launch_dir<-1>(a, stream) : launch_dir<1>(a, stream);
}

} // namespace mufft
