// One translation unit per transform length (compiled with -DROWS_N=<n>) so that the 13 sizes
// build in parallel.  Exposes rows_launch_<n>(args, stream) to dispatch.cu.
#include "rows_kernel.cuh"

#ifndef ROWS_N
#error "compile with -DROWS_N=<power of two between 2 and 8192>"
#endif

#define MUFFT_CAT_(a, b) a##b
#define MUFFT_CAT(a, b) MUFFT_CAT_(a, b)

namespace mufft {

template <int DIR>
static cudaError_t launch_dir(const mcu_rows_args &a, cudaStream_t stream)
{
    constexpr int N = ROWS_N;
    static bool configured = false; // per direction; the attribute call is idempotent, so a race is harmless
    auto kernel = fft_rows_kernel<N, DIR>;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rows_smem_bytes(N));
        if (err != cudaSuccess) return err;
        configured = true;
    }
    const size_t ctas = (a.batch + rows_per_cta(N) - 1) / rows_per_cta(N);
    if (ctas == 0) return cudaSuccess;
    if (ctas > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    kernel<<<(unsigned)ctas, rows_cta_threads(N), rows_smem_bytes(N), stream>>>(a);
    return cudaGetLastError();
}

cudaError_t MUFFT_CAT(rows_launch_, ROWS_N)(const mcu_rows_args &a, cudaStream_t stream)
{
    return a.direction < 0 ? launch_dir<-1>(a, stream) : launch_dir<1>(a, stream);
}

} // namespace mufft
