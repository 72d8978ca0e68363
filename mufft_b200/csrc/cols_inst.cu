// One translation unit per column length (compiled with -DCOLS_N=<n>).
#include "cols_kernel.cuh"
#include "launch_util.cuh"

#ifndef COLS_N
#error "compile with -DCOLS_N=<power of two between 2 and 2048>"
#endif

#define MUFFT_CAT_(a, b) a##b
#define MUFFT_CAT(a, b) MUFFT_CAT_(a, b)

namespace mufft {

template <int DIR, bool TWIDDLE>
static cudaError_t launch_dir(const mcu_cols_args &a, cudaStream_t stream)
{
    constexpr int N = COLS_N;
    static launch_cfg cfg; // per instantiation, per device inside
    auto kernel = fft_cols_kernel<N, DIR, TWIDDLE>;
    {
        const cudaError_t err = launch_prepare(cfg, kernel, cols_smem_bytes(N), 0, nullptr);
        if (err != cudaSuccess) return err;
    }
    const size_t tiles = ((size_t)a.width + cols_tile(N) - 1) / cols_tile(N);
    if (tiles == 0 || a.batch == 0) return cudaSuccess;
    if (tiles > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)tiles, (unsigned)(a.batch > 65535 ? 65535 : a.batch));
    kernel<<<grid, cols_cta_threads(N), cols_smem_bytes(N), stream>>>(a);
    return cudaGetLastError();
}

cudaError_t MUFFT_CAT(cols_launch_, COLS_N)(const mcu_cols_args &a, cudaStream_t stream)
{
#if COLS_N == 256 || COLS_N == 512 || COLS_N == 1024
    if (a.roots != nullptr) return a.direction < 0 ? launch_dir<-1, true>(a, stream) : launch_dir<1, true>(a, stream);
#else
    if (a.roots != nullptr) return cudaErrorInvalidValue;
#endif
    return a.direction < 0 ? launch_dir<-1, false>(a, stream) : launch_dir<1, false>(a, stream);
}

} // namespace mufft
