// Compiled with -DROWST_N=<n2>; exposes rowst_launch_<n2>(args, stream) to dispatch.cu.
#include "rowst_kernel.cuh"
#include "launch_util.cuh"

#ifndef ROWST_N
#error "compile with -DROWST_N=<256|512|1024>"
#endif
#define MUFFT_CAT_(a, b) a##b
#define MUFFT_CAT(a, b) MUFFT_CAT_(a, b)

namespace mufft {

template <int DIR>
static cudaError_t launch_dir(const mcu_rowst_args &a, cudaStream_t stream)
{
    static launch_cfg cfg; // per instantiation, per device inside
    auto kernel = fft_rowst_kernel<ROWST_N, DIR>;
    {
        const cudaError_t err = launch_prepare(cfg, kernel, rowst_smem_bytes(ROWST_N), 0, nullptr);
        if (err != cudaSuccess) return err;
    }
    if (a.batch == 0) return cudaSuccess;
    dim3 grid(a.n1 / MUFFT_ROWST_TILE, (unsigned)(a.batch > 65535 ? 65535 : a.batch));
    kernel<<<grid, rowst_cta_threads(ROWST_N), rowst_smem_bytes(ROWST_N), stream>>>(a);
    return cudaGetLastError();
}

cudaError_t MUFFT_CAT(rowst_launch_, ROWST_N)(const mcu_rowst_args &a, cudaStream_t stream)
{
    return a.direction < 0 ? launch_dir<-1>(a, stream) : launch_dir<1>(a, stream);
}

} // namespace mufft
