// One translation unit per complex transform length (compiled with -DCONV_N=<n>): the fused
// convolution kernels of conv_kernel.cuh, stereo and mono, plain and zero-padded first operand.
#include "conv_kernel.cuh"
#include "launch_util.cuh"

#ifndef CONV_N
#error "compile with -DCONV_N=<power of two between 2 and 8192>"
#endif

#define MUFFT_CAT_(a, b) a##b
#define MUFFT_CAT(a, b) MUFFT_CAT_(a, b)

namespace mufft {

template <int IN_MODE, bool MONO>
static cudaError_t launch_conv(const mcu_conv_args &a, cudaStream_t stream)
{
    constexpr int N = CONV_N;
    constexpr int TPC = rows_per_cta(N, 0);
    const size_t groups = (a.batch + TPC - 1) / TPC;
    if (groups == 0) return cudaSuccess;
    if (groups > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    static launch_cfg cfg; // per instantiation, per device inside
    auto kernel = fft_conv_fused_kernel<N, IN_MODE, MONO>;
    {
        const cudaError_t err = launch_prepare(cfg, kernel, rows_smem_bytes(N, 0), 0, nullptr);
        if (err != cudaSuccess) return err;
    }
    kernel<<<(unsigned)groups, rows_cta_threads(N, 0), rows_smem_bytes(N, 0), stream>>>(a);
    return cudaGetLastError();
}

cudaError_t MUFFT_CAT(conv_launch_, CONV_N)(const mcu_conv_args &a, cudaStream_t stream)
{
    const bool half = a.in_mode == MCU_IN_HALF;
    if (a.in_mode != MCU_IN_HALF && a.in_mode != MCU_IN_COMPLEX) return cudaErrorInvalidValue;
    if (a.mono) return half ? launch_conv<MCU_IN_HALF, true>(a, stream) : launch_conv<MCU_IN_COMPLEX, true>(a, stream);
    return half ? launch_conv<MCU_IN_HALF, false>(a, stream) : launch_conv<MCU_IN_COMPLEX, false>(a, stream);
}

} // namespace mufft
