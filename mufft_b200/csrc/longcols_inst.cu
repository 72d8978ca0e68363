// The single-launch four-step column kernel (longcols_kernel.cuh) for every supported first factor; exposes
// longcols_launch(args, stream) and longcols_lag(n1) to dispatch.cu.
#include "longcols_kernel.cuh"
#include "launch_util.cuh"

namespace mufft {

template <int N1, int DIR>
static cudaError_t launch_longcols(const mcu_longcols_args &a, cudaStream_t stream, unsigned *lag_only)
{
    static launch_cfg cfg;
    auto kernel = fft_longcols_kernel<N1, DIR>;
    constexpr size_t smem = longcols_smem_bytes();
    constexpr unsigned per_block = longcols_a_tiles() + longcols_b_tiles(N1);
    int ctas = 0;
    cudaError_t err = launch_prepare(cfg, kernel, smem, LONGCOLS_THREADS, &ctas);
    if (err != cudaSuccess) return err;
    // the last A ticket of a block and its first B ticket are L * per_block - NP tickets apart: ~1.5x the tickets in flight
    const unsigned lag = 1 + (3 * (unsigned)ctas / 2 + per_block - 1) / per_block;
    if (lag_only != nullptr) {
        *lag_only = lag;
        return cudaSuccess;
    }
    if (a.width == 0) return cudaSuccess;
    const unsigned blocks = (a.width + LONGCOLS_CB - 1) / LONGCOLS_CB;
    const unsigned long long tickets = (unsigned long long)blocks * per_block;
    if (tickets + (unsigned long long)ctas >= 0xffffffffull) return cudaErrorInvalidConfiguration;
    mcu_longcols_args b = a;
    if (b.lag == 0) b.lag = lag;
    const unsigned lag_eff = b.lag < blocks ? b.lag : blocks;
    if (b.scratch == nullptr || b.ring == 0 || !(b.ring > lag_eff || b.ring >= blocks)) return cudaErrorInvalidValue;
    const unsigned grid = tickets < (unsigned long long)ctas ? (unsigned)tickets : (unsigned)ctas;
    kernel<<<grid, LONGCOLS_THREADS, smem, stream>>>(b);
    return cudaGetLastError();
}

template <int N1>
static cudaError_t launch_n1(const mcu_longcols_args &a, cudaStream_t stream, unsigned *lag_only)
{
    return a.direction < 0 ? launch_longcols<N1, -1>(a, stream, lag_only) : launch_longcols<N1, 1>(a, stream, lag_only);
}

cudaError_t longcols_launch(const mcu_longcols_args &a, cudaStream_t stream, unsigned *lag_only)
{
    switch (a.n1) {
    case 16: return launch_n1<16>(a, stream, lag_only);
    case 32: return launch_n1<32>(a, stream, lag_only);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace mufft
