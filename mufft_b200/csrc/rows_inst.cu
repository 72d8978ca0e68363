// One translation unit per (transform length, direction): compiled with -DROWS_N=<n> -DROWS_DIR=<-1|1>
// so that the sizes build in parallel.  Exposes rows_launch_<n>_{fwd,inv}(args, stream) to dispatch.cu.
//
// Every (input mode, output mode) pair the host layer can ask for is a separate kernel instantiation.
#include "rows_kernel.cuh"
#include "launch_util.cuh"
#ifndef MUFFT_ROWS_PREFETCH
#define MUFFT_ROWS_PREFETCH 1 // large batches of 2048 / 8192 / 16384-point transforms go to the persistent prefetching kernel (rows_prefetch_kernel.cuh)
#endif
#if MUFFT_ROWS_PREFETCH && (ROWS_N == 16384 || ROWS_N == 8192 || ROWS_N == 2048)
#include "rows_prefetch_kernel.cuh"
#endif
#ifndef MUFFT_ROWS_PREFETCH_MIN_BATCH
#define MUFFT_ROWS_PREFETCH_MIN_BATCH 1024
#endif

#if !defined(ROWS_N) || !defined(ROWS_DIR)
#error "compile with -DROWS_N=<power of two between 2 and 16384> -DROWS_DIR=<-1 or 1>"
#endif

#define MUFFT_CAT_(a, b, c) a##b##c
#define MUFFT_CAT(a, b, c) MUFFT_CAT_(a, b, c)

namespace mufft {

template <int IN_MODE, int OUT_MODE>
static cudaError_t launch_mode(const mcu_rows_args &a, cudaStream_t stream)
{
    constexpr int N = ROWS_N;
    constexpr int DIR = ROWS_DIR;
    constexpr int V = rows_variant(N, IN_MODE, OUT_MODE);
    constexpr int TPC = rows_per_cta(N, V);
    const size_t groups = (a.batch + TPC - 1) / TPC;
    if (groups == 0) return cudaSuccess;
    if (groups > 0x7fffffffull) return cudaErrorInvalidConfiguration;
#if MUFFT_ROWS_PREFETCH && (ROWS_N == 16384 || ROWS_N == 8192 || ROWS_N == 2048)
    if constexpr (rows_prefetch_supported(N, IN_MODE, OUT_MODE)) {
        // bulk copies: 16-byte aligned transforms (the C2R flavour aligns its copies itself: 8 bytes are enough)
        if (a.batch >= MUFFT_ROWS_PREFETCH_MIN_BATCH &&
            (reinterpret_cast<uintptr_t>(a.in) & 15) == 0 && (IN_MODE == MCU_IN_C2R || (a.in_stride & 1) == 0)) {
            static launch_cfg pcfg;
            auto pkernel = fft_rows_prefetch_kernel<N, DIR, IN_MODE, OUT_MODE>;
            int ctas = 0;
            cudaError_t perr = launch_prepare(pcfg, pkernel, rows_prefetch_smem_bytes(N, V), rows_cta_threads(N, V), &ctas);
            if (perr != cudaSuccess) return perr;
            pkernel<<<(unsigned)ctas, rows_cta_threads(N, V), rows_prefetch_smem_bytes(N, V), stream>>>(a);
            return cudaGetLastError();
        }
    }
#endif
    static launch_cfg cfg; // per instantiation, per device inside
    auto kernel = fft_rows_kernel<N, DIR, IN_MODE, OUT_MODE>;
    cudaError_t err = launch_prepare(cfg, kernel, rows_smem_bytes(N, V), rows_cta_threads(N, V), nullptr);
    if (err != cudaSuccess) return err;
    kernel<<<(unsigned)groups, rows_cta_threads(N, V), rows_smem_bytes(N, V), stream>>>(a);
    return cudaGetLastError();
}

#if ROWS_DIR < 0
cudaError_t MUFFT_CAT(rows_launch_, ROWS_N, _fwd)(const mcu_rows_args &a, cudaStream_t s)
{
    switch (a.in_mode * 8 + a.out_mode) {
    case MCU_IN_COMPLEX * 8 + MCU_OUT_COMPLEX: return launch_mode<MCU_IN_COMPLEX, MCU_OUT_COMPLEX>(a, s);
    case MCU_IN_COMPLEX * 8 + MCU_OUT_PEERS: return launch_mode<MCU_IN_COMPLEX, MCU_OUT_PEERS>(a, s);
    case MCU_IN_COMPLEX * 8 + MCU_OUT_R2C: return launch_mode<MCU_IN_COMPLEX, MCU_OUT_R2C>(a, s);
    case MCU_IN_COMPLEX * 8 + MCU_OUT_R2C_FULL: return launch_mode<MCU_IN_COMPLEX, MCU_OUT_R2C_FULL>(a, s);
    case MCU_IN_COMPLEX * 8 + MCU_OUT_R2C_PEERS: return launch_mode<MCU_IN_COMPLEX, MCU_OUT_R2C_PEERS>(a, s);
    case MCU_IN_HALF * 8 + MCU_OUT_COMPLEX: return launch_mode<MCU_IN_HALF, MCU_OUT_COMPLEX>(a, s);
    case MCU_IN_HALF * 8 + MCU_OUT_R2C: return launch_mode<MCU_IN_HALF, MCU_OUT_R2C>(a, s);
    case MCU_IN_HALF * 8 + MCU_OUT_R2C_FULL: return launch_mode<MCU_IN_HALF, MCU_OUT_R2C_FULL>(a, s);
    default: return cudaErrorInvalidValue;
    }
}
#else
cudaError_t MUFFT_CAT(rows_launch_, ROWS_N, _inv)(const mcu_rows_args &a, cudaStream_t s)
{
    if (a.out_mode == MCU_OUT_PEERS && a.in_mode == MCU_IN_COMPLEX) return launch_mode<MCU_IN_COMPLEX, MCU_OUT_PEERS>(a, s);
    if (a.out_mode != MCU_OUT_COMPLEX) return cudaErrorInvalidValue;
    switch (a.in_mode) {
    case MCU_IN_COMPLEX: return launch_mode<MCU_IN_COMPLEX, MCU_OUT_COMPLEX>(a, s);
    case MCU_IN_HALF: return launch_mode<MCU_IN_HALF, MCU_OUT_COMPLEX>(a, s);
    case MCU_IN_C2R: return launch_mode<MCU_IN_C2R, MCU_OUT_COMPLEX>(a, s);
    case MCU_IN_MUL: return launch_mode<MCU_IN_MUL, MCU_OUT_COMPLEX>(a, s);
    case MCU_IN_MUL_C2R: return launch_mode<MCU_IN_MUL_C2R, MCU_OUT_COMPLEX>(a, s);
    case MCU_IN_MULACC: return launch_mode<MCU_IN_MULACC, MCU_OUT_COMPLEX>(a, s);
    case MCU_IN_MULACC_C2R: return launch_mode<MCU_IN_MULACC_C2R, MCU_OUT_COMPLEX>(a, s);
    default: return cudaErrorInvalidValue;
    }
}
#endif

} // namespace mufft
