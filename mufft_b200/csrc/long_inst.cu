// One translation unit per shape of the single-launch four-step kernel (long_kernel.cuh): compiled with
// -DLONG_N1=<16|32> -DLONG_N2=<1024|2048>.  Exposes long_launch_<n1>x<n2>(args, stream) and
// long_ctas_<n1>x<n2>() (scratch slots worth allocating on the current device) to dispatch.cu.
#include "long_kernel.cuh"
#include "launch_util.cuh"
#ifndef MUFFT_LONG_TMA_A
#define MUFFT_LONG_TMA_A 1 // inputs of step A through the TMA unit (long_kernel.cuh): 32768 points 0.575 -> 0.612 (profiles/r02ah_ab.txt)
#endif

#if !defined(LONG_N1) || !defined(LONG_N2)
#error "compile with -DLONG_N1=<n1> -DLONG_N2=<n2>"
#endif
#define MUFFT_CAT4_(a, b, c, d) a##b##c##d
#define MUFFT_CAT4(a, b, c, d) MUFFT_CAT4_(a, b, c, d)

namespace mufft {

// A shape the configured tile (MUFFT_LONG_TR) cannot take reports "unsupported": the host layer then uses the
// two-kernel split for that length instead of failing the build.
template <int DIR, int IN_MODE, bool TO_PEERS>
static cudaError_t launch_long(const mcu_long_args &a, cudaStream_t stream, int *ctas_only)
{
  if constexpr (!long_supported(LONG_N1, LONG_N2, TO_PEERS)) {
    (void)a; (void)stream; (void)ctas_only;
    return cudaErrorInvalidValue;
  } else {
    static launch_cfg cfg;
    constexpr int SLOTS = 1, THREADS = long_cta_threads(LONG_N2, TO_PEERS);
    auto kernel = fft_long_kernel<LONG_N1, LONG_N2, DIR, IN_MODE, TO_PEERS>;
    constexpr size_t smem = long_smem_bytes(LONG_N2, TO_PEERS);
    int resident = 0;
    cudaError_t err = launch_prepare(cfg, kernel, smem, THREADS, &resident);
    if (err != cudaSuccess) return err;
    if (ctas_only != nullptr) {
        *ctas_only = resident * SLOTS; // scratch slots worth allocating
        return cudaSuccess;
    }
    if (a.batch == 0) return cudaSuccess;
    size_t grid = a.batch < (size_t)resident ? a.batch : (size_t)resident;
    if (a.scratch_slots / SLOTS < grid) grid = a.scratch_slots / SLOTS;
    if (grid == 0) return cudaErrorInvalidValue;
#if MUFFT_LONG_TMA_A
    // inputs of step A through the TMA unit: 16-byte aligned transforms
    if ((reinterpret_cast<uintptr_t>(a.in) & 15) == 0 && (a.in_batch & 1) == 0) {
        static launch_cfg tcfg;
        auto tkernel = fft_long_kernel<LONG_N1, LONG_N2, DIR, IN_MODE, TO_PEERS, true>;
        int tres = 0;
        err = launch_prepare(tcfg, tkernel, smem + 32, THREADS, &tres);
        if (err != cudaSuccess) return err;
        if ((size_t)tres < grid) grid = (size_t)tres;
        tkernel<<<(unsigned)grid, THREADS, smem + 32, stream>>>(a);
        return cudaGetLastError();
    }
#endif
    kernel<<<(unsigned)grid, THREADS, smem, stream>>>(a);
    return cudaGetLastError();
  }
}

// All instantiations of a shape have the same launch bounds and shared memory, hence the same occupancy.
int MUFFT_CAT4(long_ctas_, LONG_N1, x, LONG_N2)()
{
    int ctas = 0;
    mcu_long_args none = {};
    if (launch_long<-1, MCU_IN_COMPLEX, false>(none, nullptr, &ctas) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return ctas;
}

cudaError_t MUFFT_CAT4(long_launch_, LONG_N1, x, LONG_N2)(const mcu_long_args &a, cudaStream_t s)
{
    if (a.to_peers) {
        if (a.in_mode != MCU_IN_COMPLEX) return cudaErrorInvalidValue;
        return a.direction < 0 ? launch_long<-1, MCU_IN_COMPLEX, true>(a, s, nullptr) : launch_long<1, MCU_IN_COMPLEX, true>(a, s, nullptr);
    }
    if (a.in_mode == MCU_IN_HALF)
        return a.direction < 0 ? launch_long<-1, MCU_IN_HALF, false>(a, s, nullptr) : launch_long<1, MCU_IN_HALF, false>(a, s, nullptr);
    if (a.in_mode != MCU_IN_COMPLEX) return cudaErrorInvalidValue;
    return a.direction < 0 ? launch_long<-1, MCU_IN_COMPLEX, false>(a, s, nullptr) : launch_long<1, MCU_IN_COMPLEX, false>(a, s, nullptr);
}

} // namespace mufft
