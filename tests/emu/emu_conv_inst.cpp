// TEST INFRASTRUCTURE ONLY: fused convolution kernel (conv_kernel.cuh) and the second kernel of the
// two-kernel long transform (rowst_kernel.cuh) for the emulated CUDA side, one length per translation
// unit (-DEMU_N=<n>; rowst only for 256 / 512 / 1024).
#include "conv_kernel.cuh"
#include "rowst_kernel.cuh"

using namespace mufft;
#define EMU_CAT_(a, b) a##b
#define EMU_CAT(a, b) EMU_CAT_(a, b)

extern "C" int EMU_CAT(emu_launch_conv_, EMU_N)(const mcu_conv_args *a)
{
    constexpr int N = EMU_N;
    constexpr int TPC = rows_per_cta(N, 0);
    const emu::idx3 grid{ (unsigned)((a->batch + TPC - 1) / TPC), 1, 1 };
    if (grid.x == 0) return 0;
    const unsigned threads = rows_cta_threads(N, 0);
    const size_t smem = rows_smem_bytes(N, 0);
    const bool half = a->in_mode == MCU_IN_HALF;
    if (a->in_mode != MCU_IN_HALF && a->in_mode != MCU_IN_COMPLEX) return 1;
    if (a->mono) {
        if (half) emu::launch(grid, threads, smem, 1, 1, [&] { fft_conv_fused_kernel<N, MCU_IN_HALF, true>(*a); });
        else emu::launch(grid, threads, smem, 1, 1, [&] { fft_conv_fused_kernel<N, MCU_IN_COMPLEX, true>(*a); });
    } else {
        if (half) emu::launch(grid, threads, smem, 1, 1, [&] { fft_conv_fused_kernel<N, MCU_IN_HALF, false>(*a); });
        else emu::launch(grid, threads, smem, 1, 1, [&] { fft_conv_fused_kernel<N, MCU_IN_COMPLEX, false>(*a); });
    }
    return 0;
}

#if EMU_N == 256 || EMU_N == 512 || EMU_N == 1024
extern "C" int EMU_CAT(emu_launch_rowst_, EMU_N)(const mcu_rowst_args *a)
{
    constexpr int N2 = EMU_N;
    const emu::idx3 grid{ a->n1 / MUFFT_ROWST_TILE, (unsigned)(a->batch > 65535 ? 65535 : a->batch), 1 };
    if (grid.x == 0 || grid.y == 0) return 0;
    if (a->direction < 0) emu::launch(grid, rowst_cta_threads(N2), rowst_smem_bytes(N2), 1, 1, [&] { fft_rowst_kernel<N2, -1>(*a); });
    else emu::launch(grid, rowst_cta_threads(N2), rowst_smem_bytes(N2), 1, 1, [&] { fft_rowst_kernel<N2, 1>(*a); });
    return 0;
}
#endif
