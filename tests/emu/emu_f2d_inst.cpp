// TEST INFRASTRUCTURE ONLY: the fused 2D kernel of mufft_b200/csrc/fft2d_kernel.cuh compiled for the
// CPU emulation at one image shape (-DEMU_NX=<nx> -DEMU_NY=<ny>).  `ctas` CTAs run concurrently
// (all resident, as the persistent launch on the GPU), so the ticket / rows-done protocol between
// CTAs is exercised for real.
#include "fft2d_kernel.cuh"

using namespace mufft;
#define EMU_CAT_(a, b, c, d) a##b##c##d
#define EMU_CAT(a, b, c, d) EMU_CAT_(a, b, c, d)

extern "C" int EMU_CAT(emu_2d_fused_, EMU_NX, x, EMU_NY)(int direction, void *out, const void *in, const void *tw_x, const void *tw_y,
                                                        unsigned *counters, size_t batch, unsigned lag, unsigned ctas, int lite)
{
    constexpr int NX = EMU_NX, NY = EMU_NY;
    if (!f2d_supported(NX, NY)) return -3;
    mcu_2d_args a;
    memset(&a, 0, sizeof(a));
    a.out = out; a.in = in; a.tw_x = tw_x; a.tw_y = tw_y; a.counters = counters;
    a.out_image = (size_t)NX * NY; a.in_image = (size_t)NX * NY; a.batch = batch; a.nx = NX; a.ny = NY;
    a.direction = direction; a.lag = lag;
    const emu::idx3 grid{ ctas, 1, 1 };
    if (lite == 1) { // twiddle tables in shared memory + one-sided fences
        if (direction < 0) emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY, true), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, -1, F2D_LITE>(a); });
        else emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY, true), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, 1, F2D_LITE>(a); });
    } else if (lite == 2) { // ... + derived twiddles
        if (direction < 0) emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY, true), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, -1, F2D_LITE | F2D_DERIVE>(a); });
        else emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY, true), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, 1, F2D_LITE | F2D_DERIVE>(a); });
    } else if (lite == 4) { // smem tables + one-sided fences + tickets two ahead
        if (direction < 0) emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY, true), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, -1, F2D_LITE | F2D_EARLY>(a); });
        else emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY, true), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, 1, F2D_LITE | F2D_EARLY>(a); });
    } else if (lite == 5) { // smem tables + one-sided fences, row and column tickets alternating
        if (direction < 0) emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY, true), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, -1, F2D_LITE | F2D_INTERLEAVE>(a); });
        else emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY, true), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, 1, F2D_LITE | F2D_INTERLEAVE>(a); });
    } else if (lite == 6) { // derived twiddles from the global table, tickets alternating
        if (direction < 0) emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, -1, F2D_ONE_SIDED | F2D_DERIVE | F2D_INTERLEAVE>(a); });
        else emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, 1, F2D_ONE_SIDED | F2D_DERIVE | F2D_INTERLEAVE>(a); });
    } else if (lite == 3) { // derived twiddles from the global table
        if (direction < 0) emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, -1, F2D_ONE_SIDED | F2D_DERIVE>(a); });
        else emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, 1, F2D_ONE_SIDED | F2D_DERIVE>(a); });
    } else {
        if (direction < 0) emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, -1>(a); });
        else emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, 1>(a); });
    }
    return 0;
}

// Real images (forward R2C, half or full spectrum): EMU_NX = complex points per row transform (the rows
// have 2 EMU_NX reals), output rows of pitch 2 EMU_NX bins.
extern "C" int EMU_CAT(emu_2d_fused_real_, EMU_NX, x, EMU_NY)(int full, void *out, const void *in, const void *tw_x, const void *tw_y, const void *rtw,
                                                             unsigned *counters, size_t batch, unsigned lag, unsigned ctas)
{
    constexpr int NX = EMU_NX, NY = EMU_NY;
    mcu_2d_args a;
    memset(&a, 0, sizeof(a));
    a.out = out; a.in = in; a.tw_x = tw_x; a.tw_y = tw_y; a.rtw = rtw; a.counters = counters;
    a.out_image = 2 * (size_t)NX * NY; a.in_image = (size_t)NX * NY; a.batch = batch; a.nx = NX; a.ny = NY;
    a.direction = -1; a.lag = lag; a.out_mode = full ? MCU_OUT_R2C_FULL : MCU_OUT_R2C;
    const emu::idx3 grid{ ctas, 1, 1 };
    constexpr int FLAGS = F2D_ONE_SIDED | F2D_DERIVE;
    if (full) {
        if constexpr (f2d_real_supported(NX, NY, MCU_OUT_R2C_FULL)) {
            emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY, false, f2d_rows_variant(NX, MCU_OUT_R2C_FULL)), ctas, 1,
                        [&] { fft2d_fused_kernel<NX, NY, -1, FLAGS, MCU_OUT_R2C_FULL>(a); });
            return 0;
        }
    } else {
        if constexpr (f2d_real_supported(NX, NY, MCU_OUT_R2C)) {
            emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY, false, f2d_rows_variant(NX, MCU_OUT_R2C)), ctas, 1,
                        [&] { fft2d_fused_kernel<NX, NY, -1, FLAGS, MCU_OUT_R2C>(a); });
            return 0;
        }
    }
    return -3;
}

// struct-taking entry point for the emulated CUDA side (emu_abi.cpp): the flavour the product launcher would pick
extern "C" int EMU_CAT(emu_launch_2d_, EMU_NX, x, EMU_NY)(const mcu_2d_args *a, unsigned ctas)
{
    constexpr int NX = EMU_NX, NY = EMU_NY;
    mcu_2d_args b = *a;
    if (b.lag == 0) b.lag = 1;
    const emu::idx3 grid{ ctas, 1, 1 };
    if (b.out_mode == MCU_OUT_R2C || b.out_mode == MCU_OUT_R2C_FULL)
        return EMU_CAT(emu_2d_fused_real_, EMU_NX, x, EMU_NY)(b.out_mode == MCU_OUT_R2C_FULL, b.out, b.in, b.tw_x, b.tw_y, b.rtw, b.counters, b.batch, b.lag, ctas);
    constexpr int FLAGS = F2D_ONE_SIDED | F2D_DERIVE;
    if (b.direction < 0) emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, -1, FLAGS>(b); });
    else emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, 1, FLAGS>(b); });
    return 0;
}
