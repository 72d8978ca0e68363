// TEST INFRASTRUCTURE ONLY: the fused 2D kernel of mufft_b200/csrc/fft2d_kernel.cuh compiled for the
// CPU emulation at one image shape (-DEMU_NX=<nx> -DEMU_NY=<ny>).  `ctas` CTAs run concurrently
// (all resident, as the persistent launch on the GPU), so the ticket / tiles-done protocol between
// CTAs is exercised for real.
#include "fft2d_kernel.cuh"

using namespace mufft;
#define EMU_CAT_(a, b, c, d) a##b##c##d
#define EMU_CAT(a, b, c, d) EMU_CAT_(a, b, c, d)

template <int DIR, int KIND>
static int run_kind(const mcu_2d_args &a, unsigned ctas)
{
    constexpr int NX = EMU_NX, NY = EMU_NY;
    const emu::idx3 grid{ ctas, 1, 1 };
    emu::launch(grid, MUFFT_F2D_THREADS, f2d_smem_bytes(NX, NY, KIND), ctas, 1, [&] { fft2d_fused_kernel<NX, NY, DIR, KIND>(a); });
    return 0;
}

// struct-taking entry point (emu_abi.cpp and tests/test_emu.py): lag 0 = 1
extern "C" int EMU_CAT(emu_launch_2d_, EMU_NX, x, EMU_NY)(const mcu_2d_args *a, unsigned ctas)
{
    constexpr int NX = EMU_NX, NY = EMU_NY;
    if (!f2d_supported(NX, NY)) return -3;
    if (a->batch == 0) return 0;
    mcu_2d_args b = *a;
    if (b.lag == 0) b.lag = 1;
    switch (b.kind) {
    case MCU_2D_C2C: return b.direction < 0 ? run_kind<-1, F2D_C2C>(b, ctas) : run_kind<1, F2D_C2C>(b, ctas);
    case MCU_2D_R2C: return run_kind<-1, F2D_R2C>(b, ctas);
    case MCU_2D_R2C_FULL: return run_kind<-1, F2D_R2C_FULL>(b, ctas);
    case MCU_2D_C2R: {
        const size_t lag_eff = b.lag < b.batch ? b.lag : b.batch;
        if (b.mid == nullptr || b.ring == 0 || !(b.ring > lag_eff || b.ring >= b.batch)) return 1;
        return run_kind<1, F2D_C2R>(b, ctas);
    }
    default: return 1;
    }
}
extern "C" unsigned EMU_CAT(emu_2d_row_tiles_, EMU_NX, x, EMU_NY)() { return f2d_row_tiles(EMU_NX, EMU_NY); }
extern "C" int EMU_CAT(emu_2d_rows_variant_, EMU_NX, x, EMU_NY)(int kind) { return f2d_rows_variant(EMU_NX, kind); }
