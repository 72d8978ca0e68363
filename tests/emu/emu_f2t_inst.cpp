// TEST INFRASTRUCTURE ONLY: the TMA-fed fused 2D kernel of mufft_b200/csrc/fft2d_tma_kernel.cuh compiled
// for the CPU emulation at one image shape (-DEMU_NX=<nx> -DEMU_NY=<ny>); bulk copies and the
// mbarrier are the synchronous stand-ins of tma.cuh.  `ctas` CTAs run concurrently.
#include "fft2d_tma_kernel.cuh"

using namespace mufft;
#define EMU_CAT_(a, b, c, d) a##b##c##d
#define EMU_CAT(a, b, c, d) EMU_CAT_(a, b, c, d)

extern "C" int EMU_CAT(emu_2d_tma_, EMU_NX, x, EMU_NY)(int direction, void *out, const void *in, const void *tw_x, const void *tw_y,
                                                      unsigned *counters, size_t batch, unsigned lag, unsigned ctas, int lite)
{
    constexpr int NX = EMU_NX, NY = EMU_NY;
    if (!f2t_supported(NX, NY)) return -3;
    mcu_2d_args a;
    memset(&a, 0, sizeof(a));
    a.out = out; a.in = in; a.tw_x = tw_x; a.tw_y = tw_y; a.counters = counters;
    a.out_image = (size_t)NX * NY; a.in_image = (size_t)NX * NY; a.batch = batch; a.nx = NX; a.ny = NY;
    a.direction = direction; a.lag = lag;
    tensor_map_t map;
    map.base = static_cast<const unsigned char *>(out);
    map.pitch_bytes = (size_t)NX * 8;
    map.box_w = f2d_cols_per_tile(NY);
    map.box_h = f2t_box_rows(NY);
    const emu::idx3 grid{ ctas, 1, 1 };
    if (lite) {
        if (direction < 0) emu::launch(grid, MUFFT_F2D_THREADS, f2t_smem_bytes(NX, NY, true), ctas, 1, [&] { fft2d_tma_kernel<NX, NY, -1, true>(a, map); });
        else emu::launch(grid, MUFFT_F2D_THREADS, f2t_smem_bytes(NX, NY, true), ctas, 1, [&] { fft2d_tma_kernel<NX, NY, 1, true>(a, map); });
    } else {
        if (direction < 0) emu::launch(grid, MUFFT_F2D_THREADS, f2t_smem_bytes(NX, NY), ctas, 1, [&] { fft2d_tma_kernel<NX, NY, -1>(a, map); });
        else emu::launch(grid, MUFFT_F2D_THREADS, f2t_smem_bytes(NX, NY), ctas, 1, [&] { fft2d_tma_kernel<NX, NY, 1>(a, map); });
    }
    return 0;
}
