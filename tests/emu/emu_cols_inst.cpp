// TEST INFRASTRUCTURE ONLY: the column kernel of mufft_b200/csrc/cols_kernel.cuh compiled for the CPU
// emulation at one column length (-DEMU_N=<n>).
#include "cols_kernel.cuh"

using namespace mufft;
#define EMU_CAT_(a, b) a##b
#define EMU_CAT(a, b) EMU_CAT_(a, b)

extern "C" int EMU_CAT(emu_cols_, EMU_N)(int direction, void *out, const void *in, const void *tw, size_t out_pitch, size_t in_pitch,
                                         size_t out_image, size_t in_image, size_t batch, unsigned width, unsigned rows_valid)
{
    constexpr int N = EMU_N;
    mcu_cols_args a;
    memset(&a, 0, sizeof(a));
    a.out = out; a.in = in; a.tw = tw; a.out_pitch = out_pitch; a.in_pitch = in_pitch; a.out_image = out_image; a.in_image = in_image;
    a.batch = batch; a.n = N; a.width = width; a.direction = direction; a.rows_valid = rows_valid;
    const emu::idx3 grid{ (unsigned)((width + cols_tile(N) - 1) / cols_tile(N)), (unsigned)batch, 1 };
    if (direction < 0) emu::launch(grid, cols_cta_threads(N), cols_smem_bytes(N), 1, 1, [&] { fft_cols_kernel<N, -1, false>(a); });
    else emu::launch(grid, cols_cta_threads(N), cols_smem_bytes(N), 1, 1, [&] { fft_cols_kernel<N, 1, false>(a); });
    return 0;
}

extern "C" int EMU_CAT(emu_launch_cols_, EMU_N)(const mcu_cols_args *a)
{
    constexpr int N = EMU_N;
    const emu::idx3 grid{ (unsigned)((a->width + cols_tile(N) - 1) / cols_tile(N)), (unsigned)(a->batch > 65535 ? 65535 : a->batch), 1 };
    if (grid.x == 0 || grid.y == 0) return 0;
    if (a->roots != nullptr) {
        if constexpr (N == 256 || N == 512 || N == 1024) {
            if (a->direction < 0) emu::launch(grid, cols_cta_threads(N), cols_smem_bytes(N), 1, 1, [&] { fft_cols_kernel<N, -1, true>(*a); });
            else emu::launch(grid, cols_cta_threads(N), cols_smem_bytes(N), 1, 1, [&] { fft_cols_kernel<N, 1, true>(*a); });
            return 0;
        }
        return 1;
    }
    if (a->direction < 0) emu::launch(grid, cols_cta_threads(N), cols_smem_bytes(N), 1, 1, [&] { fft_cols_kernel<N, -1, false>(*a); });
    else emu::launch(grid, cols_cta_threads(N), cols_smem_bytes(N), 1, 1, [&] { fft_cols_kernel<N, 1, false>(*a); });
    return 0;
}
