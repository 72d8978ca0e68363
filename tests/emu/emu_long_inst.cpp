// TEST INFRASTRUCTURE ONLY: the two kernels of the long-transform path (N = N1 * N2: column kernel on the
// [N1][N2] view, then rowst_kernel.cuh) and the fused convolution kernel compiled for the CPU emulation.
// -DEMU_N1=<n1> -DEMU_N2=<n2> -DEMU_CONV_N=<n>.
#include "cols_kernel.cuh"
#include "rowst_kernel.cuh"
#include "conv_kernel.cuh"

using namespace mufft;
#define EMU_CAT_(a, b) a##b
#define EMU_CAT(a, b) EMU_CAT_(a, b)
#define EMU_CAT3_(a, b, c, d) a##b##c##d
#define EMU_CAT3(a, b, c, d) EMU_CAT3_(a, b, c, d)

// Whole N1*N2-point transforms: in -> scratch (kernel A) -> out (kernel B), as mufft_host.c: run_rows does.
extern "C" int EMU_CAT3(emu_two_kernel_, EMU_N1, x, EMU_N2)(int direction, void *out, void *scratch, const void *in, const void *tw_n1, const void *tw_n2,
                                                           const void *roots, size_t batch, int half)
{
    constexpr int N1 = EMU_N1, N2 = EMU_N2;
    constexpr unsigned n = N1 * N2;
    mcu_cols_args ca;
    memset(&ca, 0, sizeof(ca));
    ca.out = scratch; ca.in = in; ca.tw = tw_n1; ca.out_pitch = N2; ca.in_pitch = N2; ca.out_image = n; ca.in_image = half ? n / 2 : n;
    ca.batch = batch; ca.n = N1; ca.width = N2; ca.direction = direction; ca.rows_valid = half ? N1 / 2 : 0;
    const emu::idx3 cgrid{ (unsigned)((N2 + cols_tile(N1) - 1) / cols_tile(N1)), (unsigned)batch, 1 };
    if (direction < 0) emu::launch(cgrid, cols_cta_threads(N1), cols_smem_bytes(N1), 1, 1, [&] { fft_cols_kernel<N1, -1, false>(ca); });
    else emu::launch(cgrid, cols_cta_threads(N1), cols_smem_bytes(N1), 1, 1, [&] { fft_cols_kernel<N1, 1, false>(ca); });
    mcu_rowst_args ra;
    memset(&ra, 0, sizeof(ra));
    ra.out = out; ra.in = scratch; ra.tw = tw_n2; ra.roots = roots; ra.out_batch = n; ra.in_batch = n;
    ra.batch = batch; ra.n = n; ra.n1 = N1; ra.n2 = N2; ra.direction = direction;
    const emu::idx3 rgrid{ N1 / MUFFT_ROWST_TILE, (unsigned)batch, 1 };
    if (direction < 0) emu::launch(rgrid, rowst_cta_threads(N2), rowst_smem_bytes(N2), 1, 1, [&] { fft_rowst_kernel<N2, -1>(ra); });
    else emu::launch(rgrid, rowst_cta_threads(N2), rowst_smem_bytes(N2), 1, 1, [&] { fft_rowst_kernel<N2, 1>(ra); });
    return 0;
}

// Whole convolution of `batch` blocks in one kernel (conv_kernel.cuh).
extern "C" int EMU_CAT(emu_conv_, EMU_CONV_N)(int mono, int half, void *out, const void *in, const void *filter, const void *tw_fwd, const void *tw_inv,
                                              const void *rtw_fwd, const void *rtw_inv, size_t out_stride, size_t in_stride, size_t batch, float norm)
{
    constexpr int N = EMU_CONV_N;
    mcu_conv_args a;
    memset(&a, 0, sizeof(a));
    a.out = out; a.in = in; a.filter = filter; a.tw_fwd = tw_fwd; a.tw_inv = tw_inv; a.rtw_fwd = rtw_fwd; a.rtw_inv = rtw_inv;
    a.out_stride = out_stride; a.in_stride = in_stride; a.batch = batch; a.n = N; a.in_mode = half ? MCU_IN_HALF : MCU_IN_COMPLEX;
    a.mono = mono; a.norm = norm;
    constexpr int TPC = rows_per_cta(N, 0);
    const emu::idx3 grid{ (unsigned)((batch + TPC - 1) / TPC), 1, 1 };
    const unsigned threads = rows_cta_threads(N, 0);
    const size_t smem = rows_smem_bytes(N, 0);
    if (mono) {
        if (half) emu::launch(grid, threads, smem, 1, 1, [&] { fft_conv_fused_kernel<N, MCU_IN_HALF, true>(a); });
        else emu::launch(grid, threads, smem, 1, 1, [&] { fft_conv_fused_kernel<N, MCU_IN_COMPLEX, true>(a); });
    } else {
        if (half) emu::launch(grid, threads, smem, 1, 1, [&] { fft_conv_fused_kernel<N, MCU_IN_HALF, false>(a); });
        else emu::launch(grid, threads, smem, 1, 1, [&] { fft_conv_fused_kernel<N, MCU_IN_COMPLEX, false>(a); });
    }
    return 0;
}
