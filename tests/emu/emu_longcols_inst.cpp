// TEST INFRASTRUCTURE ONLY: the single-launch four-step column kernel of mufft_b200/csrc/longcols_kernel.cuh on the CPU
// emulation; `ctas` CTAs run concurrently (the ticket / ring protocol between CTAs is exercised for real).
#include "longcols_kernel.cuh"

using namespace mufft;

template <int N1>
static int run_n1(const mcu_longcols_args &a, unsigned ctas)
{
    const emu::idx3 grid{ ctas, 1, 1 };
    if (a.direction < 0) emu::launch(grid, LONGCOLS_THREADS, longcols_smem_bytes(), ctas, 1, [&] { fft_longcols_kernel<N1, -1>(a); });
    else emu::launch(grid, LONGCOLS_THREADS, longcols_smem_bytes(), ctas, 1, [&] { fft_longcols_kernel<N1, 1>(a); });
    return 0;
}

extern "C" int emu_launch_longcols(const mcu_longcols_args *a, unsigned ctas)
{
    if (a->width == 0) return 0;
    mcu_longcols_args b = *a;
    if (b.lag == 0) b.lag = 1;
    const unsigned blocks = (b.width + LONGCOLS_CB - 1) / LONGCOLS_CB;
    const unsigned lag_eff = b.lag < blocks ? b.lag : blocks;
    if (b.scratch == nullptr || b.ring == 0 || !(b.ring > lag_eff || b.ring >= blocks)) return 1;
    switch (b.n1) {
    case 4: return run_n1<4>(b, ctas);
    case 8: return run_n1<8>(b, ctas);
    case 16: return run_n1<16>(b, ctas);
    case 32: return run_n1<32>(b, ctas);
    default: return 1;
    }
}
