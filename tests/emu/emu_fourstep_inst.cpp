// TEST INFRASTRUCTURE ONLY: the single-launch four-step kernel of mufft_b200/csrc/long_kernel.cuh compiled for
// the CPU emulation at one shape (-DEMU_N1=<n1> -DEMU_N2=<n2>).  `ctas` persistent CTAs, each with its own
// scratch slot, walk the batch.
#include "long_kernel.cuh"

using namespace mufft;
#define EMU_CAT_(a, b, c, d) a##b##c##d
#define EMU_CAT(a, b, c, d) EMU_CAT_(a, b, c, d)

extern "C" int EMU_CAT(emu_long_slots_per_cta_, EMU_N1, x, EMU_N2)() { return 1; }
extern "C" int EMU_CAT(emu_launch_long_, EMU_N1, x, EMU_N2)(const mcu_long_args *a, unsigned ctas)
{
    constexpr int N1 = EMU_N1, N2 = EMU_N2;
    if (a->batch == 0) return 0;
    if (a->n1 != N1 || a->n2 != N2 || a->n != (unsigned)(N1 * N2)) return 1;
    constexpr int SLOTS = 1;
    size_t grid = a->batch < ctas ? a->batch : ctas;
    if (a->scratch_slots / SLOTS < grid) grid = a->scratch_slots / SLOTS;
    if (grid == 0) return 1;
    const emu::idx3 g{ (unsigned)grid, 1, 1 };
    const bool peers = a->to_peers != 0;
    const unsigned threads = long_cta_threads(N2, peers);
    const size_t smem = long_smem_bytes(N2, peers);
    const bool fwd = a->direction < 0;
#define fft_long_kernel_for(...) fft_long_kernel<__VA_ARGS__>(*a)
    // CTAs are independent of each other: run them one after another
    if (a->to_peers) {
        if (a->in_mode != MCU_IN_COMPLEX) return 1;
        emu::launch(g, threads, smem, 1, 1, [&] { if (fwd) fft_long_kernel_for(N1, N2, -1, MCU_IN_COMPLEX, true); else fft_long_kernel_for(N1, N2, 1, MCU_IN_COMPLEX, true); });
    } else if (a->in_mode == MCU_IN_HALF) {
        emu::launch(g, threads, smem, 1, 1, [&] { if (fwd) fft_long_kernel_for(N1, N2, -1, MCU_IN_HALF, false); else fft_long_kernel_for(N1, N2, 1, MCU_IN_HALF, false); });
    } else if (a->in_mode == MCU_IN_COMPLEX) {
        emu::launch(g, threads, smem, 1, 1, [&] { if (fwd) fft_long_kernel_for(N1, N2, -1, MCU_IN_COMPLEX, false); else fft_long_kernel_for(N1, N2, 1, MCU_IN_COMPLEX, false); });
    } else {
        return 1;
    }
    return 0;
}
