// TEST INFRASTRUCTURE ONLY: the kernels that move data with the TMA unit - the warp-specialised 2D kernel
// (fft2d_tma_kernel.cuh), the persistent prefetching row kernel (rows_prefetch_kernel.cuh) and the four-step kernel with
// its step-A inputs through the exchange area (long_kernel.cuh, TMA_A) - compiled for the CPU emulation, where bulk copies
// happen at once and mbarriers are 8-byte words updated by compare-and-swap (tma_util.cuh, MUFFT_EMU branch).  What this
// checks: index maps, buffer layouts, the mbarrier / ticket protocols with OS threads as erratic "warps" - the parity
// mistake of the first GPU version (one `full` barrier per buffer, waited on by two groups) fails here.
// -DEMU_NX=<nx> -DEMU_NY=<ny> (2D shape), -DEMU_ROWS_N=<n> (row length), -DEMU_N1 / -DEMU_N2 (four-step shape).
#include "fft2d_tma_kernel.cuh"
#include "rows_prefetch_kernel.cuh"
#include "long_kernel.cuh"

using namespace mufft;
#define EMU_CAT_(a, b, c, d) a##b##c##d
#define EMU_CAT(a, b, c, d) EMU_CAT_(a, b, c, d)
#define EMU_CAT2_(a, b) a##b
#define EMU_CAT2(a, b) EMU_CAT2_(a, b)

// ---- 2D: C2C both directions, R2C where the kernel takes the kind ------------------------------------------------
template <int DIR, int KIND>
static int run_2d(const mcu_2d_args &a, unsigned ctas)
{
    constexpr int NX = EMU_NX, NY = EMU_NY;
    if constexpr (!f2d_tma_kind_supported(NX, NY, KIND)) {
        (void)a; (void)ctas;
        return -3;
    } else {
        f2d_tma_params p;
        memset(&p, 0, sizeof(p));
        p.a = a;
        p.map.base = a.out;
        p.map.dims[0] = 2ull * f2d_width(NX, KIND); // the columns that carry data, as fft2d_inst.cu encodes them
        p.map.dims[1] = (unsigned long long)a.batch * NY;
        p.map.row_bytes = (unsigned long long)f2d_pitch(NX, KIND) * sizeof(cpx);
        p.map.box[0] = 2u * f2d_tma_cols_per_tile(NY);
        p.map.box[1] = F2D_TMA_BOX_ROWS;
        const emu::idx3 grid{ ctas, 1, 1 };
        emu::launch(grid, f2d_tma_threads(NX, NY, KIND), f2d_tma_smem_bytes(NX, NY, KIND), ctas, 1, [&] { fft2d_tma_kernel<NX, NY, DIR, MUFFT_TW_SMEM | MUFFT_TW_DERIVE, KIND>(p); });
        return 0;
    }
}
extern "C" int EMU_CAT(emu_launch_2d_tma_, EMU_NX, x, EMU_NY)(const mcu_2d_args *a, unsigned ctas)
{
    if (a->batch == 0) return 0;
    mcu_2d_args b = *a;
    if (b.lag == 0) b.lag = 1;
    switch (b.kind) {
    case MCU_2D_C2C: return b.direction < 0 ? run_2d<-1, F2D_C2C>(b, ctas) : run_2d<1, F2D_C2C>(b, ctas);
    case MCU_2D_R2C: return run_2d<-1, F2D_R2C>(b, ctas);
    default: return 1;
    }
}
extern "C" int EMU_CAT(emu_2d_tma_stages_, EMU_NX, x, EMU_NY)(int kind) { return f2d_tma_stages(EMU_NX, EMU_NY, kind); }
extern "C" unsigned EMU_CAT(emu_2d_tma_row_tiles_, EMU_NX, x, EMU_NY)() { return f2d_tma_row_tiles(EMU_NX, EMU_NY, F2D_C2C); }

// ---- rows: every mode pair the prefetching kernel takes at this length -------------------------------------------
template <int DIR, int IN_MODE, int OUT_MODE>
static int run_rows(const mcu_rows_args &a, unsigned ctas)
{
    constexpr int N = EMU_ROWS_N;
    if constexpr (!rows_prefetch_supported(N, IN_MODE, OUT_MODE)) {
        (void)a; (void)ctas;
        return -3;
    } else {
        constexpr int V = rows_variant(N, IN_MODE, OUT_MODE);
        const emu::idx3 grid{ ctas, 1, 1 };
        emu::launch(grid, rows_cta_threads(N, V), rows_prefetch_smem_bytes(N, V), 1, 1, [&] { fft_rows_prefetch_kernel<N, DIR, IN_MODE, OUT_MODE>(a); });
        return 0;
    }
}
extern "C" int EMU_CAT2(emu_launch_rows_prefetch_, EMU_ROWS_N)(const mcu_rows_args *a, unsigned ctas)
{
    if (a->batch == 0) return 0;
    const bool fwd = a->direction < 0;
    switch (a->in_mode * 8 + a->out_mode) {
    case MCU_IN_COMPLEX * 8 + MCU_OUT_COMPLEX: return fwd ? run_rows<-1, MCU_IN_COMPLEX, MCU_OUT_COMPLEX>(*a, ctas) : run_rows<1, MCU_IN_COMPLEX, MCU_OUT_COMPLEX>(*a, ctas);
    case MCU_IN_HALF * 8 + MCU_OUT_COMPLEX: return fwd ? run_rows<-1, MCU_IN_HALF, MCU_OUT_COMPLEX>(*a, ctas) : run_rows<1, MCU_IN_HALF, MCU_OUT_COMPLEX>(*a, ctas);
    case MCU_IN_MUL * 8 + MCU_OUT_COMPLEX: return fwd ? -3 : run_rows<1, MCU_IN_MUL, MCU_OUT_COMPLEX>(*a, ctas);
    case MCU_IN_C2R * 8 + MCU_OUT_COMPLEX: return fwd ? -3 : run_rows<1, MCU_IN_C2R, MCU_OUT_COMPLEX>(*a, ctas);
    case MCU_IN_COMPLEX * 8 + MCU_OUT_R2C: return fwd ? run_rows<-1, MCU_IN_COMPLEX, MCU_OUT_R2C>(*a, ctas) : -3;
    case MCU_IN_COMPLEX * 8 + MCU_OUT_R2C_FULL: return fwd ? run_rows<-1, MCU_IN_COMPLEX, MCU_OUT_R2C_FULL>(*a, ctas) : -3;
    default: return 1;
    }
}

// ---- four-step rows with step A fed through the exchange area -------------------------------------------------------
extern "C" int EMU_CAT(emu_launch_long_tma_, EMU_N1, x, EMU_N2)(const mcu_long_args *a, unsigned ctas)
{
    constexpr int N1 = EMU_N1, N2 = EMU_N2;
    if (a->batch == 0) return 0;
    if (a->n1 != N1 || a->n2 != N2 || a->n != (unsigned)(N1 * N2) || a->to_peers) return 1;
    size_t grid = a->batch < ctas ? a->batch : ctas;
    if (a->scratch_slots < grid) grid = a->scratch_slots;
    if (grid == 0) return 1;
    const emu::idx3 g{ (unsigned)grid, 1, 1 };
    const unsigned threads = long_cta_threads(N2, false);
    const size_t smem = long_smem_bytes(N2, false) + 32;
    const bool fwd = a->direction < 0;
    if (a->in_mode == MCU_IN_HALF)
        emu::launch(g, threads, smem, 1, 1, [&] { if (fwd) fft_long_kernel<N1, N2, -1, MCU_IN_HALF, false, true>(*a); else fft_long_kernel<N1, N2, 1, MCU_IN_HALF, false, true>(*a); });
    else if (a->in_mode == MCU_IN_COMPLEX)
        emu::launch(g, threads, smem, 1, 1, [&] { if (fwd) fft_long_kernel<N1, N2, -1, MCU_IN_COMPLEX, false, true>(*a); else fft_long_kernel<N1, N2, 1, MCU_IN_COMPLEX, false, true>(*a); });
    else
        return 1;
    return 0;
}
