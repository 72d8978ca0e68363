// TEST INFRASTRUCTURE ONLY: the row kernel template of mufft_b200/csrc/rows_kernel.cuh compiled for
// the CPU emulation (cuda_emu.h) at one length (-DEMU_N=<n>), behind a flat C entry point for
// tests/test_emu.py.  Twiddle tables come from the real library's host code
// (mcu_engine_table_fill), so the kernel sees here exactly the tables it sees on the GPU.
#include "rows_kernel.cuh"

using namespace mufft;
#define EMU_CAT_(a, b) a##b
#define EMU_CAT(a, b) EMU_CAT_(a, b)

namespace {
template <int DIR, int IN_MODE, int OUT_MODE>
int run(const mcu_rows_args &a)
{
    constexpr int N = EMU_N;
    constexpr int V = rows_variant(N, IN_MODE, OUT_MODE);
    constexpr int TPC = rows_per_cta(N, V);
    const size_t groups = (a.batch + TPC - 1) / TPC;
    emu::launch(emu::idx3{ (unsigned)groups, 1, 1 }, rows_cta_threads(N, V), rows_smem_bytes(N, V), 1, 1,
                [&] { fft_rows_kernel<N, DIR, IN_MODE, OUT_MODE>(a); });
    return 0;
}
} // namespace

static int dispatch(const mcu_rows_args &a)
{
    if (a.direction < 0) {
        switch (a.in_mode * 8 + a.out_mode) {
        case MCU_IN_COMPLEX * 8 + MCU_OUT_COMPLEX: return run<-1, MCU_IN_COMPLEX, MCU_OUT_COMPLEX>(a);
        case MCU_IN_COMPLEX * 8 + MCU_OUT_PEERS: return run<-1, MCU_IN_COMPLEX, MCU_OUT_PEERS>(a);
        case MCU_IN_COMPLEX * 8 + MCU_OUT_R2C: return run<-1, MCU_IN_COMPLEX, MCU_OUT_R2C>(a);
        case MCU_IN_COMPLEX * 8 + MCU_OUT_R2C_FULL: return run<-1, MCU_IN_COMPLEX, MCU_OUT_R2C_FULL>(a);
        case MCU_IN_COMPLEX * 8 + MCU_OUT_R2C_PEERS: return run<-1, MCU_IN_COMPLEX, MCU_OUT_R2C_PEERS>(a);
        case MCU_IN_HALF * 8 + MCU_OUT_COMPLEX: return run<-1, MCU_IN_HALF, MCU_OUT_COMPLEX>(a);
        case MCU_IN_HALF * 8 + MCU_OUT_R2C: return run<-1, MCU_IN_HALF, MCU_OUT_R2C>(a);
        case MCU_IN_HALF * 8 + MCU_OUT_R2C_FULL: return run<-1, MCU_IN_HALF, MCU_OUT_R2C_FULL>(a);
        default: return -1;
        }
    }
    if (a.out_mode == MCU_OUT_PEERS && a.in_mode == MCU_IN_COMPLEX) return run<1, MCU_IN_COMPLEX, MCU_OUT_PEERS>(a);
    if (a.out_mode != MCU_OUT_COMPLEX) return -1;
    switch (a.in_mode) {
    case MCU_IN_COMPLEX: return run<1, MCU_IN_COMPLEX, MCU_OUT_COMPLEX>(a);
    case MCU_IN_HALF: return run<1, MCU_IN_HALF, MCU_OUT_COMPLEX>(a);
    case MCU_IN_C2R: return run<1, MCU_IN_C2R, MCU_OUT_COMPLEX>(a);
    case MCU_IN_MUL: return run<1, MCU_IN_MUL, MCU_OUT_COMPLEX>(a);
    case MCU_IN_MUL_C2R: return run<1, MCU_IN_MUL_C2R, MCU_OUT_COMPLEX>(a);
    case MCU_IN_MULACC: return run<1, MCU_IN_MULACC, MCU_OUT_COMPLEX>(a);
    case MCU_IN_MULACC_C2R: return run<1, MCU_IN_MULACC_C2R, MCU_OUT_COMPLEX>(a);
    default: return -1;
    }
}

// struct-taking entry point for the emulated CUDA side (emu_abi.cpp)
extern "C" int EMU_CAT(emu_launch_rows_, EMU_N)(const mcu_rows_args *a) { return dispatch(*a); }

extern "C" int EMU_CAT(emu_rows_, EMU_N)(int direction, int in_mode, int out_mode, void *out, const void *in, const void *in2,
                                         const void *tw, const void *rtw, size_t out_stride, size_t in_stride, size_t in2_stride,
                                         size_t batch, float norm)
{
    mcu_rows_args a;
    memset(&a, 0, sizeof(a));
    a.out = out; a.in = in; a.in2 = in2; a.tw = tw; a.rtw = rtw;
    a.out_stride = out_stride; a.in_stride = in_stride; a.in2_stride = in2_stride;
    a.batch = batch; a.n = EMU_N; a.direction = direction; a.in_mode = in_mode; a.out_mode = out_mode; a.norm = norm;
    return dispatch(a);
}

// Partitioned convolution load stage (MCU_IN_MULACC*): sum over `parts` (spectrum, filter) pairs.
extern "C" int EMU_CAT(emu_rows_parts_, EMU_N)(int in_mode, void *out, const void *in, const void *in2, const void *tw, const void *rtw,
                                               size_t out_stride, size_t in_stride, size_t batch, float norm, unsigned parts,
                                               long long in_part_stride, long long in2_part_stride)
{
    mcu_rows_args a;
    memset(&a, 0, sizeof(a));
    a.out = out; a.in = in; a.in2 = in2; a.tw = tw; a.rtw = rtw;
    a.out_stride = out_stride; a.in_stride = in_stride;
    a.batch = batch; a.n = EMU_N; a.direction = 1; a.in_mode = in_mode; a.out_mode = MCU_OUT_COMPLEX; a.norm = norm;
    a.parts = parts; a.in_part_stride = in_part_stride; a.in2_part_stride = in2_part_stride;
    return dispatch(a);
}

// Exchange-step epilogue (MCU_OUT_PEERS): element x of row b goes to slab x >> shift.
// Same for the R2C epilogue (MCU_OUT_R2C_PEERS): EMU_N + 1 bins per row, bin EMU_N in the last slab.
extern "C" int EMU_CAT(emu_rows_r2c_peers_, EMU_N)(void *const *bases, unsigned peers, unsigned shift, size_t row0, size_t pitch,
                                                   const void *in, const void *tw, const void *rtw, size_t in_stride, size_t batch)
{
    mcu_rows_args a;
    memset(&a, 0, sizeof(a));
    a.in = in; a.tw = tw; a.rtw = rtw; a.in_stride = in_stride; a.batch = batch; a.n = EMU_N; a.direction = -1;
    a.in_mode = MCU_IN_COMPLEX; a.out_mode = MCU_OUT_R2C_PEERS; a.norm = 1.0f;
    for (unsigned p = 0; p < peers && p < MCU_MAX_PEERS; p++) a.peers.base[p] = bases[p];
    a.peers.row0 = row0; a.peers.pitch = pitch; a.peers.pitch_last = pitch + 1; a.peers.shift = shift; a.peers.last = peers - 1;
    return dispatch(a);
}

extern "C" int EMU_CAT(emu_rows_peers_, EMU_N)(int direction, void *const *bases, unsigned peers, unsigned shift, size_t row0, size_t pitch,
                                               const void *in, const void *tw, size_t in_stride, size_t batch)
{
    mcu_rows_args a;
    memset(&a, 0, sizeof(a));
    a.in = in; a.tw = tw; a.in_stride = in_stride; a.batch = batch; a.n = EMU_N; a.direction = direction;
    a.in_mode = MCU_IN_COMPLEX; a.out_mode = MCU_OUT_PEERS; a.norm = 1.0f;
    for (unsigned p = 0; p < peers && p < MCU_MAX_PEERS; p++) a.peers.base[p] = bases[p];
    a.peers.row0 = row0; a.peers.pitch = pitch; a.peers.shift = shift;
    return dispatch(a);
}
