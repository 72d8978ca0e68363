// TEST INFRASTRUCTURE ONLY: the thread-block-cluster row kernel of mufft_b200/csrc/rowsc_kernel.cuh
// compiled for the CPU emulation at one length (-DEMU_N=<32768|65536>): the CTAs of a cluster run
// concurrently, remote shared-memory stores go to the peer CTA's emulated shared memory, cluster
// barriers are real barriers over all their threads.
#include "rowsc_kernel.cuh"

using namespace mufft;
#define EMU_CAT_(a, b) a##b
#define EMU_CAT(a, b) EMU_CAT_(a, b)

namespace {
template <int DIR, int IN_MODE, int OUT_MODE>
int run(const mcu_rows_args &a, unsigned clusters)
{
    constexpr int CL = rowsc_cluster(EMU_N);
    if (a.prefetch) // st.async + mbarrier flavour
        emu::launch(emu::idx3{ clusters * CL, 1, 1 }, MUFFT_ROWSC_THREADS, rowsc_smem_bytes(), CL, CL,
                    [&] { fft_rows_cluster_kernel<EMU_N, DIR, IN_MODE, OUT_MODE, true>(a); });
    else
        emu::launch(emu::idx3{ clusters * CL, 1, 1 }, MUFFT_ROWSC_THREADS, rowsc_smem_bytes(), CL, CL,
                    [&] { fft_rows_cluster_kernel<EMU_N, DIR, IN_MODE, OUT_MODE, false>(a); });
    return 0;
}
} // namespace

// out_mode: 0 = complex rows, 3 = scatter over `peers` slabs (bases/shift/row0/pitch as mcu_peer_dest)
extern "C" int EMU_CAT(emu_rows_cluster_, EMU_N)(int direction, int in_mode, int out_mode, void *out, const void *in, const void *tw,
                                                 size_t out_stride, size_t in_stride, size_t batch, unsigned clusters, int async_flavour,
                                                 void *const *bases, unsigned peers, unsigned shift, size_t row0, size_t pitch)
{
    mcu_rows_args a;
    memset(&a, 0, sizeof(a));
    a.out = out; a.in = in; a.tw = tw; a.out_stride = out_stride; a.in_stride = in_stride; a.batch = batch;
    a.n = EMU_N; a.direction = direction; a.in_mode = in_mode; a.out_mode = out_mode; a.norm = 1.0f; a.prefetch = async_flavour;
    for (unsigned p = 0; p < peers && p < MCU_MAX_PEERS; p++) a.peers.base[p] = bases[p];
    a.peers.shift = shift; a.peers.row0 = row0; a.peers.pitch = pitch;
    if (in_mode == MCU_IN_COMPLEX && out_mode == MCU_OUT_COMPLEX)
        return direction < 0 ? run<-1, MCU_IN_COMPLEX, MCU_OUT_COMPLEX>(a, clusters) : run<1, MCU_IN_COMPLEX, MCU_OUT_COMPLEX>(a, clusters);
    if (in_mode == MCU_IN_HALF && out_mode == MCU_OUT_COMPLEX)
        return direction < 0 ? run<-1, MCU_IN_HALF, MCU_OUT_COMPLEX>(a, clusters) : run<1, MCU_IN_HALF, MCU_OUT_COMPLEX>(a, clusters);
    if (in_mode == MCU_IN_COMPLEX && out_mode == MCU_OUT_PEERS)
        return direction < 0 ? run<-1, MCU_IN_COMPLEX, MCU_OUT_PEERS>(a, clusters) : run<1, MCU_IN_COMPLEX, MCU_OUT_PEERS>(a, clusters);
    return -1;
}

extern "C" int EMU_CAT(emu_launch_rows_cluster_, EMU_N)(const mcu_rows_args *a)
{
    const unsigned clusters = a->batch < 2 ? (unsigned)a->batch : 2u;
    if (clusters == 0) return 0;
    if (a->in_mode == MCU_IN_COMPLEX && a->out_mode == MCU_OUT_COMPLEX)
        return a->direction < 0 ? run<-1, MCU_IN_COMPLEX, MCU_OUT_COMPLEX>(*a, clusters) : run<1, MCU_IN_COMPLEX, MCU_OUT_COMPLEX>(*a, clusters);
    if (a->in_mode == MCU_IN_HALF && a->out_mode == MCU_OUT_COMPLEX)
        return a->direction < 0 ? run<-1, MCU_IN_HALF, MCU_OUT_COMPLEX>(*a, clusters) : run<1, MCU_IN_HALF, MCU_OUT_COMPLEX>(*a, clusters);
    if (a->in_mode == MCU_IN_COMPLEX && a->out_mode == MCU_OUT_PEERS)
        return a->direction < 0 ? run<-1, MCU_IN_COMPLEX, MCU_OUT_PEERS>(*a, clusters) : run<1, MCU_IN_COMPLEX, MCU_OUT_PEERS>(*a, clusters);
    return 1;
}
