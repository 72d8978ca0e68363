// Compiled with -DROWSC_N=<32768|65536>; exposes rowsc_launch_<n>(args, stream) to dispatch.cu.
#include "rowsc_kernel.cuh"

#ifndef ROWSC_N
#error "compile with -DROWSC_N=<32768|65536>"
#endif
#define MUFFT_CAT_(a, b) a##b
#define MUFFT_CAT(a, b) MUFFT_CAT_(a, b)

namespace mufft {

template <int DIR, int IN_MODE, int OUT_MODE, bool ASYNC>
static cudaError_t launch(const mcu_rows_args &a, cudaStream_t stream)
{
    constexpr int N = ROWSC_N;
    constexpr int CL = rowsc_cluster(N);
    static_assert(rowsc_supported(N), "no cluster shape for this length");
    static int clusters = 0; // per instantiation
    auto kernel = fft_rows_cluster_kernel<N, DIR, IN_MODE, OUT_MODE, ASYNC>;
    if (clusters == 0) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rowsc_smem_bytes());
        if (err != cudaSuccess) return err;
        int dev = 0, sms = 0;
        if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
        if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
        clusters = sms / CL > 0 ? sms / CL : 1; // one 512-thread, 128-register CTA fills an SM
    }
    if (a.batch == 0) return cudaSuccess;
    const unsigned n_clusters = a.batch < (size_t)clusters ? (unsigned)a.batch : (unsigned)clusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_clusters * CL, 1, 1);
    cfg.blockDim = dim3(MUFFT_ROWSC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = rowsc_smem_bytes();
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, a);
}

template <int DIR>
static cudaError_t launch_dir(const mcu_rows_args &a, cudaStream_t stream)
{
    switch (a.in_mode * 8 + a.out_mode) {
    // args.prefetch doubles as the flavour switch here: 1 = st.async + mbarrier exchange
#define MUFFT_ROWSC_CASE(im, om) case im * 8 + om: return a.prefetch ? launch<DIR, im, om, true>(a, stream) : launch<DIR, im, om, false>(a, stream);
        MUFFT_ROWSC_CASE(MCU_IN_COMPLEX, MCU_OUT_COMPLEX)
        MUFFT_ROWSC_CASE(MCU_IN_COMPLEX, MCU_OUT_PEERS)
        MUFFT_ROWSC_CASE(MCU_IN_HALF, MCU_OUT_COMPLEX)
#undef MUFFT_ROWSC_CASE
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t MUFFT_CAT(rowsc_launch_, ROWSC_N)(const mcu_rows_args &a, cudaStream_t stream)
{
    return a.direction < 0 ? launch_dir<-1>(a, stream) : launch_dir<1>(a, stream);
}

} // namespace mufft
