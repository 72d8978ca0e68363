// Per-device launch configuration of a kernel instantiation.  Function attributes (the opt-in to more
// than 48 KB of dynamic shared memory) and occupancy belong to a DEVICE: a process that drives several
// GPUs (mufft_cuda_set_device, slab plans) must configure every kernel once on each of them, and the
// number of resident CTAs of a persistent kernel depends on the device it is asked about.
#pragma once

#include <cuda_runtime.h>

namespace mufft {

#define MCU_MAX_DEVICES 64

struct launch_cfg {
    unsigned char ready[MCU_MAX_DEVICES]; // dynamic shared memory opted in on this device
    int ctas[MCU_MAX_DEVICES];            // resident CTAs of the kernel on the whole device (0 = not asked yet)
};

// Declare `static launch_cfg cfg;` inside the launching function template (one per instantiation) and
// call this before every launch: after the first call on a device it costs one cudaGetDevice.
// resident != nullptr: also returns SMs x CTAs per SM for `threads` threads and `smem` bytes per CTA.
// The races between threads are benign: both sides compute and store the same values.
template <typename Kernel>
static inline cudaError_t launch_prepare(launch_cfg &cfg, Kernel kernel, size_t smem, int threads, int *resident)
{
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    const bool cached = dev >= 0 && dev < MCU_MAX_DEVICES;
    if (!cached || !cfg.ready[dev]) {
        if (smem > 48 * 1024) {
            err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess) return err;
        }
        if (cached) cfg.ready[dev] = 1;
    }
    if (resident != nullptr) {
        if (!cached || cfg.ctas[dev] == 0) {
            int sms = 0, per_sm = 0;
            if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
            if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem)) != cudaSuccess) return err;
            if (per_sm < 1) return cudaErrorLaunchOutOfResources;
            if (cached) cfg.ctas[dev] = sms * per_sm;
            *resident = sms * per_sm;
        } else {
            *resident = cfg.ctas[dev];
        }
    }
    return cudaSuccess;
}

} // namespace mufft
