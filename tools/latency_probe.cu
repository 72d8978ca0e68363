// Where do the microseconds of one small synchronous call go?  (BASELINE configs[0]: `muFFT-bench 1000000 64` runs
// ONE 64-point transform per mufft_execute_plan_1d call: 0.077 us on a host core, ~10.6 us through a kernel launch.)
// Measurement tool, not part of the library: raw CUDA launch + completion variants next to the library call.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -Iinclude -o tools/_build/latency_probe tools/latency_probe.cu -Lmufft_b200 -lmufft_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../mufft_b200'
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

#include "fft.h"
#include "mufft_cuda.h"

__global__ void null_kernel() {}
__global__ void flag_kernel(volatile unsigned *flag, unsigned value) { *flag = value; }
__global__ void work_flag_kernel(float2 *out, const float2 *in, volatile unsigned *flag, unsigned value)
{
    out[threadIdx.x] = make_float2(in[threadIdx.x].x + 1.0f, in[threadIdx.x].y);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *flag = value;
}

template <typename F>
static double us_per_call(int iters, F &&f)
{
    for (int i = 0; i < 200; i++) f(i);
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; i++) f(200 + i);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::micro>(t1 - t0).count() / iters;
}

int main(int argc, char **argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    if (argc > 2 && !strcmp(argv[2], "spin")) cudaSetDeviceFlags(cudaDeviceScheduleSpin);
    cudaFree(0);
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    volatile unsigned *flag = nullptr;
    cudaHostAlloc((void **)&flag, 64, cudaHostAllocMapped);
    *flag = 0;
    float2 *din, *dout;
    cudaMalloc(&din, 64 * 8);
    cudaMalloc(&dout, 64 * 8);
    cudaMemset(din, 0, 64 * 8);

    printf("{\"what\": \"launch+cudaStreamSynchronize (legacy stream)\", \"us\": %.3f}\n", us_per_call(iters, [&](int) { null_kernel<<<1, 32>>>(); cudaStreamSynchronize(0); }));
    printf("{\"what\": \"launch+cudaStreamSynchronize (own stream)\", \"us\": %.3f}\n", us_per_call(iters, [&](int) { null_kernel<<<1, 32, 0, s>>>(); cudaStreamSynchronize(s); }));
    printf("{\"what\": \"launch+spin cudaStreamQuery\", \"us\": %.3f}\n", us_per_call(iters, [&](int) { null_kernel<<<1, 32, 0, s>>>(); while (cudaStreamQuery(s) == cudaErrorNotReady) {} }));
    printf("{\"what\": \"launch+spin on a flag in mapped host memory\", \"us\": %.3f}\n",
           us_per_call(iters, [&](int i) { flag_kernel<<<1, 32, 0, s>>>(flag, (unsigned)i + 1); while (*flag != (unsigned)i + 1) {} }));
    printf("{\"what\": \"64-point copy kernel + flag in mapped host memory\", \"us\": %.3f}\n",
           us_per_call(iters, [&](int i) { work_flag_kernel<<<1, 64, 0, s>>>(dout, din, flag, (unsigned)i + 1); while (*flag != (unsigned)i + 1) {} }));
    printf("{\"what\": \"launch only (no wait; amortised)\", \"us\": %.3f}\n", us_per_call(iters, [&](int) { null_kernel<<<1, 32, 0, s>>>(); }));
    cudaStreamSynchronize(s);
    {
        cudaPointerAttributes a;
        printf("{\"what\": \"cudaPointerGetAttributes (device pointer)\", \"us\": %.3f}\n", us_per_call(iters, [&](int) { cudaPointerGetAttributes(&a, din); }));
        char host[64];
        printf("{\"what\": \"cudaPointerGetAttributes (plain host pointer)\", \"us\": %.3f}\n", us_per_call(iters, [&](int) { cudaPointerGetAttributes(&a, host); cudaGetLastError(); }));
    }
    {   // CUDA graph with one kernel node
        cudaGraph_t g;
        cudaGraphExec_t ge;
        cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
        null_kernel<<<1, 32, 0, s>>>();
        cudaStreamEndCapture(s, &g);
        cudaGraphInstantiate(&ge, g, 0);
        printf("{\"what\": \"graph launch (1 kernel node)+cudaStreamSynchronize\", \"us\": %.3f}\n", us_per_call(iters, [&](int) { cudaGraphLaunch(ge, s); cudaStreamSynchronize(s); }));
    }
    // the library call as the reference's bench makes it: managed memory (mufft_alloc), then device memory
    mufft_plan_1d *plan = mufft_create_plan_1d_c2c(64, MUFFT_FORWARD, 0);
    void *min = mufft_calloc(64 * 8), *mout = mufft_alloc(64 * 8);
    printf("{\"what\": \"mufft_execute_plan_1d N=64, mufft_alloc memory\", \"us\": %.3f}\n", us_per_call(iters, [&](int) { mufft_execute_plan_1d(plan, mout, min); }));
    printf("{\"what\": \"mufft_execute_plan_1d N=64, device memory\", \"us\": %.3f}\n", us_per_call(iters, [&](int) { mufft_execute_plan_1d(plan, dout, din); }));
    printf("{\"what\": \"mufft_execute_plan_1d_batch_async N=64 x 1, device memory (no wait; amortised)\", \"us\": %.3f}\n",
           us_per_call(iters, [&](int) { mufft_execute_plan_1d_batch_async(plan, dout, din, 1, 0, 0); }));
    mufft_cuda_synchronize();
    mufft_free_plan_1d(plan);
    mufft_free(min);
    mufft_free(mout);
    return 0;
}
