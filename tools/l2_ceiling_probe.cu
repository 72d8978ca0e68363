// What a one-pass kernel with an L2-resident intermediate can reach on this chip, measured without any FFT:
//   trips = 0   out = in                                       (the copy behind the HBM roofline)
//   trips = k   in -> (CTA-private 64 KB scratch slot -> back) x k -> out: the slot is rewritten all the time and
//               stays in L2 (stored evict_last, read back .cg), DRAM sees the input once and the output once
// Reported: the copy, one and two trips, the cost per trip (2 D bytes written to and 2 D read from L2).  DESIGN.md section
// 3.7, profiles/README.md R2.12.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/l2_ceiling_probe tools/l2_ceiling_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int THREADS = 256;

__device__ __forceinline__ float4 ld_stream(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float4 *p, float4 v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_keep(float4 *p, float4 v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 ld_l2(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int TRIPS, int TILE_F4, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) probe(const float4 *in, float4 *out, float4 *scratch, size_t tiles)
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    float4 *slot = scratch + (size_t)blockIdx.x * TILE_F4;
    constexpr int PER = TILE_F4 / THREADS; // 16 float4 per thread
    for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        float4 v[PER];
#pragma unroll
        for (int i = 0; i < PER; i++) v[i] = ld_stream(in + tile * TILE_F4 + i * THREADS + threadIdx.x);
#pragma unroll 1
        for (int k = 0; k < TRIPS; k++) {
#pragma unroll
            for (int i = 0; i < PER; i++) st_keep(slot + i * THREADS + threadIdx.x, v[i], pol);
            __syncthreads();
            // read what ANOTHER thread wrote: the data really goes through L2
#pragma unroll
            for (int i = 0; i < PER; i++) v[i] = ld_l2(slot + i * THREADS + (threadIdx.x ^ 32));
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < PER; i++) st_stream(out + tile * TILE_F4 + i * THREADS + threadIdx.x, v[i]);
    }
}

template <int TRIPS, int TILE_F4, int CTAS>
static float run(const float4 *in, float4 *out, float4 *scratch, size_t tiles, int grid)
{
    cudaEvent_t a, b;
    CHECK(cudaEventCreate(&a));
    CHECK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 7; rep++) {
        CHECK(cudaEventRecord(a));
        probe<TRIPS, TILE_F4, CTAS><<<grid, THREADS>>>(in, out, scratch, tiles);
        CHECK(cudaEventRecord(b));
        CHECK(cudaEventSynchronize(b));
        float ms;
        CHECK(cudaEventElapsedTime(&ms, a, b));
        if (rep >= 2 && ms < best) best = ms;
    }
    return best;
}

template <int TILE_F4, int CTAS>
static void config(const float4 *in, float4 *out, float4 *scratch, size_t bytes, int sms)
{
    const size_t tiles = bytes / (TILE_F4 * sizeof(float4));
    const int grid = sms * CTAS;
    const float t0 = run<0, TILE_F4, CTAS>(in, out, scratch, tiles, grid), t1 = run<1, TILE_F4, CTAS>(in, out, scratch, tiles, grid),
                t2 = run<2, TILE_F4, CTAS>(in, out, scratch, tiles, grid);
    const double gb = bytes * 1e-9;
    printf("{\"tool\": \"l2_ceiling_probe\", \"tile_kb\": %d, \"ctas_per_sm\": %d, \"scratch_mb\": %.1f, \"bytes_each_way\": %zu, \"copy_ms\": %.4f, \"copy_dram_gbs\": %.1f, "
           "\"one_trip_ms\": %.4f, \"two_trips_ms\": %.4f, \"ms_per_trip\": %.4f, \"l2_sm_side_gbs_per_trip\": %.0f, \"one_trip_frac_of_copy\": %.3f}\n",
           TILE_F4 * 16 / 1024, CTAS, grid * TILE_F4 * sizeof(float4) / 1048576.0, bytes, t0, 2 * gb / (t0 * 1e-3), t1, t2, (t2 - t0) / 2, 2 * gb / ((t2 - t0) / 2 * 1e-3), t0 / t1);
}

int main()
{
    const size_t bytes = (size_t)2 << 30; // 2 GiB in, 2 GiB out: the traffic of 256 images of 1024 x 1024
    int sms = 0;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float4 *in, *out, *scratch;
    CHECK(cudaMalloc(&in, bytes));
    CHECK(cudaMalloc(&out, bytes));
    CHECK(cudaMalloc(&scratch, (size_t)256 << 20));
    CHECK(cudaMemset(in, 1, bytes));
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    size_t lim = 0;
    CHECK(cudaDeviceGetLimit(&lim, cudaLimitPersistingL2CacheSize));
    printf("{\"l2_bytes\": %d, \"persisting_l2_max\": %d, \"persisting_l2_limit_now\": %zu}\n", prop.l2CacheSize, prop.persistingL2CacheMaxSize, lim);
    const size_t limits_mb[] = { 0 /* as found */, 16, 32, 40, 48, 64 };
    for (size_t mb : limits_mb) {
        if (mb != 0) CHECK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, mb << 20));
        CHECK(cudaDeviceGetLimit(&lim, cudaLimitPersistingL2CacheSize));
        printf("{\"persisting_l2_limit_now\": %zu}\n", lim);
        config<2048, 4>(in, out, scratch, bytes, sms); // 32 KB tiles, 18.5 MB of slots
        config<2048, 6>(in, out, scratch, bytes, sms); // 27.8 MB
        config<2048, 8>(in, out, scratch, bytes, sms); // 37 MB
        config<4096, 8>(in, out, scratch, bytes, sms); // 64 KB tiles, 74 MB
    }
    return 0;
}
