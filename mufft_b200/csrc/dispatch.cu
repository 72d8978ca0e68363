// C-ABI entry points of the CUDA side (see abi.h): runtime helpers, the size dispatch for the
// row / column engine kernels, and the generic kernels used beyond the engine limits:
//   * stockham_pass_kernel   one radix-8/4/2 Stockham pass, any length, 1D or vertical
//                            (the role of ref kernel.c:84-453 and 455-738)
//   * resolve_kernel         stand-alone R2C / C2R resolve and spectral multiply
//                            (ref kernel.c:22-82); also backs mufft_get_convolve_func().
#include "abi.h"
#include "fft_engine.cuh"
#include "rows_kernel.cuh"

#include <atomic>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace mufft {
#define MUFFT_DECLARE_ROWS(n) cudaError_t rows_launch_##n##_fwd(const mcu_rows_args &, cudaStream_t); cudaError_t rows_launch_##n##_inv(const mcu_rows_args &, cudaStream_t);
#define MUFFT_DECLARE_COLS(n) cudaError_t cols_launch_##n(const mcu_cols_args &, cudaStream_t);
MUFFT_DECLARE_ROWS(2) MUFFT_DECLARE_ROWS(4) MUFFT_DECLARE_ROWS(8) MUFFT_DECLARE_ROWS(16)
MUFFT_DECLARE_ROWS(32) MUFFT_DECLARE_ROWS(64) MUFFT_DECLARE_ROWS(128) MUFFT_DECLARE_ROWS(256)
MUFFT_DECLARE_ROWS(512) MUFFT_DECLARE_ROWS(1024) MUFFT_DECLARE_ROWS(2048) MUFFT_DECLARE_ROWS(4096)
MUFFT_DECLARE_ROWS(8192) MUFFT_DECLARE_ROWS(16384)
#define MUFFT_DECLARE_CONV(n) cudaError_t conv_launch_##n(const mcu_conv_args &, cudaStream_t);
MUFFT_DECLARE_CONV(2) MUFFT_DECLARE_CONV(4) MUFFT_DECLARE_CONV(8) MUFFT_DECLARE_CONV(16)
MUFFT_DECLARE_CONV(32) MUFFT_DECLARE_CONV(64) MUFFT_DECLARE_CONV(128) MUFFT_DECLARE_CONV(256)
MUFFT_DECLARE_CONV(512) MUFFT_DECLARE_CONV(1024) MUFFT_DECLARE_CONV(2048) MUFFT_DECLARE_CONV(4096)
MUFFT_DECLARE_CONV(8192)
MUFFT_DECLARE_COLS(2) MUFFT_DECLARE_COLS(4) MUFFT_DECLARE_COLS(8) MUFFT_DECLARE_COLS(16)
MUFFT_DECLARE_COLS(32) MUFFT_DECLARE_COLS(64) MUFFT_DECLARE_COLS(128) MUFFT_DECLARE_COLS(256)
MUFFT_DECLARE_COLS(512) MUFFT_DECLARE_COLS(1024) MUFFT_DECLARE_COLS(2048)

#define MUFFT_F2D_SIZES(X, nx) X(nx, 256) X(nx, 512) X(nx, 1024) X(nx, 2048)
#define MUFFT_F2D_SHAPES(X) MUFFT_F2D_SIZES(X, 256) MUFFT_F2D_SIZES(X, 512) MUFFT_F2D_SIZES(X, 1024) MUFFT_F2D_SIZES(X, 2048)
#define MUFFT_DECLARE_F2D(pw, ph) cudaError_t f2d_launch_##pw##x##ph(const mcu_2d_args &, cudaStream_t); bool f2d_supported_##pw##x##ph(); bool f2d_real_supported_##pw##x##ph(int);
MUFFT_F2D_SHAPES(MUFFT_DECLARE_F2D)

cudaError_t rowst_launch_256(const mcu_rowst_args &, cudaStream_t);
cudaError_t rowst_launch_512(const mcu_rowst_args &, cudaStream_t);
cudaError_t rowst_launch_1024(const mcu_rowst_args &, cudaStream_t);
cudaError_t rowsc_launch_32768(const mcu_rows_args &, cudaStream_t);
cudaError_t rowsc_launch_65536(const mcu_rows_args &, cudaStream_t);

static std::atomic<unsigned long long> g_launches{0};

// ---- generic Stockham pass ----------------------------------------------------------------
template <int RADIX, int DIR>
__global__ void __launch_bounds__(256) stockham_pass_kernel(const mcu_pass_args a)
{
    const size_t q = a.n / RADIX;
    const size_t total = q * a.width;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const size_t x = idx % a.width;
    const size_t i = idx / a.width;
    const size_t k = i & (size_t)(a.p - 1);
    const size_t base = (i - k) * RADIX + k;
    const size_t step = a.n / ((size_t)RADIX * a.p);
    const cpx *__restrict__ roots = static_cast<const cpx *>(a.roots);

    for (size_t b = blockIdx.y; b < a.batch; b += gridDim.y) {
        const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + b * a.in_batch + x;
        cpx *__restrict__ out = static_cast<cpx *>(a.out) + b * a.out_batch + x;
        cpx v[RADIX];
#pragma unroll
        for (int m = 0; m < RADIX; m++)
            v[m] = (a.half && m >= RADIX / 2) ? make_float2(0.0f, 0.0f) : in[(i + m * q) * a.in_stride];
        if (a.p > 1) {
#pragma unroll
            for (int m = 1; m < RADIX; m++) v[m] = cmul(v[m], __ldg(roots + k * m * step));
        }
        dft_regs<RADIX, DIR>(v);
#pragma unroll
        for (int r = 0; r < RADIX; r++) out[(base + (size_t)r * a.p) * a.out_stride] = v[r];
    }
}

template <int RADIX>
static cudaError_t launch_pass_radix(const mcu_pass_args &a, cudaStream_t stream)
{
    const size_t total = (size_t)(a.n / RADIX) * a.width;
    if (total == 0 || a.batch == 0) return cudaSuccess;
    const size_t ctas = (total + 255) / 256;
    if (ctas > 0x7fffffffull) return cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)ctas, (unsigned)(a.batch > 65535 ? 65535 : a.batch));
    if (a.direction < 0) stockham_pass_kernel<RADIX, -1><<<grid, 256, 0, stream>>>(a);
    else stockham_pass_kernel<RADIX, 1><<<grid, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

// ---- stand-alone resolve / multiply -----------------------------------------------------
__global__ void __launch_bounds__(256) resolve_kernel(const mcu_resolve_args a)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n) return;
    const unsigned n = a.n;
    const cpx *__restrict__ rtw = static_cast<const cpx *>(a.rtw);
    for (size_t b = blockIdx.y; b < a.batch; b += gridDim.y) {
        const cpx *__restrict__ in = static_cast<const cpx *>(a.in) + b * a.in_stride;
        const cpx *__restrict__ in2 = a.in2 ? static_cast<const cpx *>(a.in2) + b * a.in2_stride : in;
        cpx *__restrict__ out = static_cast<cpx *>(a.out) + b * a.out_stride;
        if (a.mode == MCU_RES_MUL) {
            out[k] = cscale(a.norm, cmul(in[k], in2[k]));
        } else if (a.mode == MCU_RES_C2R || a.mode == MCU_RES_MUL_C2R) {
            cpx xa = in[k], xb = in[n - k];
            if (a.mode == MCU_RES_MUL_C2R) {
                xa = cscale(a.norm, cmul(xa, in2[k]));
                xb = cscale(a.norm, cmul(xb, in2[n - k]));
            }
            xb = cconj(xb);
            out[k] = cadd(cadd(xa, xb), cmul(csub(xa, xb), __ldg(rtw + k)));
        } else {
            const cpx za = in[k];
            if (k == 0) {
                out[0] = make_float2(za.x + za.y, 0.0f);
                out[n] = make_float2(za.x - za.y, 0.0f);
            } else {
                const cpx zb = cconj(in[n - k]);
                const cpx fe = cadd(za, zb);
                const cpx fo = cmul(__ldg(rtw + k), csub(za, zb));
                out[k] = cscale(0.5f, cadd(fe, fo));
                if (a.mode == MCU_RES_R2C_FULL) out[k + n] = cscale(0.5f, csub(fe, fo));
            }
        }
    }
}

// ---- engine twiddle tables -----------------------------------------------------------------
template <int N, int V>
static void fill_table(float *table, int direction)
{
    constexpr int passes = engine_num_passes(N, V);
    for (int j = 1; j < passes; j++) {
        const int r = engine_radix(N, j, V), p = engine_span(N, j, V);
        float *block = table + 2 * (size_t)engine_tw_offset(N, j, V);
        for (int m = 1; m < r; m++)
            for (int k = 0; k < p; k++) {
                // exp(dir * 2*pi*i * k*m / (r*p)), evaluated in double and rounded once, the way
                // the reference builds its tables (ref fft.c:102-106).
                const double phase = (M_PI * direction * 2.0 * (double)(k * m)) / (double)(r * p);
                block[2 * ((size_t)(m - 1) * p + k) + 0] = (float)cos(phase);
                block[2 * ((size_t)(m - 1) * p + k) + 1] = (float)sin(phase);
            }
    }
}

template <int N>
static void fill_table_variant(float *table, int direction, int variant)
{
    if (variant == 1) { if constexpr (engine_variant_exists(N, 1)) fill_table<N, 1>(table, direction); }
    else if (variant == 2) { if constexpr (engine_variant_exists(N, 2)) fill_table<N, 2>(table, direction); }
    else fill_table<N, 0>(table, direction);
}

#define MUFFT_FOR_CONV_SIZES(X) \
    X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192)
// row kernels also exist for MCU_ENGINE_BIG_N
#define MUFFT_FOR_ENGINE_SIZES(X) MUFFT_FOR_CONV_SIZES(X) X(16384)
// lengths with an engine twiddle table: the above and the cluster shapes (rowsc_kernel.cuh)
#define MUFFT_FOR_TABLE_SIZES(X) MUFFT_FOR_ENGINE_SIZES(X) X(32768) X(65536)

} // namespace mufft

using namespace mufft;

extern "C" {

// ---- runtime -----------------------------------------------------------------------------
int mcu_device_count(int *count) { return (int)cudaGetDeviceCount(count); }
int mcu_set_device(int device) { return (int)cudaSetDevice(device); }
int mcu_get_device(int *device) { return (int)cudaGetDevice(device); }
int mcu_sm_count(int *count)
{
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return (int)err;
    return (int)cudaDeviceGetAttribute(count, cudaDevAttrMultiProcessorCount, dev);
}
const char *mcu_error_string(int err) { return cudaGetErrorString((cudaError_t)err); }

int mcu_ipc_export(void *handle64, void *ptr)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    return (int)cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t *>(handle64), ptr);
}
int mcu_ipc_import(void **ptr, const void *handle64)
{
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}
int mcu_ipc_close(void *ptr) { return ptr ? (int)cudaIpcCloseMemHandle(ptr) : 0; }
int mcu_malloc(void **ptr, size_t bytes) { return (int)cudaMalloc(ptr, bytes ? bytes : 256); }
int mcu_free(void *ptr) { return ptr ? (int)cudaFree(ptr) : 0; }
int mcu_malloc_managed(void **ptr, size_t bytes) { return (int)cudaMallocManaged(ptr, bytes ? bytes : 256, cudaMemAttachGlobal); }
int mcu_malloc_pinned(void **ptr, size_t bytes) { return (int)cudaHostAlloc(ptr, bytes ? bytes : 256, cudaHostAllocDefault); }
int mcu_free_pinned(void *ptr) { return ptr ? (int)cudaFreeHost(ptr) : 0; }
int mcu_memset(void *ptr, int value, size_t bytes, void *stream)
{
    return (int)cudaMemsetAsync(ptr, value, bytes, (cudaStream_t)stream);
}

int mcu_pointer_kind(const void *ptr)
{
    cudaPointerAttributes attr;
    cudaError_t err = cudaPointerGetAttributes(&attr, ptr);
    if (err != cudaSuccess) {
        (void)cudaGetLastError();
        return MCU_MEM_HOST;
    }
    switch (attr.type) {
    case cudaMemoryTypeHost: return MCU_MEM_PINNED;
    case cudaMemoryTypeDevice: return MCU_MEM_DEVICE;
    case cudaMemoryTypeManaged: return MCU_MEM_MANAGED;
    default: return MCU_MEM_HOST;
    }
}

int mcu_stream_create(void **stream)
{
    cudaStream_t s;
    cudaError_t err = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    *stream = (void *)s;
    return (int)err;
}
int mcu_stream_destroy(void *stream) { return stream ? (int)cudaStreamDestroy((cudaStream_t)stream) : 0; }
int mcu_stream_sync(void *stream) { return (int)cudaStreamSynchronize((cudaStream_t)stream); }
// Busy-polls the stream instead of blocking in the driver: for work of a few microseconds the
// wake-up of a blocking synchronise costs more than the work.
int mcu_stream_sync_spin(void *stream)
{
    cudaError_t err;
    while ((err = cudaStreamQuery((cudaStream_t)stream)) == cudaErrorNotReady) {
    }
    return (int)err;
}
int mcu_event_create(void **event)
{
    cudaEvent_t e;
    cudaError_t err = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    *event = (void *)e;
    return (int)err;
}
int mcu_event_destroy(void *event) { return event ? (int)cudaEventDestroy((cudaEvent_t)event) : 0; }
int mcu_event_record(void *event, void *stream) { return (int)cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream); }
int mcu_stream_wait_event(void *stream, void *event) { return (int)cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0); }
int mcu_event_sync(void *event) { return (int)cudaEventSynchronize((cudaEvent_t)event); }
int mcu_copy_async(void *dst, const void *src, size_t bytes, void *stream)
{
    if (bytes == 0) return 0;
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
}
int mcu_copy2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t rows, void *stream)
{
    if (width_bytes == 0 || rows == 0) return 0;
    return (int)cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, rows, cudaMemcpyDefault, (cudaStream_t)stream);
}
int mcu_prefetch(const void *ptr, size_t bytes, int to_device, void *stream)
{
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return (int)err;
    err = cudaMemPrefetchAsync(ptr, bytes, to_device ? dev : cudaCpuDeviceId, (cudaStream_t)stream);
    if (err != cudaSuccess) (void)cudaGetLastError(); // advisory only
    return 0;
}

// ---- engine twiddle tables (see fill_table above) --------------------------------------------
int mcu_rows_variant(unsigned n, int in_mode, int out_mode)
{
    // the kernels' own choice (rows_kernel.cuh), so that the table handed to a kernel is always the
    // one of the shape it was compiled for
    switch (n) {
#define X(v) case v: return rows_variant(v, in_mode, out_mode);
        MUFFT_FOR_ENGINE_SIZES(X)
#undef X
    default: return 0;
    }
}

size_t mcu_engine_table_size(unsigned n, int variant)
{
    switch (n) {
#define X(v) case v: return (size_t)(variant == 1 ? engine_tw_total(v, 1) : variant == 2 ? engine_tw_total(v, 2) : engine_tw_total(v, 0)) + 1;
        MUFFT_FOR_TABLE_SIZES(X)
#undef X
    default: return 0;
    }
}

void mcu_engine_table_fill(float *table, unsigned n, int direction, int variant)
{
    switch (n) {
#define X(v) case v: fill_table_variant<v>(table, direction, variant); break;
        MUFFT_FOR_TABLE_SIZES(X)
#undef X
    default: break;
    }
}

// ---- launches -------------------------------------------------------------------------------
int mcu_launch_rows(const mcu_rows_args *a, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    switch (a->n) {
#define X(v) case v: return (int)(a->direction < 0 ? rows_launch_##v##_fwd(*a, s) : rows_launch_##v##_inv(*a, s));
        MUFFT_FOR_ENGINE_SIZES(X)
#undef X
    default: return (int)cudaErrorInvalidValue;
    }
}

int mcu_rows_cluster_supported(unsigned n) { return n == 32768 || n == 65536; }

int mcu_launch_rows_cluster(const mcu_rows_args *a, void *stream)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    switch (a->n) {
    case 32768: return (int)rowsc_launch_32768(*a, (cudaStream_t)stream);
    case 65536: return (int)rowsc_launch_65536(*a, (cudaStream_t)stream);
    default: return (int)cudaErrorInvalidValue;
    }
}

int mcu_launch_conv(const mcu_conv_args *a, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    switch (a->n) {
#define X(v) case v: return (int)conv_launch_##v(*a, s);
        MUFFT_FOR_CONV_SIZES(X)
#undef X
    default: return (int)cudaErrorInvalidValue;
    }
}

int mcu_launch_cols(const mcu_cols_args *a, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    switch (a->n) {
#define X(v) case v: return (int)cols_launch_##v(*a, s);
        X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)
#undef X
    default: return (int)cudaErrorInvalidValue;
    }
}

int mcu_2d_fused_supported(unsigned nx, unsigned ny)
{
#define X(pw, ph) if (nx == pw && ny == ph) return f2d_supported_##pw##x##ph() ? 1 : 0;
    MUFFT_F2D_SHAPES(X)
#undef X
    return 0;
}

int mcu_2d_fused_real_supported(unsigned nx, unsigned ny, int out_mode)
{
#define X(pw, ph) if (nx == pw && ny == ph) return f2d_real_supported_##pw##x##ph(out_mode) ? 1 : 0;
    MUFFT_F2D_SHAPES(X)
#undef X
    return 0;
}

int mcu_launch_2d_fused(const mcu_2d_args *a, void *stream)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
#define X(pw, ph) if (a->nx == pw && a->ny == ph) return (int)f2d_launch_##pw##x##ph(*a, (cudaStream_t)stream);
    MUFFT_F2D_SHAPES(X)
#undef X
    return (int)cudaErrorInvalidValue;
}

// n = n1 * n2 with n1 <= 2048 (column kernel) and n2 in {256, 512, 1024} (row kernel B)
unsigned mcu_split_n2(unsigned n, int columns)
{
    if (n <= MCU_COLS_MAX_N || (n & (n - 1)) != 0) return 0;
    // MUFFT_CUDA_SPLIT_N2=<256|512|1024>: measurement knob, overrides the choice where it is legal
    static const unsigned forced = [] { const char *e = getenv("MUFFT_CUDA_SPLIT_N2"); return e ? (unsigned)atoi(e) : 0u; }();
    if ((forced == 256 || forced == 512 || forced == 1024) && n / forced >= 2 && n / forced <= MCU_COLS_MAX_N) return forced;
    // Row transforms: the widest second factor wins from 32768 points on (measured on B200,
    // profiles/README.md section 3: N = 32768 4.06 -> 3.52 ms, N = 2^20 6.00 -> 4.27 ms with 1024 instead
    // of 256 / 512).  Column transforms keep the narrow split they were measured with.
    if (!columns && n >= 32768u && n <= 2048u * 1024u) return 1024;
    if (n <= 2048u * 256u) return 256;
    if (n <= 2048u * 512u) return 512;
    if (n <= 2048u * 1024u) return 1024;
    return 0;
}

int mcu_launch_rowst(const mcu_rowst_args *a, void *stream)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    switch (a->n2) {
    case 256: return (int)rowst_launch_256(*a, (cudaStream_t)stream);
    case 512: return (int)rowst_launch_512(*a, (cudaStream_t)stream);
    case 1024: return (int)rowst_launch_1024(*a, (cudaStream_t)stream);
    default: return (int)cudaErrorInvalidValue;
    }
}

int mcu_launch_pass(const mcu_pass_args *a, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    switch (a->radix) {
    case 2: return (int)launch_pass_radix<2>(*a, s);
    case 4: return (int)launch_pass_radix<4>(*a, s);
    case 8: return (int)launch_pass_radix<8>(*a, s);
    default: return (int)cudaErrorInvalidValue;
    }
}

int mcu_launch_resolve(const mcu_resolve_args *a, void *stream)
{
    if (a->n == 0 || a->batch == 0) return 0;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    dim3 grid((a->n + 255) / 256, (unsigned)(a->batch > 65535 ? 65535 : a->batch));
    resolve_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
    return (int)cudaGetLastError();
}

unsigned long long mcu_launch_count(int reset)
{
    return reset ? g_launches.exchange(0) : g_launches.load();
}

} // extern "C"
