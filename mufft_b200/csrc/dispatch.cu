// C-ABI entry points of the CUDA side (see abi.h): runtime helpers, the size dispatch for the
// row / column engine kernels, and the generic kernels used beyond the engine limits:
//   * stockham_pass_kernel   one radix-8/4/2 Stockham pass, any length, 1D or vertical
//                            (the role of ref kernel.c:84-453 and 455-738)
//   * resolve_kernel         stand-alone R2C / C2R resolve and spectral multiply
//                            (ref kernel.c:22-82); also backs mufft_get_convolve_func().
#include "abi.h"
#include "fft_engine.cuh"
#include "rows_kernel.cuh"
#include "generic_kernels.cuh"
#include "engine_tables.h"

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
#define MUFFT_DECLARE_F2D(pw, ph) cudaError_t f2d_launch_##pw##x##ph(const mcu_2d_args &, cudaStream_t); bool f2d_supported_##pw##x##ph(); \
    int f2d_rows_variant_##pw##x##ph(int); unsigned f2d_lag_##pw##x##ph(const mcu_2d_args &);
MUFFT_F2D_SHAPES(MUFFT_DECLARE_F2D)

cudaError_t rowst_launch_256(const mcu_rowst_args &, cudaStream_t);
cudaError_t rowst_launch_512(const mcu_rowst_args &, cudaStream_t);
cudaError_t rowst_launch_1024(const mcu_rowst_args &, cudaStream_t);
// single-launch four-step rows (long_kernel.cuh): n = n1 x n2
#define MUFFT_LONG_SHAPES(X) X(16, 1024) X(32, 1024) X(32, 2048)
#define MUFFT_DECLARE_LONG(n1, n2) cudaError_t long_launch_##n1##x##n2(const mcu_long_args &, cudaStream_t); int long_ctas_##n1##x##n2();
MUFFT_LONG_SHAPES(MUFFT_DECLARE_LONG)

cudaError_t longcols_launch(const mcu_longcols_args &, cudaStream_t, unsigned *lag_only);

static std::atomic<unsigned long long> g_launches{0};

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
int mcu_event_create_timing(void **event)
{
    cudaEvent_t e;
    cudaError_t err = cudaEventCreate(&e);
    *event = (void *)e;
    return (int)err;
}
int mcu_event_elapsed_ms(float *ms, void *start, void *end) { return (int)cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)end); }
int mcu_enable_peer_access(int peer_device)
{
    int dev = 0, can = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return (int)err;
    if (dev == peer_device) return 0;
    if ((err = cudaDeviceCanAccessPeer(&can, dev, peer_device)) != cudaSuccess) return (int)err;
    if (!can) return (int)cudaErrorPeerAccessUnsupported;
    err = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (err == cudaErrorPeerAccessAlreadyEnabled) {
        (void)cudaGetLastError();
        return 0;
    }
    return (int)err;
}
int mcu_capture_begin(void *stream) { return (int)cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal); }
int mcu_capture_end(void *stream, void **exec)
{
    cudaGraph_t graph = nullptr;
    *exec = nullptr;
    cudaError_t err = cudaStreamEndCapture((cudaStream_t)stream, &graph);
    if (err != cudaSuccess || graph == nullptr) {
        (void)cudaGetLastError();
        return err != cudaSuccess ? (int)err : (int)cudaErrorStreamCaptureInvalidated;
    }
    cudaGraphExec_t ge = nullptr;
    err = cudaGraphInstantiate(&ge, graph, 0);
    (void)cudaGraphDestroy(graph);
    if (err != cudaSuccess) {
        (void)cudaGetLastError();
        return (int)err;
    }
    *exec = (void *)ge;
    return 0;
}
int mcu_graph_launch(void *exec, void *stream) { return (int)cudaGraphLaunch((cudaGraphExec_t)exec, (cudaStream_t)stream); }
int mcu_graph_destroy(void *exec) { return exec ? (int)cudaGraphExecDestroy((cudaGraphExec_t)exec) : 0; }
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
int mcu_rows_variant(unsigned n, int in_mode, int out_mode) { return tables_rows_variant(n, in_mode, out_mode); }
size_t mcu_engine_table_size(unsigned n, int variant) { return tables_size(n, variant); }
void mcu_engine_table_fill(float *table, unsigned n, int direction, int variant) { tables_fill(table, n, direction, variant); }
unsigned mcu_split_n2(unsigned n, int columns) { return tables_split_n2(n, columns); }

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

// n -> n1 x n2 of the single-launch four-step kernel.  MUFFT_CUDA_LONG_16K=1 in the environment routes 16384
// points through it as well (16 x 1024) instead of the one-CTA engine kernel (measurement).
static unsigned long_n1_of(unsigned n)
{
    switch (n) {
    case 16384: return 16;
    case 32768: return 32;
    case 65536: return 32;
    default: return 0;
    }
}
unsigned mcu_long_n2(unsigned n)
{
    const unsigned n1 = long_n1_of(n);
    return n1 ? n / n1 : 0;
}
unsigned mcu_long_ctas(unsigned n)
{
    const unsigned n1 = long_n1_of(n);
    if (n1 == 0) return 0;
#define X(a1, a2) if (n1 == a1 && n / n1 == a2) return (unsigned)long_ctas_##a1##x##a2();
    MUFFT_LONG_SHAPES(X)
#undef X
    return 0;
}
int mcu_launch_long(const mcu_long_args *a, void *stream)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
#define X(a1, a2) if (a->n1 == a1 && a->n2 == a2) return (int)long_launch_##a1##x##a2(*a, (cudaStream_t)stream);
    MUFFT_LONG_SHAPES(X)
#undef X
    return (int)cudaErrorInvalidValue;
}

// 16384 and 32768 rows: 1.1x / 1.66x faster than two sweeps of the column kernel; 4096 / 8192 rows (first factor 4 / 8:
// too little work per A tile) measured 1.5-2x SLOWER and keep the two sweeps (profiles/r02h_fastbench_cols.jsonl)
int mcu_longcols_supported(unsigned n) { return n == 16384 || n == 32768; }
size_t mcu_longcols_slot_elems(unsigned n) { return (size_t)n * MUFFT_LONGCOLS_CB; } // one block of columns (longcols_kernel.cuh)
unsigned mcu_longcols_default_lag(unsigned n)
{
    if (!mcu_longcols_supported(n)) return 0;
    mcu_longcols_args a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.n1 = n / 1024; a.direction = -1;
    unsigned lag = 0;
    if (longcols_launch(a, nullptr, &lag) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return lag;
}
int mcu_launch_longcols(const mcu_longcols_args *a, void *stream)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)longcols_launch(*a, (cudaStream_t)stream, nullptr);
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

int mcu_2d_rows_variant(unsigned nx, int kind)
{
#define X(pw, ph) if (nx == pw) return f2d_rows_variant_##pw##x##ph(kind);
    MUFFT_F2D_SHAPES(X)
#undef X
    return 0;
}

unsigned mcu_2d_default_lag(unsigned nx, unsigned ny, int kind)
{
    mcu_2d_args a;
    memset(&a, 0, sizeof(a));
    a.nx = nx; a.ny = ny; a.kind = kind; a.direction = kind == MCU_2D_C2R ? 1 : -1;
#define X(pw, ph) if (nx == pw && ny == ph) return f2d_lag_##pw##x##ph(a);
    MUFFT_F2D_SHAPES(X)
#undef X
    return 0;
}

size_t mcu_2d_mid_elems(unsigned nx, unsigned ny) { return ((size_t)nx + 2) * ny; }

int mcu_launch_2d_fused(const mcu_2d_args *a, void *stream)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
#define X(pw, ph) if (a->nx == pw && a->ny == ph) return (int)f2d_launch_##pw##x##ph(*a, (cudaStream_t)stream);
    MUFFT_F2D_SHAPES(X)
#undef X
    return (int)cudaErrorInvalidValue;
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

void mcu_launch_count_add(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
unsigned long long mcu_launch_count(int reset)
{
    return reset ? g_launches.exchange(0) : g_launches.load();
}

} // extern "C"
