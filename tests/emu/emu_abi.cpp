// TEST INFRASTRUCTURE ONLY: a stand-in for the CUDA side of the library (everything abi.h declares,
// i.e. what dispatch.cu and the *_inst.cu units provide) on top of the CPU emulation of the kernels.
// Linked with the UNCHANGED C host layer (mufft_b200/csrc/mufft_host.c) it gives
// tests/emu/_build/libmufft_hostemu.so: the whole fft.h / mufft_cuda.h surface running on the CPU, so
// that the host logic - plan construction, path selection, table choice, chunking, host staging, the
// 2D and convolution orchestration - is tested against the oracle where no GPU exists
// (tests/test_hostemu.py).  Never shipped, never loaded by the package or the bench.
//
// "Device" memory is ordinary memory; a registry of allocations answers mcu_pointer_kind, so that
// host pointers still take the staged path of mufft_host.c.  Streams and events are dummies: every
// operation completes before its call returns.
#include "abi.h"
#include "engine_tables.h"
#include "generic_kernels.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

using namespace mufft;

// ---- per-size kernel entry points (emu_*_inst.cpp) ---------------------------------------------------
extern "C" {
#define X(n) int emu_launch_rows_##n(const mcu_rows_args *);
MUFFT_FOR_ENGINE_SIZES(X)
#undef X
#define X(n) int emu_launch_conv_##n(const mcu_conv_args *);
MUFFT_FOR_CONV_SIZES(X)
#undef X
#define EMU_FOR_COLS_SIZES(X) X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)
#define X(n) int emu_launch_cols_##n(const mcu_cols_args *);
EMU_FOR_COLS_SIZES(X)
#undef X
int emu_launch_rowst_256(const mcu_rowst_args *);
int emu_launch_rowst_512(const mcu_rowst_args *);
int emu_launch_rowst_1024(const mcu_rowst_args *);
#define EMU_FOR_FOURSTEP_SHAPES(X) X(16, 1024) X(32, 1024) X(32, 2048)
#define X(n1, n2) int emu_launch_long_##n1##x##n2(const mcu_long_args *, unsigned);
EMU_FOR_FOURSTEP_SHAPES(X)
#undef X
int emu_launch_longcols(const mcu_longcols_args *, unsigned);
// fused 2D shapes compiled for the emulation (tests/emu/Makefile: F2D_SHAPES)
#define EMU_FOR_F2D_SHAPES(X) X(256, 256) X(512, 256) X(2048, 256)
#define X(nx, ny) int emu_launch_2d_##nx##x##ny(const mcu_2d_args *, unsigned); int emu_2d_rows_variant_##nx##x##ny(int);
EMU_FOR_F2D_SHAPES(X)
#undef X
}

namespace {

std::atomic<unsigned long long> g_launches{ 0 };
std::mutex g_mem_mutex;
struct block { size_t size; int kind; };
std::map<uintptr_t, block> g_blocks;

void *alloc_kind(size_t bytes, int kind)
{
    bytes = bytes ? bytes : 256;
    void *p = nullptr;
    if (posix_memalign(&p, 256, (bytes + 255) & ~(size_t)255) != 0) return nullptr;
    memset(p, 0xA5, bytes); // "device" memory is not zero-initialised
    std::lock_guard<std::mutex> lock(g_mem_mutex);
    g_blocks[(uintptr_t)p] = block{ bytes, kind };
    return p;
}
int free_kind(void *p)
{
    if (p == nullptr) return 0;
    {
        std::lock_guard<std::mutex> lock(g_mem_mutex);
        g_blocks.erase((uintptr_t)p);
    }
    free(p);
    return 0;
}

} // namespace

extern "C" {

// ---- runtime ---------------------------------------------------------------------------------------
// two "devices" (memory is not tied to either): enough to test the plan / current-device check of the host layer
static thread_local int t_device = 0;
int mcu_device_count(int *count) { *count = 2; return 0; }
int mcu_set_device(int device) { if (device < 0 || device > 1) return 101; t_device = device; return 0; }
int mcu_get_device(int *device) { *device = t_device; return 0; }
int mcu_sm_count(int *count) { *count = 4; return 0; }
const char *mcu_error_string(int err) { return err == 0 ? "no error" : "emulated CUDA side: error"; }

int mcu_ipc_export(void *handle64, void *ptr) { memset(handle64, 0, 64); memcpy(handle64, &ptr, sizeof(ptr)); return 0; }
int mcu_ipc_import(void **ptr, const void *handle64) { memcpy(ptr, handle64, sizeof(*ptr)); return 0; }
int mcu_ipc_close(void *) { return 0; }
int mcu_malloc(void **ptr, size_t bytes) { *ptr = alloc_kind(bytes, MCU_MEM_DEVICE); return *ptr ? 0 : 2; }
int mcu_free(void *ptr) { return free_kind(ptr); }
int mcu_malloc_managed(void **ptr, size_t bytes) { *ptr = alloc_kind(bytes, MCU_MEM_MANAGED); return *ptr ? 0 : 2; }
int mcu_malloc_pinned(void **ptr, size_t bytes) { *ptr = alloc_kind(bytes, MCU_MEM_PINNED); return *ptr ? 0 : 2; }
int mcu_free_pinned(void *ptr) { return free_kind(ptr); }
int mcu_memset(void *ptr, int value, size_t bytes, void *) { memset(ptr, value, bytes); return 0; }

int mcu_pointer_kind(const void *ptr)
{
    std::lock_guard<std::mutex> lock(g_mem_mutex);
    auto it = g_blocks.upper_bound((uintptr_t)ptr);
    if (it == g_blocks.begin()) return MCU_MEM_HOST;
    --it;
    return (uintptr_t)ptr < it->first + it->second.size ? it->second.kind : MCU_MEM_HOST;
}

int mcu_stream_create(void **stream) { *stream = malloc(1); return 0; }
int mcu_stream_destroy(void *stream) { free(stream); return 0; }
int mcu_stream_sync(void *) { return 0; }
int mcu_event_create(void **event) { *event = malloc(1); return 0; }
int mcu_event_destroy(void *event) { free(event); return 0; }
int mcu_event_record(void *, void *) { return 0; }
int mcu_stream_wait_event(void *, void *) { return 0; }
int mcu_event_sync(void *) { return 0; }
int mcu_event_create_timing(void **event) { *event = malloc(1); return 0; }
int mcu_event_elapsed_ms(float *ms, void *, void *) { *ms = 1.0f; return 0; }
int mcu_enable_peer_access(int peer) { return peer < 0 || peer > 1 ? 101 : 0; }
// graphs: the emulated side executes everything at once, so a "capture" has already run its work; report captures as
// unsupported and let the host layer take its ordinary path (counted, so that the fallback is tested)
static std::atomic<unsigned> g_capture_attempts;
int mcu_capture_begin(void *) { g_capture_attempts++; return 801; }
int mcu_capture_end(void *, void **exec) { *exec = nullptr; return 801; }
int mcu_graph_launch(void *, void *) { return 801; }
int mcu_graph_destroy(void *) { return 0; }
unsigned emu_capture_attempts(int reset) { return reset ? g_capture_attempts.exchange(0) : g_capture_attempts.load(); }
int mcu_copy_async(void *dst, const void *src, size_t bytes, void *) { if (bytes) memmove(dst, src, bytes); return 0; }
int mcu_copy2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t rows, void *)
{
    for (size_t r = 0; r < rows; r++) memmove((char *)dst + r * dpitch, (const char *)src + r * spitch, width_bytes);
    return 0;
}
// counted, so that the host layer's managed-memory prefetch policy can be tested: calls to the device / to the host
static std::atomic<unsigned> g_prefetch_calls[2];
int mcu_prefetch(const void *, size_t, int to_device, void *) { g_prefetch_calls[to_device ? 1 : 0]++; return 0; }
unsigned emu_prefetch_calls(int to_device, int reset)
{
    return reset ? g_prefetch_calls[to_device ? 1 : 0].exchange(0) : g_prefetch_calls[to_device ? 1 : 0].load();
}

// ---- tables ------------------------------------------------------------------------------------------
int mcu_rows_variant(unsigned n, int in_mode, int out_mode) { return tables_rows_variant(n, in_mode, out_mode); }
size_t mcu_engine_table_size(unsigned n, int variant) { return tables_size(n, variant); }
void mcu_engine_table_fill(float *table, unsigned n, int direction, int variant) { tables_fill(table, n, direction, variant); }
unsigned mcu_split_n2(unsigned n, int columns) { return tables_split_n2(n, columns); }

// ---- kernels -----------------------------------------------------------------------------------------
int mcu_launch_rows(const mcu_rows_args *a, void *)
{
    g_launches++;
    if (a->batch == 0) return 0;
    switch (a->n) {
#define X(v) case v: return emu_launch_rows_##v(a) == 0 ? 0 : 1;
        MUFFT_FOR_ENGINE_SIZES(X)
#undef X
    default: return 1;
    }
}
// single-launch four-step rows: 3 "resident" CTAs walk the batch
static unsigned long_n1_of(unsigned n) { return n == 16384 ? 16 : (n == 32768 || n == 65536 ? 32 : 0); }
unsigned mcu_long_n2(unsigned n) { return long_n1_of(n) ? n / long_n1_of(n) : 0; }
unsigned mcu_long_ctas(unsigned n) { return long_n1_of(n) ? 3 : 0; }
int mcu_launch_long(const mcu_long_args *a, void *)
{
    g_launches++;
#define X(a1, a2) if (a->n1 == a1 && a->n2 == a2) return emu_launch_long_##a1##x##a2(a, 3);
    EMU_FOR_FOURSTEP_SHAPES(X)
#undef X
    return 1;
}
int mcu_longcols_supported(unsigned n) { return n == 16384 || n == 32768; }
size_t mcu_longcols_slot_elems(unsigned n) { return (size_t)n * 32; }
unsigned mcu_longcols_default_lag(unsigned n) { return mcu_longcols_supported(n) ? 1 : 0; }
int mcu_launch_longcols(const mcu_longcols_args *a, void *)
{
    g_launches++;
    return emu_launch_longcols(a, 3);
}
int mcu_launch_conv(const mcu_conv_args *a, void *)
{
    g_launches++;
    switch (a->n) {
#define X(v) case v: return emu_launch_conv_##v(a);
        MUFFT_FOR_CONV_SIZES(X)
#undef X
    default: return 1;
    }
}
int mcu_launch_cols(const mcu_cols_args *a, void *)
{
    g_launches++;
    switch (a->n) {
#define X(v) case v: return emu_launch_cols_##v(a);
        EMU_FOR_COLS_SIZES(X)
#undef X
    default: return 1;
    }
}
int mcu_launch_rowst(const mcu_rowst_args *a, void *)
{
    g_launches++;
    switch (a->n2) {
    case 256: return emu_launch_rowst_256(a);
    case 512: return emu_launch_rowst_512(a);
    case 1024: return emu_launch_rowst_1024(a);
    default: return 1;
    }
}
int mcu_2d_fused_supported(unsigned nx, unsigned ny)
{
#define X(px, py) if (nx == px && ny == py) return 1;
    EMU_FOR_F2D_SHAPES(X)
#undef X
    return 0;
}
int mcu_2d_rows_variant(unsigned nx, int kind)
{
#define X(px, py) if (nx == px) return emu_2d_rows_variant_##px##x##py(kind);
    EMU_FOR_F2D_SHAPES(X)
#undef X
    return 0;
}
unsigned mcu_2d_default_lag(unsigned, unsigned, int) { return 1; }
size_t mcu_2d_mid_elems(unsigned nx, unsigned ny) { return ((size_t)nx + 2) * ny; }
int mcu_launch_2d_fused(const mcu_2d_args *a, void *)
{
    g_launches++;
    if (a->batch == 0) return 0;
#define X(px, py) if (a->nx == px && a->ny == py) return emu_launch_2d_##px##x##py(a, 3);
    EMU_FOR_F2D_SHAPES(X)
#undef X
    return 1;
}

int mcu_launch_pass(const mcu_pass_args *a, void *)
{
    g_launches++;
    const size_t total = (size_t)(a->n / a->radix) * a->width;
    if (total == 0 || a->batch == 0) return 0;
    const emu::idx3 grid{ (unsigned)((total + 255) / 256), (unsigned)(a->batch > 65535 ? 65535 : a->batch), 1 };
    const bool fwd = a->direction < 0;
    switch (a->radix) {
    case 2: emu::launch(grid, 256, 0, 1, 1, [&] { if (fwd) stockham_pass_kernel<2, -1>(*a); else stockham_pass_kernel<2, 1>(*a); }); return 0;
    case 4: emu::launch(grid, 256, 0, 1, 1, [&] { if (fwd) stockham_pass_kernel<4, -1>(*a); else stockham_pass_kernel<4, 1>(*a); }); return 0;
    case 8: emu::launch(grid, 256, 0, 1, 1, [&] { if (fwd) stockham_pass_kernel<8, -1>(*a); else stockham_pass_kernel<8, 1>(*a); }); return 0;
    default: return 1;
    }
}
int mcu_launch_resolve(const mcu_resolve_args *a, void *)
{
    if (a->n == 0 || a->batch == 0) return 0;
    g_launches++;
    const emu::idx3 grid{ (a->n + 255) / 256, (unsigned)(a->batch > 65535 ? 65535 : a->batch), 1 };
    emu::launch(grid, 256, 0, 1, 1, [&] { resolve_kernel(*a); });
    return 0;
}

void mcu_launch_count_add(unsigned long long n) { g_launches += n; }
unsigned long long mcu_launch_count(int reset) { return reset ? g_launches.exchange(0) : g_launches.load(); }

} // extern "C"
