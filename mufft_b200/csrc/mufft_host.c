/* Host layer of the B200 muFFT implementation: the fft.h entry points (plans, validation,
 * twiddle tables, memory classification, staging) in C, calling the CUDA side only through
 * the C ABI of abi.h.  There is no CPU compute path: every transform runs in a CUDA kernel,
 * and plan creation fails (NULL) when no CUDA device is usable.
 *
 * Mapping to the reference (paths relative to /root/reference):
 *   plan structs            fft.c:54-99      -> struct mufft_plan_1d/_2d/_conv below
 *   build_twiddles          fft.c:102-134    -> axis_init (engine tables: dispatch.cu fill_table)
 *   build_r2c_twiddles      fft.c:441-455    -> make_resolve_twiddles
 *   build_plan_1d/_2d       fft.c:325-435    -> static schedule (engine: fft_engine.cuh; generic: axis_init)
 *   mufft_execute_plan_1d   fft.c:811-850    -> exec_1d -> run_rows (one fused kernel when N <= 8192)
 *   mufft_execute_plan_2d   fft.c:852-989    -> exec_2d -> run_2d_device (row kernel + column kernel)
 *   conv                    fft.c:565-624, 794-809 -> multiply fused into the inverse transform's load
 *   mufft_alloc/free        fft.c:1032-1084  -> CUDA managed memory
 */
#include "fft.h"
#include "mufft_cuda.h"
#include "abi.h"

#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define MUFFT_FLAG_R2C_INTERNAL (1u << 26)
#define MUFFT_FLAG_C2R_INTERNAL (1u << 27)
#define MUFFT_PADDING_COMPLEX_SAMPLES 4 /* ref fft_internal.h:203 */

enum { KIND_C2C = 0, KIND_R2C = 1, KIND_R2C_FULL = 2, KIND_C2R = 3 };

#define CPX 8u                              /* bytes per complex sample */
#define STAGE_SLOTS_MAX 8
/* Staging pipeline of the host-pointer path: slots (one stream each: copy in, kernels, copy out) and
 * bytes per chunk.  Fixed: what a plan does never depends on process-global state.  A build with
 * -DMUFFT_CUDA_DEBUG reads MUFFT_CUDA_STAGE_SLOTS / MUFFT_CUDA_STAGE_CHUNK_MB once, for tuning runs. */
static int g_stage_slots = 4;
static size_t g_stage_chunk_bytes = (size_t)64 << 20;
#ifdef MUFFT_CUDA_DEBUG
__attribute__((constructor)) static void read_env(void)
{
    const char *e = getenv("MUFFT_CUDA_STAGE_SLOTS");
    if (e != NULL && atoi(e) >= 1 && atoi(e) <= STAGE_SLOTS_MAX) g_stage_slots = atoi(e);
    e = getenv("MUFFT_CUDA_STAGE_CHUNK_MB");
    if (e != NULL && atoi(e) >= 1 && atoi(e) <= 1024) g_stage_chunk_bytes = (size_t)atoi(e) << 20;
}
#endif
#define STAGE_SLOTS g_stage_slots
#define STAGE_CHUNK_BYTES g_stage_chunk_bytes
/* Two-kernel transforms go through scratch between their two kernels; the batch is walked in chunks of
 * this many bytes.  L2-sized chunks (8..64 MiB) were measured and LOSE: the many short launches cost more
 * in launch gaps and wave tails than the L2 hits save (profiles/README.md, "large N"). */
#define SPLIT_CHUNK_BYTES ((size_t)1 << 30)
/* 2D transforms on the row kernel + column kernel path: images per launch pair, as bytes of row->column
 * intermediate.  L2-sized groups measured SLOWER than whole-batch launches (profiles/README.md section 3:
 * 2D R2C 1024^2 x 256 1.40 ms with 86 launches against 0.92 ms with 2), so the cap only bounds the scratch. */
#define MID_GROUP_BYTES ((size_t)2048 << 20)
/* Long columns on two sweeps of the column kernel (lengths the one-launch kernel of longcols_kernel.cuh does not take):
 * columns per block, as bytes of the block's intermediate.  L2-sized blocks (16 / 32 / 64 MiB, second sweep discarding the
 * intermediate as it reads it) were measured and LOSE to whole-slab sweeps - 1.47 / 1.23 / 1.09 ms against 1.04 ms at
 * 32768 x 4096 (profiles/r02g_ab_cols_blocks.txt): launch gaps and wave tails again.  The cap only bounds the scratch. */
#ifndef MUFFT_COLS_BLOCK_MB
#define MUFFT_COLS_BLOCK_MB 1024
#endif
#define COLS_BLOCK_BYTES ((size_t)MUFFT_COLS_BLOCK_MB << 20)
#define SCRATCH_CAP_BYTES ((size_t)1 << 30)
#define ZERO_COPY_BYTES ((size_t)64 << 10)   /* pinned buffers up to this size are used in place */
#define MAX_STEPS 32
/* Synchronous executes on CUDA managed memory (what mufft_alloc returns and the reference's callers fill and
 * read on the host between calls): buffers of this many bytes or more are prefetched to the device on the
 * stream in front of the kernel - bulk migration instead of page fault by page fault (measured: 1.00 -> 0.77 ms
 * for one 65536-point transform, 165 -> 145 ms for a 4096 x 4096 image, profiles/r02a_probes.jsonl). */
#define MANAGED_PREFETCH_MIN ((size_t)256 << 10)
/* Synchronous calls that touch at most this many bytes of device-visible memory may be replayed from a CUDA graph */
#define GRAPH_MAX_BYTES ((size_t)1 << 20)
/* Path flags of mufft_cuda.h that a plan keeps */
#define PATH_FLAGS (MUFFT_CUDA_FLAG_SEPARATE_2D | MUFFT_CUDA_FLAG_TWO_KERNEL | MUFFT_CUDA_FLAG_GENERIC_PASSES | MUFFT_CUDA_FLAG_FOUR_STEP_16K)

/* ---- per-thread state --------------------------------------------------------------------- */
static __thread void *tls_stream = NULL;
static __thread char tls_error[256] = "";

/* Error convention of this file: 0 = success; MCU_ERR_REPORTED = failure whose message is already in
 * tls_error (set by the innermost layer that knew what went wrong); any other value = a raw CUDA error
 * code from abi.h that nobody has described yet.  fail() turns the latter into the former exactly once, so
 * that enclosing TRY() layers pass an inner message through instead of overwriting it. */
#define MCU_ERR_REPORTED (-1)
static __thread unsigned tls_warned = 0;
static void warn_stderr(void)
{
    /* the fft.h execute calls return void: stderr is the only signal a caller that does not poll
     * mufft_cuda_last_error() gets.  Rate-limited per thread, not once per process. */
    if (tls_warned < 8) {
        tls_warned++;
        fprintf(stderr, "mufft_b200: %s%s\n", tls_error, tls_warned == 8 ? " (further messages on this thread suppressed)" : "");
    }
}
#if defined(__GNUC__)
__attribute__((format(printf, 1, 2)))
#endif
static int set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tls_error, sizeof(tls_error), fmt, ap);
    va_end(ap);
    warn_stderr();
    return MCU_ERR_REPORTED;
}
static int fail(int err, const char *what)
{
    if (err == 0) return 0;
    if (err != MCU_ERR_REPORTED) return set_error("%s: %s", what, mcu_error_string(err));
    return MCU_ERR_REPORTED;
}
#define TRY(expr) do { int e_ = (expr); if (e_ != 0) return fail(e_, #expr); } while (0)

static int have_device(void)
{
    int count = 0;
    if (mcu_device_count(&count) != 0 || count <= 0) {
        (void)set_error("no usable CUDA device (this library has no CPU path)");
        return 0;
    }
    return 1;
}

int mufft_cuda_device_count(void)
{
    int count = 0;
    if (mcu_device_count(&count) != 0) return 0;
    return count;
}
int mufft_cuda_set_device(int device) { return fail(mcu_set_device(device), "cudaSetDevice"); }
void mufft_cuda_set_stream(void *stream) { tls_stream = stream; }
void *mufft_cuda_get_stream(void) { return tls_stream; }
int mufft_cuda_synchronize(void) { return fail(mcu_stream_sync(tls_stream), "cudaStreamSynchronize"); }
const char *mufft_cuda_last_error(void) { return tls_error; }
unsigned long long mufft_cuda_launch_count(int reset) { return mcu_launch_count(reset); }

/* ---- memory (ref fft.c:1032-1084) ----------------------------------------------------------- */
/* Live allocations made through this library (mufft_alloc / mufft_calloc / mufft_cuda_device_alloc): the kind of
 * memory of a pointer inside one of them is known without asking the driver.  cudaPointerGetAttributes costs
 * about a microsecond per pointer, a fifth of a small synchronous call (tools/latency_probe.cu); the reference's
 * callers hand in mufft_alloc memory (test.c:33-62, bench.c:203-219), so the common case never pays it.
 * Foreign pointers (cudaMalloc, torch, plain host memory) are not in the table and are classified by the driver. */
#define REGISTRY_MAX 512
static struct { uintptr_t base; size_t size; int kind; } g_registry[REGISTRY_MAX];
static int g_registry_count = 0;
static pthread_rwlock_t g_registry_lock = PTHREAD_RWLOCK_INITIALIZER;

static void registry_add(void *ptr, size_t size, int kind)
{
    pthread_rwlock_wrlock(&g_registry_lock);
    if (g_registry_count < REGISTRY_MAX) {
        g_registry[g_registry_count].base = (uintptr_t)ptr;
        g_registry[g_registry_count].size = size ? size : 1;
        g_registry[g_registry_count].kind = kind;
        g_registry_count++;
    }
    pthread_rwlock_unlock(&g_registry_lock);
}

static void registry_remove(void *ptr)
{
    pthread_rwlock_wrlock(&g_registry_lock);
    for (int i = 0; i < g_registry_count; i++)
        if (g_registry[i].base == (uintptr_t)ptr) {
            g_registry[i] = g_registry[--g_registry_count];
            break;
        }
    pthread_rwlock_unlock(&g_registry_lock);
}

static int pointer_kind(const void *ptr)
{
    int kind = -1;
    pthread_rwlock_rdlock(&g_registry_lock);
    for (int i = 0; i < g_registry_count; i++)
        if ((uintptr_t)ptr - g_registry[i].base < g_registry[i].size) {
            kind = g_registry[i].kind;
            break;
        }
    pthread_rwlock_unlock(&g_registry_lock);
    return kind >= 0 ? kind : mcu_pointer_kind(ptr);
}

void *mufft_alloc(size_t size)
{
    void *ptr = NULL;
    if (!have_device()) return NULL;
    if (fail(mcu_malloc_managed(&ptr, size), "cudaMallocManaged") != 0) return NULL;
    registry_add(ptr, size, MCU_MEM_MANAGED);
    return ptr;
}

void *mufft_calloc(size_t size)
{
    void *ptr = mufft_alloc(size);
    if (ptr != NULL) memset(ptr, 0, size);
    return ptr;
}

void mufft_free(void *ptr)
{
    if (ptr == NULL) return;
    registry_remove(ptr);
    (void)mcu_free(ptr);
}

static int is_pow2_ge2(unsigned n) { return n >= 2 && (n & (n - 1)) == 0; }

static int device_visible(int kind, size_t bytes)
{
    if (kind == MCU_MEM_DEVICE || kind == MCU_MEM_MANAGED) return 1;
    return kind == MCU_MEM_PINNED && bytes <= ZERO_COPY_BYTES;
}

/* Upload a host array to a fresh device allocation. */
static void *upload(const void *host, size_t bytes)
{
    void *dev = NULL;
    if (fail(mcu_malloc(&dev, bytes), "cudaMalloc") != 0) return NULL;
    if (fail(mcu_copy_async(dev, host, bytes, NULL), "cudaMemcpyAsync") != 0 ||
        fail(mcu_stream_sync(NULL), "cudaStreamSynchronize") != 0) {
        mcu_free(dev);
        return NULL;
    }
    return dev;
}

/* ---- one transform axis: tables and the generic-pass schedule ---------------------------------- */
typedef struct axis {
    unsigned n;
    int direction;
    void *d_engine_tw; /* engine table: non-NULL = row transforms of this axis run in one kernel */
    unsigned long_n2;  /* != 0: rows run on the single-launch four-step kernel, n = (n / long_n2) x long_n2 (long_kernel.cuh) */
    int generic_only;  /* MUFFT_CUDA_FLAG_GENERIC_PASSES: no two-kernel split either */
    int long_cols;     /* columns run on the single-launch four-step column kernel, n = (n / 1024) x 1024 (longcols_kernel.cuh) */
    void *d_engine_tw_real; /* table of the mirrored-pass shape the R2C (forward) / C2R (inverse) row kernel uses, or NULL */
    void *d_roots;     /* exp(dir*2*pi*i*j/n), j < n: generic multi-pass kernels, two-kernel split */
    unsigned n1, n2;   /* two-kernel split n = n1 * n2 (0 = not used) */
    void *d_tw_n1, *d_tw_n2; /* engine tables of the two factors */
    unsigned nsteps;
    unsigned radix[MAX_STEPS], span[MAX_STEPS];
} axis;

static void axis_free(axis *ax)
{
    mcu_free(ax->d_engine_tw);
    mcu_free(ax->d_engine_tw_real);
    mcu_free(ax->d_roots);
    mcu_free(ax->d_tw_n1);
    mcu_free(ax->d_tw_n2);
    memset(ax, 0, sizeof(*ax));
}

static void *make_engine_table(unsigned n, int direction, int variant)
{
    size_t entries = mcu_engine_table_size(n, variant);
    float *host = calloc(entries ? entries : 1, CPX);
    if (!host) return NULL;
    mcu_engine_table_fill(host, n, direction, variant);
    void *dev = upload(host, (entries ? entries : 1) * CPX);
    free(host);
    return dev;
}

/* Longest row transform one CTA does on chip, given the plan's path flags. */
static unsigned onchip_limit(unsigned flags)
{
    return (flags & (MUFFT_CUDA_FLAG_TWO_KERNEL | MUFFT_CUDA_FLAG_GENERIC_PASSES | MUFFT_CUDA_FLAG_FOUR_STEP_16K)) ? MCU_ENGINE_MAX_N : MCU_ENGINE_BIG_N;
}

/* want_generic: also build what the multi-pass kernels need (AXIS_ROWS / AXIS_COLS: which
 * split the axis takes); 0 = the axis runs on the single-kernel engine only. */
enum { AXIS_ROWS = 1, AXIS_COLS = 2 };
static int axis_init(axis *ax, unsigned n, int direction, int want_generic, unsigned flags)
{
    const int columns = want_generic == AXIS_COLS;
    memset(ax, 0, sizeof(*ax));
    ax->n = n;
    ax->direction = direction;
    ax->generic_only = (flags & MUFFT_CUDA_FLAG_GENERIC_PASSES) != 0;
    if (n <= onchip_limit(flags)) {
        ax->d_engine_tw = make_engine_table(n, direction, 0);
        if (!ax->d_engine_tw) return -1;
        /* the real transforms may run on a different engine shape (abi.h: mcu_rows_variant) */
        const int variant = direction < 0 ? mcu_rows_variant(n, MCU_IN_COMPLEX, MCU_OUT_R2C) : mcu_rows_variant(n, MCU_IN_C2R, MCU_OUT_COMPLEX);
        if (variant != 0) {
            ax->d_engine_tw_real = make_engine_table(n, direction, variant);
            if (!ax->d_engine_tw_real) return -1;
        }
    }
    if (want_generic == AXIS_ROWS && ax->d_engine_tw == NULL && !(flags & (MUFFT_CUDA_FLAG_TWO_KERNEL | MUFFT_CUDA_FLAG_GENERIC_PASSES)) &&
        mcu_long_n2(n) != 0 && mcu_long_ctas(n) != 0) {
        /* one launch, one pass over HBM: n = n1 x n2 inside a persistent CTA, intermediate in L2 */
        ax->long_n2 = mcu_long_n2(n);
        ax->n2 = ax->long_n2;
        ax->n1 = n / ax->n2;
        ax->d_tw_n2 = make_engine_table(ax->n2, direction, 0);
        if (!ax->d_tw_n2) return -1;
    } else if (columns && !(flags & (MUFFT_CUDA_FLAG_TWO_KERNEL | MUFFT_CUDA_FLAG_GENERIC_PASSES)) && mcu_longcols_supported(n)) {
        /* one launch, one pass over HBM: blocks of 32 columns, intermediate in L2 */
        ax->long_cols = 1;
        ax->n2 = 1024;
        ax->n1 = n / 1024;
        ax->d_tw_n2 = make_engine_table(1024, direction, 0);
        if (!ax->d_tw_n2) return -1;
    } else if (want_generic && !ax->generic_only && mcu_split_n2(n, columns) != 0) { /* rows use it above the on-chip limit, columns above MCU_COLS_MAX_N */
        ax->n2 = mcu_split_n2(n, columns);
        ax->n1 = n / ax->n2;
        ax->d_tw_n1 = make_engine_table(ax->n1, direction, 0);
        ax->d_tw_n2 = make_engine_table(ax->n2, direction, 0);
        if (!ax->d_tw_n1 || !ax->d_tw_n2) return -1;
    }
    if (want_generic) {
        /* Static radix-8/4/2 decomposition: widest radix first, the remainder last. */
        unsigned rem = n, p = 1;
        while (rem > 1) {
            unsigned r = rem % 8 == 0 ? 8 : (rem % 4 == 0 ? 4 : 2);
            ax->radix[ax->nsteps] = r;
            ax->span[ax->nsteps] = p;
            ax->nsteps++;
            rem /= r;
            p *= r;
        }
        float *host = malloc((size_t)n * CPX);
        if (!host) return -1;
        for (unsigned j = 0; j < n; j++) {
            /* double precision, rounded once (ref fft.c:104-105) */
            double phase = (M_PI * direction * 2.0 * (double)j) / (double)n;
            host[2 * (size_t)j + 0] = (float)cos(phase);
            host[2 * (size_t)j + 1] = (float)sin(phase);
        }
        ax->d_roots = upload(host, (size_t)n * CPX);
        free(host);
        if (!ax->d_roots) return -1;
    }
    return 0;
}

/* (0 + i*dir) * exp(i*pi*dir*k/n), k < n (ref fft.c:441-455). */
static void *make_resolve_twiddles(int direction, unsigned n)
{
    float *host = malloc((size_t)n * CPX);
    if (!host) return NULL;
    for (unsigned k = 0; k < n; k++) {
        double phase = (M_PI * direction * (double)k) / (double)n;
        float c = (float)cos(phase), s = (float)sin(phase);
        /* (0, dir) * (c, s) = (-dir*s, dir*c) */
        host[2 * (size_t)k + 0] = -(float)direction * s;
        host[2 * (size_t)k + 1] = (float)direction * c;
    }
    void *dev = upload(host, (size_t)n * CPX);
    free(host);
    return dev;
}

/* ---- scratch and staging owned by a plan -------------------------------------------------------- */
typedef struct scratch {
    void *buf[2];
    size_t bytes;
    void *slots;       /* scratch slots of the four-step kernels: one transform per resident CTA / a ring of column blocks; lives in L2 */
    size_t slots_bytes;
    void *counters;    /* ticket and tiles-done counters of the four-step column kernel */
    size_t counters_bytes;
} scratch;

static int scratch_reserve(scratch *s, size_t bytes)
{
    if (bytes <= s->bytes) return 0;
    for (int i = 0; i < 2; i++) {
        mcu_free(s->buf[i]);
        s->buf[i] = NULL;
    }
    s->bytes = 0;
    for (int i = 0; i < 2; i++) TRY(mcu_malloc(&s->buf[i], bytes));
    s->bytes = bytes;
    return 0;
}

static void scratch_free(scratch *s)
{
    mcu_free(s->buf[0]);
    mcu_free(s->buf[1]);
    mcu_free(s->slots);
    mcu_free(s->counters);
    memset(s, 0, sizeof(*s));
}

typedef struct stage_slot {
    void *stream;
    void *d_in, *d_in2, *d_out;
    size_t in_bytes, in2_bytes, out_bytes;
} stage_slot;

typedef struct stage {
    stage_slot slot[STAGE_SLOTS_MAX];
    void *d_shared;
    size_t shared_bytes;
    void *ev_begin;
    void *ev_kernels;   /* recorded after the kernels of each chunk: the next chunk's kernels wait for it */
    int ready;
    int device_plus_1;  /* device the owning plan's tables and scratch live on (0 = not bound to a plan) */
    /* latency path: the last small synchronous call on device-visible memory, as an executable CUDA graph */
    struct {
        void *exec, *capture_stream;
        const void *in, *in2;
        void *out;
        size_t count, in_stride, out_stride, in2_stride;
        unsigned long long launches; /* kernel launches inside the graph (mufft_cuda_launch_count counts them per replay) */
        int key, seen, disabled;
    } graph;
} stage;

/* A plan belongs to the device that was current when it was created: its twiddle tables, scratch and
 * staging buffers live there.  With one host thread per GPU, executing it while another device is
 * current would fault inside the kernel; say so instead. */
static void stage_bind_device(stage *st)
{
    int dev = 0;
    if (mcu_get_device(&dev) == 0) st->device_plus_1 = dev + 1;
}

static int grow(void **ptr, size_t *have, size_t need)
{
    if (need <= *have) return 0;
    mcu_free(*ptr);
    *ptr = NULL;
    *have = 0;
    TRY(mcu_malloc(ptr, need));
    *have = need;
    return 0;
}

static int stage_init(stage *st)
{
    if (st->ready) return 0;
    for (int i = 0; i < STAGE_SLOTS; i++) TRY(mcu_stream_create(&st->slot[i].stream));
    TRY(mcu_event_create(&st->ev_begin));
    TRY(mcu_event_create(&st->ev_kernels));
    st->ready = 1;
    return 0;
}

static void stage_free(stage *st)
{
    for (int i = 0; i < STAGE_SLOTS_MAX; i++) {
        mcu_stream_destroy(st->slot[i].stream);
        mcu_free(st->slot[i].d_in);
        mcu_free(st->slot[i].d_in2);
        mcu_free(st->slot[i].d_out);
    }
    mcu_free(st->d_shared);
    mcu_graph_destroy(st->graph.exec);
    mcu_stream_destroy(st->graph.capture_stream);
    mcu_event_destroy(st->ev_begin);
    mcu_event_destroy(st->ev_kernels);
    memset(st, 0, sizeof(*st));
}

static int stage_check_device(const stage *st)
{
    int dev = -1;
    if (st->device_plus_1 != 0 && mcu_get_device(&dev) == 0 && dev != st->device_plus_1 - 1) {
        return set_error("plan was created on CUDA device %d but device %d is current "
                         "(mufft_cuda_set_device)", st->device_plus_1 - 1, dev);
    }
    return 0;
}

/* Runs `count` units on device-visible pointers; strides in bytes. */
typedef int (*unit_launch)(void *ctx, void *out, size_t out_stride, const void *in, size_t in_stride,
                           const void *in2, size_t in2_stride, size_t count, void *stream);

typedef struct io_desc {
    void *out;
    const void *in, *in2;
    size_t out_stride, in_stride, in2_stride; /* user strides in bytes (in2_stride 0 = shared operand) */
    size_t out_unit, in_unit, in2_unit;       /* bytes actually touched per unit (tight device pitch) */
    size_t count;
} io_desc;

#define ALIGNED(ptr) (((uintptr_t)(ptr) % CPX) == 0)

static int copy_units(void *dst, size_t dpitch, const void *src, size_t spitch, size_t unit, size_t count, void *stream)
{
    if (dpitch == unit && spitch == unit) return mcu_copy_async(dst, src, unit * count, stream);
    return mcu_copy2d_async(dst, dpitch, src, spitch, unit, count, stream);
}

/* Chunked, pipelined host<->device staging: slot s handles chunks s, s+SLOTS, ... on its own
 * stream (copy in, kernels, copy out), so the copies of one chunk overlap the kernels and the
 * opposite-direction copies of its neighbours. */
static int run_any(stage *st, unit_launch fn, void *ctx, const io_desc *io, int async, int graph_key)
{
    if (io->count == 0) return 0;
    if (stage_check_device(st) != 0) return -1;
    const size_t span_in = io->in_stride * (io->count - 1) + io->in_unit;
    const size_t span_out = io->out_stride * (io->count - 1) + io->out_unit;
    /* The kernels move 8-byte complex values: a buffer that is not 8-byte aligned (legal for the real
     * side of R2C / C2R, e.g. a float array on the caller's stack; the reference's scalar path accepts
     * it) cannot be used in place, whatever memory it is.  It takes the staged path below, whose
     * copies have no alignment rule. */
    const int kind_in = pointer_kind(io->in), kind_out = pointer_kind(io->out);
    const int vis_in = ALIGNED(io->in) && device_visible(kind_in, span_in);
    const int vis_out = ALIGNED(io->out) && device_visible(kind_out, span_out);
    int vis_in2 = 1;
    if (io->in2) {
        size_t span2 = io->in2_stride ? io->in2_stride * (io->count - 1) + io->in2_unit : io->in2_unit;
        vis_in2 = ALIGNED(io->in2) && device_visible(pointer_kind(io->in2), span2);
    }

    if (vis_in && vis_out && vis_in2) {
        const int pf_in = !async && kind_in == MCU_MEM_MANAGED && span_in >= MANAGED_PREFETCH_MIN;
        const int pf_out = !async && kind_out == MCU_MEM_MANAGED && span_out >= MANAGED_PREFETCH_MIN;
        if (pf_in) (void)mcu_prefetch(io->in, span_in, 1, tls_stream);   /* advisory: errors are ignored */
        if (pf_out) (void)mcu_prefetch(io->out, span_out, 1, tls_stream);
        /* Latency path (BASELINE configs[0]: one small transform per fft.h call).  Such a call is
         * launch-bound: on the measured box a bare kernel launch + cudaStreamSynchronize is 8.7 us, the same launch
         * replayed from a CUDA graph 6.8 us (tools/latency_probe.cu, profiles/r02d_latency.jsonl).  The second time a
         * small synchronous call comes with the same pointers it is captured into a graph, from then on replayed. */
        if (!async && graph_key != 0 && !st->graph.disabled && span_in + span_out <= GRAPH_MAX_BYTES) {
            const int same = st->graph.key == graph_key && st->graph.in == io->in && st->graph.in2 == io->in2 &&
                             st->graph.out == io->out && st->graph.count == io->count && st->graph.in_stride == io->in_stride &&
                             st->graph.out_stride == io->out_stride && st->graph.in2_stride == io->in2_stride;
            if (same && st->graph.exec != NULL) {
                mcu_launch_count_add(st->graph.launches);
                TRY(mcu_graph_launch(st->graph.exec, tls_stream));
                TRY(mcu_stream_sync(tls_stream));
                return 0;
            }
            if (same && st->graph.seen >= 1) {
                /* second identical call (the first ran uncaptured and did every allocation / attribute call): capture */
                int ok = st->graph.capture_stream != NULL || mcu_stream_create(&st->graph.capture_stream) == 0;
                if (ok && mcu_capture_begin(st->graph.capture_stream) == 0) {
                    const unsigned long long before = mcu_launch_count(0);
                    const int err = fn(ctx, io->out, io->out_stride, io->in, io->in_stride, io->in2, io->in2_stride, io->count, st->graph.capture_stream);
                    void *exec = NULL;
                    ok = mcu_capture_end(st->graph.capture_stream, &exec) == 0 && err == 0 && exec != NULL;
                    if (ok) {
                        mcu_graph_destroy(st->graph.exec);
                        st->graph.exec = exec;
                        st->graph.launches = mcu_launch_count(0) - before;
                        TRY(mcu_graph_launch(st->graph.exec, tls_stream));
                        TRY(mcu_stream_sync(tls_stream));
                        return 0;
                    }
                    mcu_graph_destroy(exec);
                } else {
                    ok = 0;
                }
                if (!ok) {
                    st->graph.disabled = 1; /* this path cannot be captured: never try again, run it the ordinary way */
                    tls_error[0] = 0;
                }
            } else if (!same) {
                mcu_graph_destroy(st->graph.exec);
                st->graph.exec = NULL;
                st->graph.key = graph_key; st->graph.in = io->in; st->graph.in2 = io->in2; st->graph.out = io->out; st->graph.count = io->count;
                st->graph.in_stride = io->in_stride; st->graph.out_stride = io->out_stride; st->graph.in2_stride = io->in2_stride;
                st->graph.seen = 0;
            }
            st->graph.seen++;
        }
        TRY(fn(ctx, io->out, io->out_stride, io->in, io->in_stride, io->in2, io->in2_stride, io->count, tls_stream));
        if (!async) TRY(mcu_stream_sync(tls_stream));
        return 0;
    }
    if (async) {
        return set_error("asynchronous execute needs device-accessible (device or managed), 8-byte aligned pointers");
    }

    TRY(stage_init(st));
    /* Work already queued on the caller's stream may produce our device-side operands. */
    TRY(mcu_event_record(st->ev_begin, tls_stream));
    for (int s = 0; s < STAGE_SLOTS; s++) TRY(mcu_stream_wait_event(st->slot[s].stream, st->ev_begin));

    const void *in2_dev = io->in2;
    if (io->in2 && !vis_in2 && io->in2_stride == 0) {
        TRY(grow(&st->d_shared, &st->shared_bytes, io->in2_unit));
        TRY(mcu_copy_async(st->d_shared, io->in2, io->in2_unit, st->slot[0].stream));
        TRY(mcu_stream_sync(st->slot[0].stream));
        in2_dev = st->d_shared;
    }

    size_t biggest = io->in_unit > io->out_unit ? io->in_unit : io->out_unit;
    size_t chunk = STAGE_CHUNK_BYTES / (biggest ? biggest : 1);
    if (chunk == 0) chunk = 1;
    /* keep every slot busy even for mid-sized jobs */
    if (chunk * STAGE_SLOTS > io->count) chunk = (io->count + STAGE_SLOTS - 1) / STAGE_SLOTS;

    size_t done = 0;
    for (unsigned c = 0; done < io->count; c++) {
        stage_slot *sl = &st->slot[c % STAGE_SLOTS];
        const size_t cnt = io->count - done < chunk ? io->count - done : chunk;
        const char *src = (const char *)io->in + done * io->in_stride;
        char *dst = (char *)io->out + done * io->out_stride;
        const void *d_in = src, *d_in2 = in2_dev;
        void *d_out = dst;
        size_t s_in = io->in_stride, s_out = io->out_stride, s_in2 = io->in2_stride;

        if (!vis_in) {
            TRY(grow(&sl->d_in, &sl->in_bytes, cnt * io->in_unit));
            TRY(copy_units(sl->d_in, io->in_unit, src, io->in_stride, io->in_unit, cnt, sl->stream));
            d_in = sl->d_in;
            s_in = io->in_unit;
        }
        if (io->in2 && io->in2_stride != 0) {
            const char *src2 = (const char *)io->in2 + done * io->in2_stride;
            d_in2 = src2;
            if (!vis_in2) {
                TRY(grow(&sl->d_in2, &sl->in2_bytes, cnt * io->in2_unit));
                TRY(copy_units(sl->d_in2, io->in2_unit, src2, io->in2_stride, io->in2_unit, cnt, sl->stream));
                d_in2 = sl->d_in2;
                s_in2 = io->in2_unit;
            }
        }
        if (!vis_out) {
            TRY(grow(&sl->d_out, &sl->out_bytes, cnt * io->out_unit));
            d_out = sl->d_out;
            s_out = io->out_unit;
        }
        /* The kernels of consecutive chunks run on different streams but may share plan-owned
         * scratch (2D intermediates and ticket counters, two-kernel transforms, convolution
         * spectra): chain them.  The copies on either side still overlap the neighbours' kernels. */
        if (c > 0) TRY(mcu_stream_wait_event(sl->stream, st->ev_kernels));
        TRY(fn(ctx, d_out, s_out, d_in, s_in, d_in2, s_in2, cnt, sl->stream));
        TRY(mcu_event_record(st->ev_kernels, sl->stream));
        if (!vis_out) TRY(copy_units(dst, io->out_stride, sl->d_out, io->out_unit, io->out_unit, cnt, sl->stream));
        done += cnt;
    }
    for (int s = 0; s < STAGE_SLOTS; s++) TRY(mcu_stream_sync(st->slot[s].stream));
    return 0;
}

/* ---- 1D plans (ref fft.c:54-66, 476-547, 626-662) ------------------------------------------------ */
struct mufft_plan_1d {
    unsigned n; /* complex transform length (N/2 for R2C / C2R) */
    int kind;
    int direction;
    int zero_pad;
    int use_engine;    /* one launch per batch: on-chip engine or the L2-resident four-step kernel */
    unsigned flags;    /* path flags of mufft_cuda.h (PATH_FLAGS) */
    axis ax;
    axis ax_cols;      /* tables for mufft_execute_plan_1d_columns, built on first use */
    void *d_rtw;
    scratch scr;
    scratch scr_cols;
    stage st;
};

void mufft_free_plan_1d(mufft_plan_1d *plan)
{
    if (plan == NULL) return;
    axis_free(&plan->ax);
    axis_free(&plan->ax_cols);
    scratch_free(&plan->scr_cols);
    mcu_free(plan->d_rtw);
    scratch_free(&plan->scr);
    stage_free(&plan->st);
    free(plan);
}

mufft_plan_1d *mufft_create_plan_1d_c2c(unsigned N, int direction, unsigned flags)
{
    /* ref fft.c:628-631 rejects non powers of two and N == 1; N == 0 is rejected here too. */
    if (!is_pow2_ge2(N)) return NULL;
    if (direction != MUFFT_FORWARD && direction != MUFFT_INVERSE) return NULL;
    if (!have_device()) return NULL;

    mufft_plan_1d *plan = calloc(1, sizeof(*plan));
    if (plan == NULL) return NULL;
    plan->n = N;
    plan->kind = KIND_C2C;
    plan->direction = direction;
    plan->zero_pad = (flags & MUFFT_FLAG_ZERO_PAD_UPPER_HALF) != 0;
    plan->flags = flags & PATH_FLAGS;
    stage_bind_device(&plan->st);
    if (axis_init(&plan->ax, N, direction, N <= onchip_limit(flags) ? 0 : AXIS_ROWS, flags) != 0) {
        mufft_free_plan_1d(plan);
        return NULL;
    }
    plan->use_engine = plan->ax.d_engine_tw != NULL || plan->ax.long_n2 != 0;
    return plan;
}

mufft_plan_1d *mufft_create_plan_1d_r2c(unsigned N, unsigned flags)
{
    if (!is_pow2_ge2(N)) return NULL;
    mufft_plan_1d *plan = mufft_create_plan_1d_c2c(N / 2, MUFFT_FORWARD, flags); /* NULL for N == 2 */
    if (plan == NULL) return NULL;
    plan->kind = (flags & MUFFT_FLAG_FULL_R2C) ? KIND_R2C_FULL : KIND_R2C;
    plan->d_rtw = make_resolve_twiddles(MUFFT_FORWARD, N / 2);
    if (plan->d_rtw == NULL) {
        mufft_free_plan_1d(plan);
        return NULL;
    }
    return plan;
}

mufft_plan_1d *mufft_create_plan_1d_c2r(unsigned N, unsigned flags)
{
    if (!is_pow2_ge2(N)) return NULL;
    mufft_plan_1d *plan = mufft_create_plan_1d_c2c(N / 2, MUFFT_INVERSE, flags);
    if (plan == NULL) return NULL;
    plan->kind = KIND_C2R;
    plan->d_rtw = make_resolve_twiddles(MUFFT_INVERSE, N / 2);
    if (plan->d_rtw == NULL) {
        mufft_free_plan_1d(plan);
        return NULL;
    }
    return plan;
}

int mufft_cuda_plan_1d_uses_engine(const mufft_plan_1d *plan) { return plan ? plan->use_engine : 0; }

/* Destination of MCU_OUT_PEERS row transforms (set by mufft_execute_plan_1d_batch_scatter_async). */
static __thread const mcu_peer_dest *tls_peers = NULL;

/* Row transforms on device-visible memory; strides in complex elements.  One fused kernel on
 * the engine, or pre-resolve / passes / post-resolve through scratch beyond its limit. */
typedef struct part_desc {
    unsigned parts;
    long long in_step, in2_step; /* complex elements between consecutive (spectrum, filter partition) pairs */
} part_desc;

/* The n = n1 x n2 core of a long row transform on device memory, `cnt` transforms: the single-launch
 * four-step kernel, the two-kernel split through scr->buf[0], or one generic pass per radix ping-ponging
 * through scr->buf[0..1].  to_peers: the output goes to the slabs of tls_peers (rows offset by row_offset).
 * dst == NULL: the result stays in scratch (tight, n elements apart) and *where says in which buffer. */
static int run_long_core(const axis *ax, scratch *scr, void *dst, size_t dst_stride, const void *cur, size_t cur_stride,
                         size_t cnt, int half, int to_peers, size_t row_offset, void *stream, const void **where);

static int run_rows(const axis *ax, scratch *scr, const void *d_rtw, int in_mode, int out_mode, float norm,
                    void *out, size_t out_stride, const void *in, size_t in_stride,
                    const void *in2, size_t in2_stride, size_t batch, void *stream, const part_desc *pd)
{
    const unsigned n = ax->n;
    if ((in_mode == MCU_IN_MULACC || in_mode == MCU_IN_MULACC_C2R) && (pd == NULL || ax->d_engine_tw == NULL)) {
        return set_error("partitioned convolution needs a transform that fits the on-chip engine "
                         "(stereo N <= %u, mono N <= %u)", MCU_ENGINE_BIG_N, 2 * MCU_ENGINE_BIG_N);
    }
    if (ax->d_engine_tw != NULL) {
        mcu_rows_args a;
        memset(&a, 0, sizeof(a));
        a.out = out; a.in = in; a.in2 = in2; a.rtw = d_rtw;
        a.tw = mcu_rows_variant(n, in_mode, out_mode) != 0 ? ax->d_engine_tw_real : ax->d_engine_tw;
        a.out_stride = out_stride; a.in_stride = in_stride; a.in2_stride = in2_stride;
        a.batch = batch; a.n = n; a.direction = ax->direction;
        a.in_mode = in_mode; a.out_mode = out_mode; a.norm = norm;
        if (out_mode == MCU_OUT_PEERS || out_mode == MCU_OUT_R2C_PEERS) a.peers = *tls_peers;
        if (pd != NULL) { a.parts = pd->parts; a.in_part_stride = pd->in_step; a.in2_part_stride = pd->in2_step; }
        return mcu_launch_rows(&a, stream);
    }
    if (out_mode == MCU_OUT_PEERS && ax->n2 == 0) return set_error("scatter: this length has neither a four-step nor a two-kernel path");
    if (out_mode == MCU_OUT_R2C_PEERS) return set_error("scatter: R2C rows must fit the single-kernel engine");

    const int pre = in_mode == MCU_IN_C2R || in_mode == MCU_IN_MUL || in_mode == MCU_IN_MUL_C2R;
    const int post = out_mode != MCU_OUT_COMPLEX && out_mode != MCU_OUT_PEERS;
    if (ax->long_n2 != 0 && !pre && !post) /* plain or zero-padded C2C: one launch for the whole batch, no chunking */
        return run_long_core(ax, scr, out, out_stride, in, in_stride, batch, in_mode == MCU_IN_HALF, out_mode == MCU_OUT_PEERS, 0, stream, NULL);

    /* With a resolve / multiply sweep in front or behind, the transform works through bounded scratch:
     * walk the batch in chunks. */
    size_t chunk = (ax->n2 != 0 ? SPLIT_CHUNK_BYTES : SCRATCH_CAP_BYTES) / ((size_t)n * CPX);
    if (chunk == 0) chunk = 1;
    if (chunk > batch) chunk = batch;
    TRY(scratch_reserve(scr, chunk * n * CPX));

    for (size_t done = 0; done < batch; done += chunk) {
        const size_t cnt = batch - done < chunk ? batch - done : chunk;
        const char *src = (const char *)in + done * in_stride * CPX;
        const char *src2 = in2 ? (const char *)in2 + done * in2_stride * CPX : NULL;
        char *dst = (char *)out + done * out_stride * CPX;
        const void *cur = src;
        size_t cur_stride = in_stride;

        if (pre) {
            mcu_resolve_args r;
            memset(&r, 0, sizeof(r));
            r.out = scr->buf[1]; r.in = src; r.in2 = src2; r.rtw = d_rtw;
            r.out_stride = n; r.in_stride = in_stride; r.in2_stride = in2_stride; r.batch = cnt; r.n = n;
            r.mode = in_mode == MCU_IN_C2R ? MCU_RES_C2R : (in_mode == MCU_IN_MUL ? MCU_RES_MUL : MCU_RES_MUL_C2R);
            r.norm = norm;
            TRY(mcu_launch_resolve(&r, stream));
            cur = scr->buf[1];
            cur_stride = n;
        }
        if (post) {
            /* transform into scratch (dst NULL: run_long_core says where), then resolve into the output.  An input
             * and an output sweep never occur together: R2C has no input sweep, C2R / multiply no output sweep. */
            const void *mid = NULL;
            TRY(run_long_core(ax, scr, NULL, n, cur, cur_stride, cnt, in_mode == MCU_IN_HALF, 0, 0, stream, &mid));
            mcu_resolve_args r;
            memset(&r, 0, sizeof(r));
            r.out = dst; r.in = mid; r.rtw = d_rtw;
            r.out_stride = out_stride; r.in_stride = n; r.batch = cnt; r.n = n;
            r.mode = out_mode == MCU_OUT_R2C_FULL ? MCU_RES_R2C_FULL : MCU_RES_R2C;
            TRY(mcu_launch_resolve(&r, stream));
        } else {
            TRY(run_long_core(ax, scr, dst, out_stride, cur, cur_stride, cnt, in_mode == MCU_IN_HALF, out_mode == MCU_OUT_PEERS, done, stream, NULL));
        }
    }
    return 0;
}

static int run_long_core(const axis *ax, scratch *scr, void *dst, size_t dst_stride, const void *cur, size_t cur_stride,
                         size_t cnt, int half, int to_peers, size_t row_offset, void *stream, const void **where)
{
    const unsigned n = ax->n;
    const int in_scratch = dst == NULL && !to_peers;
    if (in_scratch && (ax->long_n2 != 0 || ax->n2 != 0)) { /* neither path uses buf[1] itself */
        dst = scr->buf[1];
        dst_stride = n;
        *where = dst;
    }
    if (ax->long_n2 != 0) {
        size_t slots = mcu_long_ctas(n); /* one per resident CTA */
        if (slots > cnt) slots = cnt;
        if (slots == 0) return set_error("four-step kernel: no resident CTA on this device");
        TRY(grow(&scr->slots, &scr->slots_bytes, slots * n * CPX));
        mcu_long_args la;
        memset(&la, 0, sizeof(la));
        la.out = dst; la.in = cur; la.tw = ax->d_tw_n2; la.roots = ax->d_roots;
        la.scratch = scr->slots; la.scratch_slots = scr->slots_bytes / ((size_t)n * CPX);
        la.out_batch = dst_stride; la.in_batch = cur_stride; la.batch = cnt;
        la.n = n; la.n1 = ax->n1; la.n2 = ax->n2; la.direction = ax->direction;
        la.in_mode = half ? MCU_IN_HALF : MCU_IN_COMPLEX;
        if (to_peers) { la.to_peers = 1; la.peers = *tls_peers; la.peers.row0 += row_offset; }
        return mcu_launch_long(&la, stream);
    }
    if (ax->n2 != 0) {
        /* two kernels: N2 column transforms of length N1 on the [N1][N2] view into scratch,
         * then twiddle + N1 row transforms of length N2 with the transposed store */
        mcu_cols_args ca;
        memset(&ca, 0, sizeof(ca));
        ca.out = scr->buf[0]; ca.in = cur; ca.tw = ax->d_tw_n1;
        ca.out_pitch = ax->n2; ca.in_pitch = ax->n2; ca.out_image = n; ca.in_image = cur_stride;
        ca.batch = cnt; ca.n = ax->n1; ca.width = ax->n2; ca.direction = ax->direction;
        ca.rows_valid = half ? ax->n1 / 2 : 0;
        TRY(mcu_launch_cols(&ca, stream));
        mcu_rowst_args ra;
        memset(&ra, 0, sizeof(ra));
        ra.in = scr->buf[0]; ra.in_batch = n; ra.tw = ax->d_tw_n2; ra.roots = ax->d_roots;
        if (to_peers) { ra.to_peers = 1; ra.peers = *tls_peers; ra.peers.row0 += row_offset; }
        else { ra.out = dst; ra.out_batch = dst_stride; }
        ra.batch = cnt; ra.n = n; ra.n1 = ax->n1; ra.n2 = ax->n2; ra.direction = ax->direction;
        return mcu_launch_rowst(&ra, stream);
    }
    if (to_peers) return set_error("scatter: this length has neither a four-step nor a two-kernel path");
    for (unsigned s = 0; s < ax->nsteps; s++) {
        const int last = s + 1 == ax->nsteps;
        mcu_pass_args p;
        memset(&p, 0, sizeof(p));
        p.in = cur; p.in_batch = cur_stride; p.in_stride = 1;
        if (last && !in_scratch) { p.out = dst; p.out_batch = dst_stride; }
        else { p.out = scr->buf[s & 1]; p.out_batch = n; }
        if (last && in_scratch) *where = p.out;
        p.out_stride = 1;
        p.roots = ax->d_roots; p.batch = cnt; p.n = n; p.width = 1;
        p.radix = ax->radix[s]; p.p = ax->span[s]; p.direction = ax->direction;
        p.half = half && s == 0;
        TRY(mcu_launch_pass(&p, stream));
        cur = p.out;
        cur_stride = p.out_batch;
    }
    return 0;
}

typedef struct rows_ctx {
    mufft_plan_1d *plan;
    int in_mode, out_mode;
    float norm;
} rows_ctx;

static int rows_unit_launch(void *ctx_, void *out, size_t out_stride, const void *in, size_t in_stride,
                            const void *in2, size_t in2_stride, size_t count, void *stream)
{
    rows_ctx *ctx = ctx_;
    mufft_plan_1d *plan = ctx->plan;
    return run_rows(&plan->ax, &plan->scr, plan->d_rtw, ctx->in_mode, ctx->out_mode, ctx->norm,
                    out, out_stride / CPX, in, in_stride / CPX, in2, in2_stride / CPX, count, stream, NULL);
}

static size_t in_unit_bytes(unsigned n, int in_mode)
{
    switch (in_mode) {
    case MCU_IN_HALF: return (size_t)(n / 2) * CPX;
    case MCU_IN_C2R:
    case MCU_IN_MUL_C2R: return ((size_t)n + 1) * CPX;
    default: return (size_t)n * CPX;
    }
}

static size_t out_unit_bytes(unsigned n, int out_mode)
{
    switch (out_mode) {
    case MCU_OUT_R2C: return ((size_t)n + 1) * CPX;
    case MCU_OUT_R2C_FULL: return (size_t)n * 2 * CPX;
    default: return (size_t)n * CPX;
    }
}

static void plan_modes(const mufft_plan_1d *plan, int *in_mode, int *out_mode)
{
    *in_mode = plan->kind == KIND_C2R ? MCU_IN_C2R : (plan->zero_pad ? MCU_IN_HALF : MCU_IN_COMPLEX);
    *out_mode = plan->kind == KIND_R2C ? MCU_OUT_R2C : (plan->kind == KIND_R2C_FULL ? MCU_OUT_R2C_FULL : MCU_OUT_COMPLEX);
}

static int exec_1d(mufft_plan_1d *plan, int in_mode, int out_mode, float norm,
                   void *output, size_t out_stride, const void *input, size_t in_stride,
                   const void *input2, size_t in2_stride, size_t in2_unit, size_t batch, int async)
{
    rows_ctx ctx = { plan, in_mode, out_mode, norm };
    io_desc io;
    memset(&io, 0, sizeof(io));
    io.out = output; io.in = input; io.in2 = input2;
    io.in_unit = in_unit_bytes(plan->n, in_mode);
    io.out_unit = out_unit_bytes(plan->n, out_mode);
    io.in2_unit = in2_unit;
    io.in_stride = in_stride ? in_stride : io.in_unit;
    io.out_stride = out_stride ? out_stride : io.out_unit;
    io.in2_stride = in2_stride;
    io.count = batch;
    if ((io.in_stride | io.out_stride | io.in2_stride) % CPX != 0) {
        return set_error("strides must be multiples of 8 bytes");
    }
    return run_any(&plan->st, rows_unit_launch, &ctx, &io, async, 1 + in_mode * 16 + out_mode);
}

void mufft_execute_plan_1d(mufft_plan_1d *plan, void *MUFFT_RESTRICT output, const void *MUFFT_RESTRICT input)
{
    int in_mode, out_mode;
    plan_modes(plan, &in_mode, &out_mode);
    (void)exec_1d(plan, in_mode, out_mode, 1.0f, output, 0, input, 0, NULL, 0, 0, 1, 0);
}

int mufft_execute_plan_1d_batch(mufft_plan_1d *plan, void *output, const void *input, size_t batch,
                                size_t output_stride_bytes, size_t input_stride_bytes)
{
    int in_mode, out_mode;
    plan_modes(plan, &in_mode, &out_mode);
    return exec_1d(plan, in_mode, out_mode, 1.0f, output, output_stride_bytes, input, input_stride_bytes, NULL, 0, 0, batch, 0);
}

int mufft_execute_plan_1d_batch_async(mufft_plan_1d *plan, void *output, const void *input, size_t batch,
                                      size_t output_stride_bytes, size_t input_stride_bytes)
{
    int in_mode, out_mode;
    plan_modes(plan, &in_mode, &out_mode);
    return exec_1d(plan, in_mode, out_mode, 1.0f, output, output_stride_bytes, input, input_stride_bytes, NULL, 0, 0, batch, 1);
}

/* Exchange step of a slab-decomposed 2D transform fused into the row pass (see mufft_cuda.h). */
int mufft_execute_plan_1d_batch_scatter_async(mufft_plan_1d *plan, void *const *peer_outputs, unsigned peers,
                                              size_t dest_row0, const void *input, size_t batch, size_t input_stride_bytes)
{
    const int real = plan != NULL && plan->kind == KIND_R2C;
    if (plan == NULL || (plan->kind != KIND_C2C && !real) || plan->zero_pad || peers == 0 || peers > MCU_MAX_PEERS ||
        (peers & (peers - 1)) != 0 || plan->n % peers != 0 || input_stride_bytes % CPX != 0) {
        return set_error("scatter: needs a plain C2C or R2C plan and a power-of-two number of devices dividing its complex length");
    }
    if (stage_check_device(&plan->st) != 0) return -1;
    if (real && plan->ax.d_engine_tw == NULL) {
        return set_error("scatter: R2C rows must fit the single-kernel engine (N <= %u reals)", 2 * MCU_ENGINE_BIG_N);
    }
    /* plan->n complex points per row either way (an R2C row of N reals is N/2 complex) */
    const size_t in_stride = input_stride_bytes ? input_stride_bytes : (size_t)plan->n * CPX;
    if (batch == 0) return 0;
    if (!device_visible(pointer_kind(input), in_stride * (batch - 1) + (size_t)plan->n * CPX)) {
        return set_error("scatter: needs device-accessible (device or managed) pointers");
    }
    mcu_peer_dest d;
    memset(&d, 0, sizeof(d));
    const unsigned cols = plan->n / peers;
    for (unsigned p = 0; p < peers; p++) d.base[p] = peer_outputs[p];
    d.row0 = dest_row0;
    d.pitch = cols;
    d.pitch_last = real ? cols + 1 : cols; /* the Nyquist bin is one more column of the last slab */
    d.last = peers - 1;
    while ((1u << d.shift) < cols) d.shift++;
    tls_peers = &d;
    int err = run_rows(&plan->ax, &plan->scr, plan->d_rtw, MCU_IN_COMPLEX, real ? MCU_OUT_R2C_PEERS : MCU_OUT_PEERS, 1.0f, NULL, 0,
                       input, in_stride / CPX, NULL, 0, batch, tls_stream, NULL);
    tls_peers = NULL;
    return err;
}

void *mufft_cuda_device_alloc(size_t size)
{
    void *ptr = NULL;
    if (!have_device()) return NULL;
    if (fail(mcu_malloc(&ptr, size), "cudaMalloc") != 0) return NULL;
    registry_add(ptr, size, MCU_MEM_DEVICE);
    return ptr;
}
void mufft_cuda_device_free(void *ptr)
{
    if (ptr == NULL) return;
    registry_remove(ptr);
    (void)mcu_free(ptr);
}
void *mufft_cuda_host_alloc(size_t size)
{
    void *ptr = NULL;
    if (!have_device()) return NULL;
    if (fail(mcu_malloc_pinned(&ptr, size), "cudaHostAlloc") != 0) return NULL;
    registry_add(ptr, size, MCU_MEM_PINNED);
    return ptr;
}
void mufft_cuda_host_free(void *ptr)
{
    if (ptr == NULL) return;
    registry_remove(ptr);
    (void)mcu_free_pinned(ptr);
}
int mufft_cuda_copy(void *dst, const void *src, size_t bytes)
{
    TRY(mcu_copy_async(dst, src, bytes, tls_stream));
    TRY(mcu_stream_sync(tls_stream));
    return 0;
}
int mufft_cuda_memset(void *ptr, int value, size_t bytes)
{
    TRY(mcu_memset(ptr, value, bytes, tls_stream));
    TRY(mcu_stream_sync(tls_stream));
    return 0;
}
int mufft_cuda_ipc_export(void *handle64, void *device_ptr) { return fail(mcu_ipc_export(handle64, device_ptr), "cudaIpcGetMemHandle"); }
void *mufft_cuda_ipc_import(const void *handle64)
{
    void *ptr = NULL;
    if (fail(mcu_ipc_import(&ptr, handle64), "cudaIpcOpenMemHandle") != 0) return NULL;
    return ptr;
}
int mufft_cuda_ipc_close(void *ptr) { return fail(mcu_ipc_close(ptr), "cudaIpcCloseMemHandle"); }

static int run_cols_axis(const axis *ax, scratch *scr_cols, void *out, size_t out_pitch, size_t out_image,
                         const void *in, size_t in_pitch, size_t in_image, unsigned width, size_t images, void *stream);

/* The plan's C2C transform applied down the columns of a row-major [N][width] matrix on device
 * memory (what the column pass of a slab-decomposed 2D transform needs: mufft_b200/slab.py).
 * Enqueued on the current stream. */
int mufft_execute_plan_1d_columns(mufft_plan_1d *plan, void *output, const void *input, size_t width,
                                  size_t output_pitch_bytes, size_t input_pitch_bytes)
{
    if (plan == NULL || plan->kind != KIND_C2C || plan->zero_pad || width == 0 || width > 0xffffffffu ||
        (output_pitch_bytes | input_pitch_bytes) % CPX != 0) {
        return set_error("columns: needs a plain C2C plan and pitches that are multiples of 8 bytes");
    }
    if (stage_check_device(&plan->st) != 0) return -1;
    const size_t span_in = input_pitch_bytes * ((size_t)plan->n - 1) + width * CPX;
    const size_t span_out = output_pitch_bytes * ((size_t)plan->n - 1) + width * CPX;
    if (!device_visible(pointer_kind(input), span_in) || !device_visible(pointer_kind(output), span_out)) {
        return set_error("columns: needs device-accessible (device or managed) pointers");
    }
    if (plan->ax_cols.n == 0 && axis_init(&plan->ax_cols, plan->n, plan->direction, plan->n > MCU_COLS_MAX_N ? AXIS_COLS : 0, plan->flags) != 0)
        return set_error("columns: could not build the tables of a %u-point column transform", plan->n);
    return run_cols_axis(&plan->ax_cols, &plan->scr_cols, output, output_pitch_bytes / CPX, 0, input, input_pitch_bytes / CPX, 0,
                         (unsigned)width, 1, tls_stream);
}

/* ---- convolution (ref fft.c:89-99, 565-624, 794-809) --------------------------------------------- */
struct mufft_plan_conv {
    mufft_plan_1d *plans[2];
    mufft_plan_1d *output_plan;
    size_t block_size;
    float normalization;
    unsigned n;
    int stereo;
    void *d_spectra;      /* mufft_execute_conv_batch beyond the engine limit: block spectra of one chunk */
    size_t spectra_bytes;
};

void mufft_free_plan_conv(mufft_plan_conv *plan)
{
    if (plan == NULL) return;
    mufft_free_plan_1d(plan->plans[0]);
    mufft_free_plan_1d(plan->plans[1]);
    mufft_free_plan_1d(plan->output_plan);
    mcu_free(plan->d_spectra);
    free(plan);
}

mufft_plan_conv *mufft_create_plan_conv(unsigned N, unsigned flags, unsigned method)
{
    if (!is_pow2_ge2(N)) return NULL;
    mufft_plan_conv *conv = calloc(1, sizeof(*conv));
    if (conv == NULL) return NULL;

    const unsigned zp0 = (method & MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_FIRST) ? MUFFT_FLAG_ZERO_PAD_UPPER_HALF : 0;
    const unsigned zp1 = (method & MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_SECOND) ? MUFFT_FLAG_ZERO_PAD_UPPER_HALF : 0;
    conv->n = N;
    conv->stereo = (method & 1u) == MUFFT_CONV_METHOD_FLAG_STEREO_MONO;
    if (conv->stereo) {
        conv->block_size = (size_t)N * CPX;
        conv->plans[0] = mufft_create_plan_1d_c2c(N, MUFFT_FORWARD, flags | zp0);
        conv->plans[1] = mufft_create_plan_1d_r2c(N, flags | zp1 | MUFFT_FLAG_FULL_R2C);
        conv->output_plan = mufft_create_plan_1d_c2c(N, MUFFT_INVERSE, flags & ~MUFFT_FLAG_ZERO_PAD_UPPER_HALF);
    } else {
        /* the reference pads the block so its SIMD multiply may overrun N/2+1 (fft.c:586) */
        conv->block_size = ((size_t)N / 2 + MUFFT_PADDING_COMPLEX_SAMPLES) * CPX;
        conv->plans[0] = mufft_create_plan_1d_r2c(N, flags | zp0);
        conv->plans[1] = mufft_create_plan_1d_r2c(N, flags | zp1);
        conv->output_plan = mufft_create_plan_1d_c2r(N, flags & ~MUFFT_FLAG_ZERO_PAD_UPPER_HALF);
    }
    conv->normalization = 1.0f / (float)N;
    if (conv->plans[0] == NULL || conv->plans[1] == NULL || conv->output_plan == NULL) {
        mufft_free_plan_conv(conv);
        return NULL;
    }
    return conv;
}

size_t mufft_conv_get_transformed_block_size(mufft_plan_conv *plan) { return plan->block_size; }

static int conv_input(mufft_plan_conv *plan, unsigned block, void *output, const void *input, size_t batch, int async)
{
    mufft_plan_1d *p = plan->plans[block & 1u];
    int in_mode, out_mode;
    plan_modes(p, &in_mode, &out_mode);
    return exec_1d(p, in_mode, out_mode, 1.0f, output, plan->block_size, input, 0, NULL, 0, 0, batch, async);
}

static int conv_output(mufft_plan_conv *plan, void *output, const void *first, const void *second,
                       size_t batch, size_t second_stride, int async)
{
    /* The spectral multiply and the 1/N happen in the load stage of the inverse transform. */
    const int in_mode = plan->stereo ? MCU_IN_MUL : MCU_IN_MUL_C2R;
    const size_t used = plan->stereo ? (size_t)plan->n * CPX : ((size_t)plan->n / 2 + 1) * CPX;
    return exec_1d(plan->output_plan, in_mode, MCU_OUT_COMPLEX, plan->normalization,
                   output, 0, first, plan->block_size, second, second_stride, used, batch, async);
}

void mufft_execute_conv_input(mufft_plan_conv *plan, unsigned block, void *output, const void *input)
{
    (void)conv_input(plan, block, output, input, 1, 0);
}

void mufft_execute_conv_output(mufft_plan_conv *plan, void *output, const void *input_first, const void *input_second)
{
    (void)conv_output(plan, output, input_first, input_second, 1, 0, 0);
}

int mufft_execute_conv_input_batch(mufft_plan_conv *plan, unsigned block, void *output, const void *input, size_t batch)
{
    return conv_input(plan, block, output, input, batch, 0);
}
int mufft_execute_conv_input_batch_async(mufft_plan_conv *plan, unsigned block, void *output, const void *input, size_t batch)
{
    return conv_input(plan, block, output, input, batch, 1);
}
int mufft_execute_conv_output_batch(mufft_plan_conv *plan, void *output, const void *input_first,
                                    const void *input_second, size_t batch, size_t second_stride_bytes)
{
    return conv_output(plan, output, input_first, input_second, batch, second_stride_bytes, 0);
}
int mufft_execute_conv_output_batch_async(mufft_plan_conv *plan, void *output, const void *input_first,
                                          const void *input_second, size_t batch, size_t second_stride_bytes)
{
    return conv_output(plan, output, input_first, input_second, batch, second_stride_bytes, 1);
}

/* ---- whole convolution in one call (additive; mufft_cuda.h) -------------------------------------
 * conv_input(FIRST) + conv_output per block without the block spectrum ever reaching HBM when the
 * transform fits the on-chip engine (one kernel: conv_kernel.cuh); beyond it the two library
 * steps run back to back through plan-owned scratch, chunk by chunk. */
typedef struct conv_ctx {
    mufft_plan_conv *plan;
} conv_ctx;

static int conv_unit_launch(void *ctx_, void *out, size_t out_stride, const void *in, size_t in_stride,
                            const void *in2, size_t in2_stride, size_t count, void *stream)
{
    mufft_plan_conv *plan = ((conv_ctx *)ctx_)->plan;
    mufft_plan_1d *fwd = plan->plans[0], *inv = plan->output_plan;
    int in_mode, out_mode;
    plan_modes(fwd, &in_mode, &out_mode);
    if (fwd->use_engine && fwd->n <= MCU_ENGINE_MAX_N) { /* the one-kernel convolution stops at MCU_ENGINE_MAX_N */
        mcu_conv_args a;
        memset(&a, 0, sizeof(a));
        a.out = out; a.in = in; a.filter = in2;
        a.tw_fwd = fwd->ax.d_engine_tw; a.tw_inv = inv->ax.d_engine_tw;
        a.rtw_fwd = fwd->d_rtw; a.rtw_inv = inv->d_rtw;
        a.out_stride = out_stride / CPX; a.in_stride = in_stride / CPX; a.filter_stride = in2_stride / CPX;
        a.batch = count; a.n = fwd->n; a.in_mode = in_mode; a.mono = !plan->stereo;
        a.norm = plan->normalization;
        return mcu_launch_conv(&a, stream);
    }
    /* long transforms: the spectra of a chunk of blocks (<= 256 MiB) go through plan-owned scratch */
    size_t chunk = ((size_t)256 << 20) / plan->block_size;
    if (chunk == 0) chunk = 1;
    if (chunk > count) chunk = count;
    TRY(grow(&plan->d_spectra, &plan->spectra_bytes, chunk * plan->block_size));
    const int mul_mode = plan->stereo ? MCU_IN_MUL : MCU_IN_MUL_C2R;
    for (size_t done = 0; done < count; done += chunk) {
        const size_t cnt = count - done < chunk ? count - done : chunk;
        TRY(run_rows(&fwd->ax, &fwd->scr, fwd->d_rtw, in_mode, out_mode, 1.0f, plan->d_spectra, plan->block_size / CPX,
                     (const char *)in + done * in_stride, in_stride / CPX, NULL, 0, cnt, stream, NULL));
        TRY(run_rows(&inv->ax, &inv->scr, inv->d_rtw, mul_mode, MCU_OUT_COMPLEX, plan->normalization,
                     (char *)out + done * out_stride, out_stride / CPX, plan->d_spectra, plan->block_size / CPX,
                     in2_stride ? (const char *)in2 + done * in2_stride : (const char *)in2, in2_stride / CPX, cnt, stream, NULL));
    }
    return 0;
}

static int conv_fused(mufft_plan_conv *plan, void *output, const void *input, const void *filter,
                      size_t batch, size_t filter_stride, int async)
{
    if (plan == NULL) return -1;
    conv_ctx ctx = { plan };
    int in_mode, out_mode;
    plan_modes(plan->plans[0], &in_mode, &out_mode);
    io_desc io;
    memset(&io, 0, sizeof(io));
    io.out = output; io.in = input; io.in2 = filter;
    io.in_unit = in_unit_bytes(plan->plans[0]->n, in_mode);
    io.out_unit = (size_t)plan->output_plan->n * CPX; /* stereo: N complex; mono: N floats */
    io.in2_unit = plan->stereo ? (size_t)plan->n * CPX : ((size_t)plan->n / 2 + 1) * CPX;
    io.in_stride = io.in_unit;
    io.out_stride = io.out_unit;
    io.in2_stride = filter_stride;
    io.count = batch;
    if (filter_stride % CPX != 0) {
        return set_error("strides must be multiples of 8 bytes");
    }
    return run_any(&plan->plans[0]->st, conv_unit_launch, &ctx, &io, async, 0);
}

int mufft_execute_conv_batch(mufft_plan_conv *plan, void *output, const void *input_first,
                             const void *filter_spectrum, size_t batch, size_t filter_stride_bytes)
{
    return conv_fused(plan, output, input_first, filter_spectrum, batch, filter_stride_bytes, 0);
}
int mufft_execute_conv_batch_async(mufft_plan_conv *plan, void *output, const void *input_first,
                                   const void *filter_spectrum, size_t batch, size_t filter_stride_bytes)
{
    return conv_fused(plan, output, input_first, filter_spectrum, batch, filter_stride_bytes, 1);
}

/* ---- partitioned convolution: filters longer than one block (additive; mufft_cuda.h) --------------
 * The reference leaves the accumulation over filter partitions to the caller (fft.h:170-174): P
 * conv_output calls and P-1 time-domain additions per block.  Here the P spectral products are
 * summed in the load stage of ONE inverse transform (frequency-domain delay line). */
static int conv_partitioned(mufft_plan_conv *plan, void *output, const void *spectra, const void *filter_spectra,
                            unsigned parts, size_t batch, int async)
{
    if (plan == NULL || parts == 0) {
        return set_error("partitioned convolution: needs a plan and at least one partition");
    }
    if (batch == 0) return 0;
    if (stage_check_device(&plan->output_plan->st) != 0) return -1;
    const size_t blk = plan->block_size;
    const size_t out_unit = (size_t)plan->output_plan->n * CPX;
    if ((((uintptr_t)output | (uintptr_t)spectra | (uintptr_t)filter_spectra) % CPX) != 0) {
        return set_error("buffers must be 8-byte aligned");
    }
    /* the history blocks in front of `spectra` are read too: classify the whole range */
    const char *first = (const char *)spectra - (size_t)(parts - 1) * blk;
    if (!device_visible(pointer_kind(first), (batch + parts - 1) * blk) ||
        !device_visible(pointer_kind(spectra), batch * blk) ||
        !device_visible(pointer_kind(filter_spectra), (size_t)parts * blk) ||
        !device_visible(pointer_kind(output), batch * out_unit)) {
        return set_error("partitioned convolution needs device-accessible (device or managed) pointers");
    }
    part_desc pd = { parts, -(long long)(blk / CPX), (long long)(blk / CPX) };
    mufft_plan_1d *inv = plan->output_plan;
    TRY(run_rows(&inv->ax, &inv->scr, inv->d_rtw, plan->stereo ? MCU_IN_MULACC : MCU_IN_MULACC_C2R, MCU_OUT_COMPLEX,
                 plan->normalization, output, out_unit / CPX, spectra, blk / CPX, filter_spectra, 0, batch, tls_stream, &pd));
    if (!async) TRY(mcu_stream_sync(tls_stream));
    return 0;
}

int mufft_execute_conv_output_partitioned_batch(mufft_plan_conv *plan, void *output, const void *spectra,
                                                const void *filter_spectra, unsigned parts, size_t batch)
{
    return conv_partitioned(plan, output, spectra, filter_spectra, parts, batch, 0);
}
int mufft_execute_conv_output_partitioned_batch_async(mufft_plan_conv *plan, void *output, const void *spectra,
                                                      const void *filter_spectra, unsigned parts, size_t batch)
{
    return conv_partitioned(plan, output, spectra, filter_spectra, parts, batch, 1);
}

/* Stand-alone spectral multiply (ref fft.h:201-207, kernel.c:22-34): runs the CUDA kernel on
 * whatever memory it is given and synchronises. */
static __thread stage tls_mul_stage; /* bound to the device it was first used on; re-created when the thread switches device */

static int mul_unit_launch(void *ctx, void *out, size_t out_stride, const void *in, size_t in_stride,
                           const void *in2, size_t in2_stride, size_t count, void *stream)
{
    (void)out_stride; (void)in_stride; (void)in2_stride; (void)count;
    mcu_resolve_args r;
    memset(&r, 0, sizeof(r));
    r.out = out; r.in = in; r.in2 = in2; r.batch = 1;
    r.n = *(unsigned *)ctx;
    r.mode = MCU_RES_MUL;
    r.norm = ((float *)ctx)[1];
    return mcu_launch_resolve(&r, stream);
}

static void mufft_convolve_cuda(void *output, const void *a, const void *b, float normalization, unsigned samples)
{
    if (samples == 0) return;
    union { unsigned u; float f; } ctx[2];
    ctx[0].u = samples;
    ctx[1].f = normalization;
    io_desc io;
    memset(&io, 0, sizeof(io));
    io.out = output; io.in = a; io.in2 = b;
    io.in_unit = io.out_unit = io.in2_unit = (size_t)samples * CPX;
    io.in_stride = io.out_stride = io.in_unit;
    io.in2_stride = io.in2_unit;
    io.count = 1;
    int dev = 0;
    if (mcu_get_device(&dev) == 0 && tls_mul_stage.device_plus_1 != dev + 1) {
        if (tls_mul_stage.device_plus_1 != 0) stage_free(&tls_mul_stage); /* streams and buffers of the other device */
        tls_mul_stage.device_plus_1 = dev + 1;
    }
    (void)run_any(&tls_mul_stage, mul_unit_launch, ctx, &io, 0, 0);
}

mufft_convolve_func mufft_get_convolve_func(unsigned flags)
{
    (void)flags;
    if (!have_device()) return NULL;
    return mufft_convolve_cuda;
}

/* ---- 2D plans (ref fft.c:69-86, 664-792, 852-989) --------------------------------------------------- */
struct mufft_plan_2d {
    unsigned nx, ny; /* nx = complex row length (Nx/2 for R2C / C2R) */
    int kind;
    unsigned vertical_nx;
    axis ax_x, ax_y;
    void *d_rtw;
    int cols_engine;
    int fused;          /* the fused persistent kernel does the whole transform (fft2d_kernel.cuh) */
    unsigned flags;     /* path flags of mufft_cuda.h */
    scratch scr_rows;   /* generic row passes */
    scratch scr_cols;   /* generic column passes */
    void *d_mid;        /* intermediate image(s) between the row and the column kernel */
    size_t mid_bytes;
    void *d_counters;   /* fused kernel: ticket counter + producer / consumer tiles done per image */
    size_t counters_bytes;
    stage st;
};

void mufft_free_plan_2d(mufft_plan_2d *plan)
{
    if (plan == NULL) return;
    axis_free(&plan->ax_x);
    axis_free(&plan->ax_y);
    mcu_free(plan->d_rtw);
    scratch_free(&plan->scr_rows);
    scratch_free(&plan->scr_cols);
    mcu_free(plan->d_mid);
    mcu_free(plan->d_counters);
    stage_free(&plan->st);
    free(plan);
}

static mufft_plan_2d *plan_2d_common(unsigned nx, unsigned ny, int direction, int kind, unsigned flags)
{
    if (!have_device()) return NULL;
    mufft_plan_2d *plan = calloc(1, sizeof(*plan));
    if (plan == NULL) return NULL;
    plan->nx = nx;
    plan->ny = ny;
    plan->kind = kind;
    plan->vertical_nx = nx;
    plan->cols_engine = ny <= MCU_COLS_MAX_N;
    plan->flags = flags & PATH_FLAGS;
    plan->fused = !(flags & MUFFT_CUDA_FLAG_SEPARATE_2D) && mcu_2d_fused_supported(nx, ny);
    /* C2R images whose rows have fewer than 512 complex points: 8 such rows are too little work per ticket of the fused
     * kernel, the column kernel + row kernel through scratch are faster (512 x 512: 1.03 against 1.42 ms per 1024 images;
     * from 512 complex points on the fused kernel wins or ties: profiles/README.md R2.5) */
    if (kind == KIND_C2R && nx < 512) plan->fused = 0;
    stage_bind_device(&plan->st);
    if (axis_init(&plan->ax_x, nx, direction, nx > onchip_limit(flags) ? AXIS_ROWS : 0, flags) != 0 ||
        axis_init(&plan->ax_y, ny, direction, plan->cols_engine ? 0 : AXIS_COLS, flags) != 0) {
        mufft_free_plan_2d(plan);
        return NULL;
    }
    return plan;
}

mufft_plan_2d *mufft_create_plan_2d_c2c(unsigned Nx, unsigned Ny, int direction, unsigned flags)
{
    if (!is_pow2_ge2(Nx) || !is_pow2_ge2(Ny)) return NULL;
    if (direction != MUFFT_FORWARD && direction != MUFFT_INVERSE) return NULL;
    return plan_2d_common(Nx, Ny, direction, KIND_C2C, flags);
}

mufft_plan_2d *mufft_create_plan_2d_r2c(unsigned Nx, unsigned Ny, unsigned flags)
{
    if (!is_pow2_ge2(Nx) || !is_pow2_ge2(Ny) || !is_pow2_ge2(Nx / 2)) return NULL;
    const int full = (flags & MUFFT_FLAG_FULL_R2C) != 0;
    mufft_plan_2d *plan = plan_2d_common(Nx / 2, Ny, MUFFT_FORWARD, full ? KIND_R2C_FULL : KIND_R2C, flags);
    if (plan == NULL) return NULL;
    plan->vertical_nx = full ? Nx : Nx / 2 + 1; /* ref fft.c:738-746 */
    plan->d_rtw = make_resolve_twiddles(MUFFT_FORWARD, Nx / 2);
    if (plan->d_rtw == NULL) {
        mufft_free_plan_2d(plan);
        return NULL;
    }
    return plan;
}

mufft_plan_2d *mufft_create_plan_2d_c2r(unsigned Nx, unsigned Ny, unsigned flags)
{
    if (!is_pow2_ge2(Nx) || !is_pow2_ge2(Ny) || !is_pow2_ge2(Nx / 2)) return NULL;
    mufft_plan_2d *plan = plan_2d_common(Nx / 2, Ny, MUFFT_INVERSE, KIND_C2R, flags);
    if (plan == NULL) return NULL;
    plan->vertical_nx = Nx / 2 + 1;
    plan->d_rtw = make_resolve_twiddles(MUFFT_INVERSE, Nx / 2);
    if (plan->d_rtw == NULL) {
        mufft_free_plan_2d(plan);
        return NULL;
    }
    return plan;
}

int mufft_cuda_plan_2d_paths(const mufft_plan_2d *plan)
{
    if (plan == NULL) return 0;
    return (plan->ax_x.d_engine_tw != NULL || plan->ax_x.long_n2 != 0 ? 1 : 0) | (plan->cols_engine ? 2 : 0) | (plan->fused ? 4 : 0);
}

/* Column transforms of `images` images on device-visible memory; pitches in complex elements. */
static int run_cols_axis(const axis *ax, scratch *scr_cols, void *out, size_t out_pitch, size_t out_image,
                         const void *in, size_t in_pitch, size_t in_image, unsigned width, size_t images, void *stream)
{
    if (ax->n <= MCU_COLS_MAX_N) {
        mcu_cols_args a;
        memset(&a, 0, sizeof(a));
        a.out = out; a.in = in; a.tw = ax->d_engine_tw;
        a.out_pitch = out_pitch; a.in_pitch = in_pitch; a.out_image = out_image; a.in_image = in_image;
        a.batch = images; a.n = ax->n; a.width = width; a.direction = ax->direction;
        return mcu_launch_cols(&a, stream);
    }
    if (ax->long_cols) {
        const unsigned blocks = (width + MUFFT_LONGCOLS_CB - 1) / MUFFT_LONGCOLS_CB;
        const unsigned lag = mcu_longcols_default_lag(ax->n);
        size_t ring = (size_t)lag + 3;
        if (ring > blocks) ring = blocks;
        if (lag == 0 || ring == 0) return set_error("four-step column kernel: no resident CTA on this device");
        TRY(grow(&scr_cols->slots, &scr_cols->slots_bytes, ring * mcu_longcols_slot_elems(ax->n) * CPX));
        TRY(grow(&scr_cols->counters, &scr_cols->counters_bytes, (2 * (size_t)blocks + 1) * sizeof(unsigned)));
        for (size_t img = 0; img < images; img++) {
            mcu_longcols_args la;
            memset(&la, 0, sizeof(la));
            TRY(mcu_memset(scr_cols->counters, 0, (2 * (size_t)blocks + 1) * sizeof(unsigned), stream));
            la.out = (char *)out + img * out_image * CPX; la.in = (const char *)in + img * in_image * CPX;
            la.tw = ax->d_tw_n2; la.roots = ax->d_roots; la.scratch = scr_cols->slots; la.counters = scr_cols->counters;
            la.ring = (unsigned)ring; la.lag = lag; la.out_pitch = out_pitch; la.in_pitch = in_pitch; la.width = width;
            la.n = ax->n; la.n1 = ax->n1; la.direction = ax->direction;
            TRY(mcu_launch_longcols(&la, stream));
        }
        return 0;
    }
    if (ax->n2 != 0) {
        /* Long columns, n = n1 * n2 (row y = y1*n2 + y2, output row k = k1 + n1*k2), two sweeps of
         * the column kernel per image:
         *   A  n2 "images" (one per y2) of n1-point column transforms -> scratch [k1][y2][c]
         *   B  n1 "images" (one per k1) of n2-point column transforms with the twiddle
         *      w_n^(k1*y2) fused into the load -> out rows k1 + n1*k2 */
        const unsigned n1 = ax->n1, n2 = ax->n2;
        /* blocks of columns whose scratch ([k1][y2][block]) stays in L2 between the two sweeps */
        size_t wb = COLS_BLOCK_BYTES / ((size_t)ax->n * CPX);
        wb = wb < 16 ? 16 : wb - wb % 16;
        if (wb > width) wb = width;
        TRY(scratch_reserve(scr_cols, wb * ax->n * CPX));
        for (size_t img = 0; img < images; img++) {
            for (size_t c0 = 0; c0 < width; c0 += wb) {
                const unsigned w = (unsigned)(width - c0 < wb ? width - c0 : wb);
                mcu_cols_args ca;
                memset(&ca, 0, sizeof(ca));
                ca.in = (const char *)in + (img * in_image + c0) * CPX; ca.in_pitch = (size_t)n2 * in_pitch; ca.in_image = in_pitch;
                ca.out = scr_cols->buf[0]; ca.out_pitch = (size_t)n2 * w; ca.out_image = w;
                ca.tw = ax->d_tw_n1; ca.batch = n2; ca.n = n1; ca.width = w; ca.direction = ax->direction;
                TRY(mcu_launch_cols(&ca, stream));
                mcu_cols_args cb;
                memset(&cb, 0, sizeof(cb));
                cb.in = scr_cols->buf[0]; cb.in_pitch = w; cb.in_image = (size_t)n2 * w;
                cb.out = (char *)out + (img * out_image + c0) * CPX; cb.out_pitch = (size_t)n1 * out_pitch; cb.out_image = out_pitch;
                cb.tw = ax->d_tw_n2; cb.batch = n1; cb.n = n2; cb.width = w; cb.direction = ax->direction;
                cb.roots = ax->d_roots; cb.root_mask = ax->n - 1;
                cb.discard_input = w % 16 == 0; /* whole 128-byte segments of plan-owned scratch */
                TRY(mcu_launch_cols(&cb, stream));
            }
        }
        return 0;
    }
    /* Generic vertical passes, image by image, ping-ponging through scratch of pitch `width`. */
    TRY(scratch_reserve(scr_cols, (size_t)width * ax->n * CPX));
    for (size_t img = 0; img < images; img++) {
        const void *cur = (const char *)in + img * in_image * CPX;
        size_t cur_pitch = in_pitch;
        for (unsigned s = 0; s < ax->nsteps; s++) {
            const int last = s + 1 == ax->nsteps;
            mcu_pass_args p;
            memset(&p, 0, sizeof(p));
            p.in = cur; p.in_stride = cur_pitch;
            if (last) { p.out = (char *)out + img * out_image * CPX; p.out_stride = out_pitch; }
            else { p.out = scr_cols->buf[s & 1]; p.out_stride = width; }
            p.roots = ax->d_roots; p.batch = 1; p.n = ax->n; p.width = width;
            p.radix = ax->radix[s]; p.p = ax->span[s]; p.direction = ax->direction;
            TRY(mcu_launch_pass(&p, stream));
            cur = p.out;
            cur_pitch = p.out_stride;
        }
    }
    return 0;
}

static int run_cols(mufft_plan_2d *plan, void *out, size_t out_pitch, size_t out_image,
                    const void *in, size_t in_pitch, size_t in_image, unsigned width, size_t images, void *stream)
{
    return run_cols_axis(&plan->ax_y, &plan->scr_cols, out, out_pitch, out_image, in, in_pitch, in_image, width, images, stream);
}

/* Whole 2D transforms of `images` tightly packed images on device-visible memory.  The
 * intermediate between the row and the column kernel is plan-owned and sized for a small
 * group of images, so that on B200 it is produced and consumed inside the 126 MB L2 instead
 * of making a round trip through HBM. */

static int run_2d_device(mufft_plan_2d *plan, void *out, const void *in, size_t images, void *stream)
{
    const unsigned nx = plan->nx, ny = plan->ny;
    if (plan->fused) {
        /* one persistent launch for the whole batch: the intermediate lives in `out` (C2C, R2C) or in a
         * plan-owned ring of a few images (C2R) and stays in L2 */
        static const int kind_of[4] = { MCU_2D_C2C, MCU_2D_R2C, MCU_2D_R2C_FULL, MCU_2D_C2R }; /* KIND_* order */
        mcu_2d_args f;
        memset(&f, 0, sizeof(f));
        f.kind = kind_of[plan->kind];
        TRY(grow(&plan->d_counters, &plan->counters_bytes, (2 * images + 1) * sizeof(unsigned)));
        TRY(mcu_memset(plan->d_counters, 0, (2 * images + 1) * sizeof(unsigned), stream));
        f.out = out; f.in = in; f.tw_y = plan->ax_y.d_engine_tw; f.rtw = plan->d_rtw;
        f.tw_x = mcu_2d_rows_variant(nx, f.kind) != 0 ? plan->ax_x.d_engine_tw_real : plan->ax_x.d_engine_tw;
        f.counters = plan->d_counters;
        f.in_image = plan->kind == KIND_C2R ? 2 * (size_t)nx * ny : (size_t)nx * ny;
        f.out_image = (plan->kind == KIND_R2C || plan->kind == KIND_R2C_FULL) ? 2 * (size_t)nx * ny : (size_t)nx * ny;
        f.batch = images; f.nx = nx; f.ny = ny; f.direction = plan->ax_x.direction;
        if (plan->kind == KIND_C2R) {
            /* columns first into the ring (ref fft.c:861-916 uses the output buffer, which fft.h therefore wants
             * twice as large; the batched API has dense outputs, so the ring is the plan's) */
            f.lag = mcu_2d_default_lag(nx, ny, f.kind);
            size_t ring = (size_t)f.lag + 3;
            if (ring > images) ring = images;
            TRY(grow(&plan->d_mid, &plan->mid_bytes, ring * mcu_2d_mid_elems(nx, ny) * CPX));
            f.mid = plan->d_mid;
            f.ring = (unsigned)ring;
        }
        return mcu_launch_2d_fused(&f, stream);
    }
    const int real_rows = plan->kind != KIND_C2C;
    const size_t mid_pitch = real_rows ? 2 * (size_t)nx : nx; /* ref fft.c:920 doubles the pitch for R2C */
    const size_t mid_image = mid_pitch * ny;                  /* complex elements */
    size_t group = MID_GROUP_BYTES / (mid_image * CPX);
    if (group == 0) group = 1;
    if (group > images) group = images;
    TRY(grow(&plan->d_mid, &plan->mid_bytes, group * mid_image * CPX));

    for (size_t done = 0; done < images; done += group) {
        const size_t cnt = images - done < group ? images - done : group;
        if (plan->kind == KIND_C2R) {
            /* columns first (ref fft.c:863-916): input pitch 2*nx, nx+1 columns carry data */
            const char *src = (const char *)in + done * mid_image * CPX;
            char *dst = (char *)out + done * (size_t)nx * ny * CPX;
            TRY(run_cols(plan, plan->d_mid, mid_pitch, mid_image, src, mid_pitch, mid_image, nx + 1, cnt, stream));
            /* rows: C2R resolve fused into the load, result dense (pitch nx complex = Nx floats) */
            /* (the rows of all images of the group form one batch: image stride = ny rows) */
            TRY(run_rows(&plan->ax_x, &plan->scr_rows, plan->d_rtw, MCU_IN_C2R, MCU_OUT_COMPLEX, 1.0f,
                         dst, nx, plan->d_mid, mid_pitch, NULL, 0, cnt * ny, stream, NULL));
        } else {
            const size_t in_image = (size_t)nx * ny; /* complex elements (real rows: Nx floats = nx complex) */
            const char *src = (const char *)in + done * in_image * CPX;
            char *dst = (char *)out + done * mid_image * CPX;
            const int out_mode = plan->kind == KIND_R2C ? MCU_OUT_R2C : (plan->kind == KIND_R2C_FULL ? MCU_OUT_R2C_FULL : MCU_OUT_COMPLEX);
            TRY(run_rows(&plan->ax_x, &plan->scr_rows, plan->d_rtw, MCU_IN_COMPLEX, out_mode, 1.0f,
                         plan->d_mid, mid_pitch, src, nx, NULL, 0, cnt * ny, stream, NULL));
            TRY(run_cols(plan, dst, mid_pitch, mid_image, plan->d_mid, mid_pitch, mid_image, plan->vertical_nx, cnt, stream));
        }
    }
    return 0;
}

static int image_unit_launch(void *ctx, void *out, size_t out_stride, const void *in, size_t in_stride,
                             const void *in2, size_t in2_stride, size_t count, void *stream)
{
    (void)in2; (void)in2_stride; (void)out_stride; (void)in_stride; /* images are tightly packed */
    return run_2d_device(ctx, out, in, count, stream);
}

static int exec_2d(mufft_plan_2d *plan, void *output, const void *input, size_t batch, int async)
{
    const size_t nx = plan->nx, ny = plan->ny;
    io_desc io;
    memset(&io, 0, sizeof(io));
    io.out = output; io.in = input;
    switch (plan->kind) {
    case KIND_C2C: io.in_unit = io.out_unit = nx * ny * CPX; break;
    case KIND_C2R: io.in_unit = 2 * nx * ny * CPX; io.out_unit = nx * ny * CPX; break;
    default: io.in_unit = nx * ny * CPX; io.out_unit = 2 * nx * ny * CPX; break;
    }
    io.in_stride = io.in_unit;
    io.out_stride = io.out_unit;
    io.count = batch;
    return run_any(&plan->st, image_unit_launch, plan, &io, async, 0);
}

void mufft_execute_plan_2d(mufft_plan_2d *plan, void *MUFFT_RESTRICT output, const void *MUFFT_RESTRICT input)
{
    (void)exec_2d(plan, output, input, 1, 0);
}

int mufft_execute_plan_2d_batch(mufft_plan_2d *plan, void *output, const void *input, size_t batch)
{
    return exec_2d(plan, output, input, batch, 0);
}

int mufft_execute_plan_2d_batch_async(mufft_plan_2d *plan, void *output, const void *input, size_t batch)
{
    return exec_2d(plan, output, input, batch, 1);
}

#include "mufft_slab.inc"
