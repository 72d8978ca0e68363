/* Internal C ABI between the C host layer (mufft_host.c) and the CUDA translation units.
 * Plain pointers and sizes only; every function returns 0 on success or a cudaError_t value.
 * Nothing here is exported to users: the public surface is include/fft.h + include/mufft_cuda.h.
 */
#ifndef MUFFT_B200_ABI_H
#define MUFFT_B200_ABI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Largest transform length handled by the single-kernel shared-memory engine. */
#define MCU_ENGINE_MAX_N 8192
/* Row kernels are also built for this length (one 512-thread CTA per transform, one CTA per SM); plans use them
 * unless created with MUFFT_CUDA_FLAG_TWO_KERNEL / _GENERIC_PASSES / _FOUR_STEP_16K.  Convolution in one kernel and the
 * column kernels stop at MCU_ENGINE_MAX_N / MCU_COLS_MAX_N. */
#define MCU_ENGINE_BIG_N 16384
/* Largest column length handled by the shared-memory column kernel. */
#define MCU_COLS_MAX_N 2048

/* How a row transform reads its input (first pass prologue). */
enum {
    MCU_IN_COMPLEX = 0,   /* x[i] = in[i]                                   (C2C, and R2C: reals viewed as N/2 complex) */
    MCU_IN_HALF = 1,      /* x[i] = i < N/2 ? in[i] : 0, upper half never read (ref kernel.c:104-122,200-226,333-387) */
    MCU_IN_C2R = 2,       /* x[k] = (X[k]+conj X[N-k]) + (X[k]-conj X[N-k])*rtw[k]   (ref kernel.c:36-45) */
    MCU_IN_MUL = 3,       /* x[i] = norm * in[i] * in2[i]                            (ref kernel.c:22-34) */
    MCU_IN_MUL_C2R = 4,   /* X = norm * in .* in2, then as MCU_IN_C2R */
    /* Partitioned convolution (frequency-domain delay line): the product is summed over `parts`
     * (spectrum, filter-partition) pairs before the transform, see mcu_rows_args.parts */
    MCU_IN_MULACC = 5,    /* x[i] = norm * sum_p in[p*in_part_stride + i] * in2[p*in2_part_stride + i] */
    MCU_IN_MULACC_C2R = 6 /* X = that sum, then as MCU_IN_C2R */
};
/* How a row transform writes its output (last pass epilogue). */
enum {
    MCU_OUT_COMPLEX = 0,  /* out[i] = y[i] */
    MCU_OUT_R2C = 1,      /* N+1 bins from the N-point half-length spectrum  (ref kernel.c:47-63) */
    MCU_OUT_R2C_FULL = 2, /* all 2N bins                                      (ref kernel.c:65-82) */
    MCU_OUT_PEERS = 3,    /* out[i] = y[i], scattered over the column slabs of several devices (mcu_peer_dest) */
    MCU_OUT_R2C_PEERS = 4 /* the N+1 bins of MCU_OUT_R2C scattered the same way; bin N rides with the last device */
};

/* Destination of a row transform whose output is the exchange step of a slab-decomposed 2D
 * transform: element x of transform b goes to device p = x >> shift, into
 * base[p] + (row0 + b) * pitch + (x & ((1 << shift) - 1)), all in complex elements.  The bases are
 * device pointers valid in this process (local memory or peer memory mapped through CUDA IPC),
 * so the stores of the row kernel ARE the all-to-all transpose: they travel over NVLink. */
#define MCU_MAX_PEERS 16
typedef struct mcu_peer_dest {
    void *base[MCU_MAX_PEERS];
    size_t row0, pitch;
    unsigned shift;       /* log2(columns per device) */
    unsigned last;        /* MCU_OUT_R2C_PEERS: index of the last device, which also takes bin N (column 2^shift of its slab) */
    size_t pitch_last;    /* MCU_OUT_R2C_PEERS: row pitch of the last device's slab (pitch + 1 when tight) */
} mcu_peer_dest;

typedef struct mcu_rows_args {
    void *out;
    const void *in;
    const void *in2;       /* second operand of the spectral multiply, else NULL */
    const void *tw;        /* engine twiddle table for n (device) */
    const void *rtw;       /* resolve twiddles, n entries (device), or NULL */
    size_t out_stride;     /* distance between consecutive transforms, in complex elements */
    size_t in_stride;
    size_t in2_stride;     /* 0 = every transform multiplies by the same spectrum */
    size_t batch;
    unsigned n;            /* complex transform length */
    int direction;         /* -1 forward, +1 inverse */
    int in_mode, out_mode;
    float norm;
    unsigned parts;        /* MCU_IN_MULACC*: number of (spectrum, filter partition) pairs summed per transform */
    long long in_part_stride;  /* complex elements from the spectrum of pair p to that of pair p+1 (may be negative) */
    long long in2_part_stride; /* same for the filter partitions */
    mcu_peer_dest peers;   /* out_mode == MCU_OUT_PEERS only */
} mcu_rows_args;

typedef struct mcu_cols_args {
    void *out;
    const void *in;
    const void *tw;        /* engine twiddle table for n */
    size_t out_pitch;      /* complex elements between consecutive rows */
    size_t in_pitch;
    size_t out_image;      /* complex elements between consecutive images of a batch */
    size_t in_image;
    size_t batch;
    unsigned n;            /* column length (rows) */
    unsigned width;        /* number of columns to transform */
    int direction;
    unsigned rows_valid;   /* rows >= rows_valid are taken as zero and never read (0 = all n rows valid) */
    const void *roots;     /* NULL, or exp(dir*2*pi*i*j/n_total): element (image, row) is multiplied by */
    unsigned root_mask;    /* roots[(image*row) & root_mask] before the transform (n = 256, 512, 1024 only) */
    int discard_input;     /* 1 = the input is scratch that nobody reads again: its lines are dropped from L2 once they
                            * have been read, without a write-back (needs 16-column tiles of 128-byte aligned row segments) */
} mcu_cols_args;

/* Fused convolution kernel (conv_kernel.cuh): forward transform, spectral multiply, inverse
 * transform of every block in one launch; n <= MCU_ENGINE_MAX_N only. */
typedef struct mcu_conv_args {
    void *out;             /* stereo: n complex per block; mono: 2n floats (= n complex) per block */
    const void *in;        /* stereo: n (n/2 when zero padded) complex; mono: 2n (n) floats */
    const void *filter;    /* filter spectrum: stereo n bins, mono n+1 bins */
    const void *tw_fwd, *tw_inv;   /* engine tables (variant 0) of n, forward and inverse */
    const void *rtw_fwd, *rtw_inv; /* mono: resolve twiddles of both directions, n entries each */
    size_t out_stride, in_stride, filter_stride; /* complex elements between blocks (filter_stride 0 = one filter) */
    size_t batch;
    unsigned n;            /* complex transform length (stereo: N, mono: N/2) */
    int in_mode;           /* MCU_IN_COMPLEX or MCU_IN_HALF */
    int mono;
    float norm;
} mcu_conv_args;

/* Second kernel of the two-kernel transform for n = n1 * n2 > MCU_ENGINE_MAX_N (rowst_kernel.cuh). */
typedef struct mcu_rowst_args {
    void *out;             /* X[k1 + n1*k2] */
    const void *in;        /* Y[k1][n2], rows contiguous */
    const void *tw;        /* engine twiddle table of n2 */
    const void *roots;     /* exp(dir*2*pi*i*j/n), j < n */
    size_t out_batch, in_batch; /* complex elements between consecutive transforms */
    size_t batch;
    unsigned n, n1, n2;
    int direction;
    int to_peers;          /* 1 = scatter X over the devices of `peers` instead of writing `out` */
    mcu_peer_dest peers;
} mcu_rowst_args;

/* Fused 2D kernel (fft2d_kernel.cuh): whole batch in one persistent launch. */
enum { MCU_2D_C2C = 0, MCU_2D_R2C = 1, MCU_2D_R2C_FULL = 2, MCU_2D_C2R = 3 };
typedef struct mcu_2d_args {
    void *out;             /* C2C: [batch][ny][nx] complex, also holds the row-transformed intermediate; R2C: rows of pitch
                            * 2 nx bins; C2R: [batch][ny][nx] complex = 2 nx floats per row */
    const void *in;        /* C2C: as out; R2C: 2 nx floats per row; C2R: spectrum rows of pitch 2 nx bins, nx + 1 carry data */
    const void *tw_x, *tw_y; /* engine twiddle tables of nx (the variant of mcu_2d_rows_variant) and ny (variant 0) */
    const void *rtw;       /* R2C / C2R: resolve twiddles of the rows, nx entries */
    int kind;              /* MCU_2D_*; nx = complex points per row transform (half the real row length) */
    unsigned *counters;    /* 1 + 2 * batch zero-initialised words: ticket counter, producer / consumer tiles done per image */
    void *mid;             /* C2R: ring of `ring` intermediate images of mcu_2d_mid_elems(nx, ny) complex each */
    unsigned ring;
    size_t out_image, in_image; /* complex elements between consecutive images */
    size_t batch;
    unsigned nx, ny;
    int direction;
    unsigned lag;          /* consumer tickets run this many images behind the producer tickets (0 = the launcher's choice) */
} mcu_2d_args;

/* Single-launch four-step kernel for rows of n = n1 * n2 points (long_kernel.cuh): both steps in one
 * persistent CTA, the intermediate in a scratch slot of n complex per CTA that stays in L2. */
typedef struct mcu_long_args {
    void *out;             /* X[k], natural order */
    const void *in;
    const void *tw;        /* engine twiddle table of n2 (variant 0) */
    const void *roots;     /* exp(dir*2*pi*i*j/n), j < n */
    void *scratch;         /* scratch_slots slots of n complex each */
    size_t scratch_slots;  /* at most this many CTAs are launched */
    size_t out_batch, in_batch; /* complex elements between consecutive transforms */
    size_t batch;
    unsigned n, n1, n2;
    int direction;
    int in_mode;           /* MCU_IN_COMPLEX or MCU_IN_HALF */
    int to_peers;          /* 1 = scatter X over the devices of `peers` instead of writing `out` */
    mcu_peer_dest peers;
} mcu_long_args;

/* Single-launch four-step COLUMN transforms of n = n1 x 1024 points down a row-major [n][width] matrix
 * (longcols_kernel.cuh): blocks of 32 columns, intermediate in a ring of L2-resident scratch slots. */
#ifndef MUFFT_LONGCOLS_CB
#define MUFFT_LONGCOLS_CB 32 /* columns per block of the four-step column kernel (a multiple of 16) */
#endif
typedef struct mcu_longcols_args {
    void *out;             /* may equal `in` */
    const void *in;
    const void *tw;        /* engine twiddle table of 1024 points (variant 0) */
    const void *roots;     /* exp(dir*2*pi*i*j/n), j < n */
    void *scratch;         /* `ring` slots of mcu_longcols_slot_elems(n) complex each */
    unsigned *counters;    /* 1 + 2 * ceil(width / 32) zero-initialised words */
    unsigned ring;         /* > lag, or >= the number of column blocks */
    unsigned lag;          /* B tiles run this many blocks behind the A tiles (0 = the launcher's choice) */
    size_t out_pitch, in_pitch; /* complex elements between consecutive rows */
    unsigned width;        /* columns */
    unsigned n, n1;
    int direction;
} mcu_longcols_args;

/* Generic one-pass Stockham kernel (any length; used beyond the engine limits). */
typedef struct mcu_pass_args {
    void *out;
    const void *in;
    const void *roots;     /* master table exp(dir*2*pi*i*j/n), j < n (device) */
    size_t out_stride;     /* complex elements between consecutive points of one transform */
    size_t in_stride;
    size_t out_batch;      /* complex elements between consecutive batch entries */
    size_t in_batch;
    size_t batch;
    unsigned n;            /* transform length */
    unsigned width;        /* independent transforms side by side (columns); 1 for 1D */
    unsigned radix, p;
    int direction;
    int half;              /* first pass of a zero-padded transform */
} mcu_pass_args;

/* Stand-alone resolve / multiply kernel (generic path and mufft_get_convolve_func). */
enum { MCU_RES_R2C = 0, MCU_RES_R2C_FULL = 1, MCU_RES_C2R = 2, MCU_RES_MUL = 3, MCU_RES_MUL_C2R = 4 };

typedef struct mcu_resolve_args {
    void *out;
    const void *in;
    const void *in2;       /* second spectrum for the MUL modes, else NULL */
    const void *rtw;       /* resolve twiddles, n entries (unused by MCU_RES_MUL) */
    size_t out_stride, in_stride, in2_stride, batch; /* strides in complex elements */
    unsigned n;            /* complex length (MCU_RES_MUL: number of samples, any value) */
    int mode;              /* MCU_RES_* */
    float norm;
} mcu_resolve_args;

/* ---- runtime ---------------------------------------------------------------------------- */
int mcu_device_count(int *count);
int mcu_set_device(int device);
int mcu_get_device(int *device);
int mcu_sm_count(int *count);
const char *mcu_error_string(int err);

/* CUDA IPC: 64-byte handle of a cudaMalloc allocation / mapping of another process's allocation */
int mcu_ipc_export(void *handle64, void *ptr);
int mcu_ipc_import(void **ptr, const void *handle64);
int mcu_ipc_close(void *ptr);
int mcu_malloc(void **ptr, size_t bytes);
int mcu_free(void *ptr);
int mcu_malloc_managed(void **ptr, size_t bytes);
int mcu_malloc_pinned(void **ptr, size_t bytes);
int mcu_free_pinned(void *ptr);
int mcu_memset(void *ptr, int value, size_t bytes, void *stream);

/* 0 = ordinary host memory, 1 = pinned host, 2 = device, 3 = managed */
enum { MCU_MEM_HOST = 0, MCU_MEM_PINNED = 1, MCU_MEM_DEVICE = 2, MCU_MEM_MANAGED = 3 };
int mcu_pointer_kind(const void *ptr);

int mcu_stream_create(void **stream);
int mcu_stream_destroy(void *stream);
int mcu_stream_sync(void *stream);
int mcu_event_create(void **event);
int mcu_event_destroy(void *event);
int mcu_event_record(void *event, void *stream);
int mcu_stream_wait_event(void *stream, void *event);
int mcu_event_sync(void *event);
int mcu_event_create_timing(void **event);                 /* an event that records time stamps */
int mcu_event_elapsed_ms(float *ms, void *start, void *end); /* both recorded on streams of the same device */
/* lets kernels and copies of the current device address memory of `peer_device` (NVLink / PCIe peer to peer);
 * already enabled = success */
int mcu_enable_peer_access(int peer_device);
/* CUDA graphs for the latency path: everything enqueued on `stream` between begin and end becomes an executable
 * graph (*exec); replaying it costs one driver call however many nodes it has. */
int mcu_capture_begin(void *stream);
int mcu_capture_end(void *stream, void **exec); /* also ends a capture that failed; *exec = NULL then */
int mcu_graph_launch(void *exec, void *stream);
int mcu_graph_destroy(void *exec);
int mcu_copy_async(void *dst, const void *src, size_t bytes, void *stream);
int mcu_copy2d_async(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes, size_t rows, void *stream);
int mcu_prefetch(const void *ptr, size_t bytes, int to_device, void *stream);

/* ---- kernels ------------------------------------------------------------------------------ */
int mcu_launch_rows(const mcu_rows_args *args, void *stream);
/* Single-launch four-step rows: n2 of the split the kernel uses for n (0 = no such kernel for n), the
 * number of CTAs it keeps resident on the current device (= scratch slots worth allocating), the launch. */
unsigned mcu_long_n2(unsigned n);
unsigned mcu_long_ctas(unsigned n);
int mcu_launch_long(const mcu_long_args *args, void *stream);
/* Single-launch four-step columns: 1 if the kernel exists for n rows; complex elements of one ring slot; the lag the
 * launcher would pick on the current device (the ring must hold more slots); the launch. */
int mcu_longcols_supported(unsigned n);
size_t mcu_longcols_slot_elems(unsigned n);
unsigned mcu_longcols_default_lag(unsigned n);
int mcu_launch_longcols(const mcu_longcols_args *args, void *stream);
int mcu_launch_cols(const mcu_cols_args *args, void *stream);
int mcu_launch_pass(const mcu_pass_args *args, void *stream);
/* n2 of the two-kernel split of n for a row (columns = 0) or column (columns = 1) transform
 * (0 = n is outside its range and takes the generic passes) */
unsigned mcu_split_n2(unsigned n, int columns);
int mcu_launch_rowst(const mcu_rowst_args *args, void *stream);
/* 1 if the fused 2D kernel exists for this image shape (nx = complex length of a row transform); the engine
 * shape variant its row transforms use for a kind (selects the table passed as tw_x); the lag the launcher
 * would pick on the current device (C2R: the ring must hold more than lag images); elements of a C2R ring slot */
int mcu_2d_fused_supported(unsigned nx, unsigned ny);
int mcu_2d_rows_variant(unsigned nx, int kind);
unsigned mcu_2d_default_lag(unsigned nx, unsigned ny, int kind);
size_t mcu_2d_mid_elems(unsigned nx, unsigned ny);
int mcu_launch_2d_fused(const mcu_2d_args *args, void *stream);
int mcu_launch_resolve(const mcu_resolve_args *args, void *stream);
int mcu_launch_conv(const mcu_conv_args *args, void *stream);
/* Engine shape variant the row kernel of (n, in_mode, out_mode) uses: 0 general, 1 / 2 the
 * mirrored-pass shapes of the real transforms (fft_engine.cuh).  The twiddle table passed in
 * mcu_rows_args.tw must be the one of that variant. */
int mcu_rows_variant(unsigned n, int in_mode, int out_mode);
/* number of complex entries of the engine twiddle table for length n, and its host fill */
size_t mcu_engine_table_size(unsigned n, int variant);
void mcu_engine_table_fill(float *table, unsigned n, int direction, int variant);
/* number of kernel launches issued through mcu_launch_* since the last reset (for bench.py) */
unsigned long long mcu_launch_count(int reset);
void mcu_launch_count_add(unsigned long long n); /* launches replayed from a graph */

#ifdef __cplusplus
}
#endif
#endif
