/* Additive GPU extension of the muFFT API (include/fft.h stays exactly the reference surface).
 *
 * fft.h has no notion of batch, stream or device (ref fft.h:86-121): one call = one transform,
 * synchronous.  A B200 is only busy when thousands of transforms are in flight, so the entry
 * points below run the SAME plans over a batch of independent transforms with one kernel
 * launch, optionally asynchronously on a caller-provided CUDA stream.
 *
 * Memory: every pointer may be ordinary host memory, pinned host memory, CUDA managed memory
 * (what mufft_alloc returns) or device memory.  Host memory is staged through plan-owned device
 * buffers in pipelined chunks; device/managed memory is used in place.
 *
 * Strides are in BYTES between consecutive transforms; 0 selects the tight default:
 *   C2C      in N*8 (N/2*8 with ZERO_PAD_UPPER_HALF)   out N*8
 *   R2C      in N*4 (N/2*4 zero padded)                out (N/2+1)*8, or N*8 with FULL_R2C
 *   C2R      in (N/2+1)*8                              out N*4
 * Strides must be multiples of 8 bytes.  Buffers should be 8-byte aligned (mufft_alloc gives 256): a
 * buffer that is not - a float array with 4-byte alignment on the real side of R2C / C2R - is never
 * used in place but staged like host memory (synchronous calls only; the async calls refuse it).
 */
#ifndef MUFFT_CUDA_H__
#define MUFFT_CUDA_H__

#include "fft.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Number of visible CUDA devices (0 when there is none or the driver is missing). */
int mufft_cuda_device_count(void);
/* Device used by plans created afterwards on this thread, and by mufft_alloc. 0 on success. */
int mufft_cuda_set_device(int device);
/* CUDA stream (a cudaStream_t passed as void*) used by executes issued from this thread.
 * NULL (the default) is the legacy default stream. */
void mufft_cuda_set_stream(void *stream);
void *mufft_cuda_get_stream(void);
/* Blocks until the current stream has drained.  Returns 0 or a CUDA error code. */
int mufft_cuda_synchronize(void);
/* Message of the last failure on this thread ("" if none).  Never NULL. */
const char *mufft_cuda_last_error(void);
/* Kernel launches issued by the library since the previous call with reset != 0. */
unsigned long long mufft_cuda_launch_count(int reset);

/* `batch` independent 1D transforms with the plan's kind; synchronous (returns when `output`
 * is complete and host-visible), any memory kind. */
int mufft_execute_plan_1d_batch(mufft_plan_1d *plan, void *output, const void *input, size_t batch,
                                size_t output_stride_bytes, size_t input_stride_bytes);
/* Same, but only enqueues work on the current stream; both pointers must be device-accessible
 * (device or managed memory).  Returns 0, or nonzero if the pointers are not device-accessible
 * or a launch failed. */
int mufft_execute_plan_1d_batch_async(mufft_plan_1d *plan, void *output, const void *input, size_t batch,
                                      size_t output_stride_bytes, size_t input_stride_bytes);

/* The plan's N-point C2C transform applied down the COLUMNS of a row-major [N][width] complex
 * matrix (pitches in bytes between consecutive rows), device-accessible pointers, enqueued on the
 * current stream.  This is the column pass of a slab-decomposed 2D transform (mufft_b200/slab.py). */
int mufft_execute_plan_1d_columns(mufft_plan_1d *plan, void *output, const void *input, size_t width,
                                  size_t output_pitch_bytes, size_t input_pitch_bytes);

/* Row pass and exchange step of a slab-decomposed 2D transform in ONE kernel: `batch` rows are
 * transformed with the plan (plain C2C, length N) and element x of row b is stored to device
 * p = x / (N / peers), at peer_outputs[p] + ((dest_row0 + b) * (N / peers) + x % (N / peers)) complex
 * elements - i.e. into the [rows][N / peers] column slab of device p.  The pointers are device
 * pointers valid in this process: local memory for the own slab, peer memory mapped with
 * mufft_cuda_ipc_import for the others, so the kernel's stores are the all-to-all transpose and
 * travel over NVLink; no pack pass, no separate collective.  `peers` must be a power of two that
 * divides N.  Enqueued on the current stream; the caller synchronises the devices (any barrier
 * ordered after this call on every device) before reading the slabs.
 * With an R2C plan (N reals per row, no FULL_R2C / zero padding, N/2 <= the single-kernel limit) the
 * N/2 + 1 bins of row b are scattered the same way: bin k to device min(k / (N/2 / peers), peers - 1);
 * the slabs are [rows][N/2 / peers] complex, except the last device's, which is one column wider and
 * takes the Nyquist bin N/2 as its last column (the layout of mufft_b200/slab.py: SlabRFFT2D). */
int mufft_execute_plan_1d_batch_scatter_async(mufft_plan_1d *plan, void *const *peer_outputs, unsigned peers,
                                              size_t dest_row0, const void *input, size_t batch, size_t input_stride_bytes);

/* Plain device memory (cudaMalloc) that can be shared between the processes of one node, and
 * CUDA IPC handles (64 bytes) for it.  mufft_cuda_ipc_import maps another process's allocation
 * into this one (peer access over NVLink is enabled on demand). */
void *mufft_cuda_device_alloc(size_t size);
void mufft_cuda_device_free(void *ptr);
/* Page-locked (pinned) host memory: the synchronous batch calls stage ordinary host memory through the same pipeline, but
 * the copies of pinned buffers run at the full rate of the bus and overlap in both directions. */
void *mufft_cuda_host_alloc(size_t size);
void mufft_cuda_host_free(void *ptr);
/* Synchronous helpers for callers that keep their data in device memory: copy `bytes` between any two kinds of
 * memory (host, pinned, managed, device), fill device-accessible memory with a byte value.  Both are ordered after the
 * work already queued on the current stream and complete on return.  0 on success. */
int mufft_cuda_copy(void *dst, const void *src, size_t bytes);
int mufft_cuda_memset(void *ptr, int value, size_t bytes);
int mufft_cuda_ipc_export(void *handle64, void *device_ptr);
void *mufft_cuda_ipc_import(const void *handle64);
int mufft_cuda_ipc_close(void *ptr);

/* `batch` independent images, tightly packed one after another (each laid out as fft.h
 * describes for the plan's kind: [Ny][Nx] complex, R2C output / C2R input with row pitch Nx
 * complex).  C2R output images are Nx*Ny floats each, densely packed. */
int mufft_execute_plan_2d_batch(mufft_plan_2d *plan, void *output, const void *input, size_t batch);
int mufft_execute_plan_2d_batch_async(mufft_plan_2d *plan, void *output, const void *input, size_t batch);

/* Batched convolution stages.  Blocks are mufft_conv_get_transformed_block_size() bytes apart;
 * inputs and outputs are tightly packed.  If second_stride_bytes is 0 every block of the batch
 * is multiplied by the same `input_second` spectrum (one filter, many signal blocks). */
int mufft_execute_conv_input_batch(mufft_plan_conv *plan, unsigned block, void *output, const void *input, size_t batch);
int mufft_execute_conv_output_batch(mufft_plan_conv *plan, void *output, const void *input_first,
                                    const void *input_second, size_t batch, size_t second_stride_bytes);
int mufft_execute_conv_input_batch_async(mufft_plan_conv *plan, unsigned block, void *output, const void *input, size_t batch);
int mufft_execute_conv_output_batch_async(mufft_plan_conv *plan, void *output, const void *input_first,
                                          const void *input_second, size_t batch, size_t second_stride_bytes);

/* Whole convolution of `batch` blocks in one call: output[b] = IFFT(FFT(input_first[b]) .* filter / N),
 * i.e. mufft_execute_conv_input(plan, MUFFT_CONV_BLOCK_FIRST, ...) followed by
 * mufft_execute_conv_output(...) for every block, with `filter_spectrum` exactly what
 * mufft_execute_conv_input(plan, MUFFT_CONV_BLOCK_SECOND, filter_spectrum, taps) produced
 * (mufft_conv_get_transformed_block_size() bytes).  When the transform fits the on-chip engine
 * (stereo N <= 8192, mono N <= 16384) this is ONE kernel and the block spectrum never reaches HBM:
 * a zero-padded stereo block moves 12N bytes instead of the 28N of the two-call sequence.
 * Inputs are tightly packed blocks laid out as conv_input(FIRST) expects them, outputs as
 * conv_output writes them (stereo N complex, mono N floats per block).  filter_stride_bytes 0 =
 * one filter for every block, else one spectrum per block that many bytes apart. */
int mufft_execute_conv_batch(mufft_plan_conv *plan, void *output, const void *input_first,
                             const void *filter_spectrum, size_t batch, size_t filter_stride_bytes);
int mufft_execute_conv_batch_async(mufft_plan_conv *plan, void *output, const void *input_first,
                                   const void *filter_spectrum, size_t batch, size_t filter_stride_bytes);

/* Uniformly partitioned convolution for filters longer than one block (frequency-domain delay
 * line).  The reference leaves the accumulation over filter partitions to the caller
 * (ref fft.h:170-174: one mufft_execute_conv_output per partition, outputs added in the time
 * domain); here the products are summed in front of ONE inverse transform per block:
 *     output[b] = IFFT( sum_{p < parts} S[b - p] .* F[p] ) / N,      b = 0 .. batch-1
 * S[j]: block spectra in time order (what mufft_execute_conv_input(plan, MUFFT_CONV_BLOCK_FIRST, ..)
 * wrote), mufft_conv_get_transformed_block_size() bytes apart; `spectra` points at S[0], and the
 * parts-1 blocks in front of it (S[-1] .. S[-(parts-1)]: the tail of the previous call, zeros at the
 * start of a stream) must be valid memory.  F[p]: spectra of the filter partitions
 * (mufft_execute_conv_input(plan, MUFFT_CONV_BLOCK_SECOND, ..) of taps [p*L, (p+1)*L)), the same
 * distance apart.  Outputs as mufft_execute_conv_output writes them, tightly packed; the caller
 * overlap-adds consecutive output blocks exactly as with a one-block filter.  Device-accessible
 * (device or managed) memory only; the transform must fit the on-chip engine (stereo N <= 16384,
 * mono N <= 32768), otherwise nonzero is returned. */
int mufft_execute_conv_output_partitioned_batch(mufft_plan_conv *plan, void *output, const void *spectra,
                                                const void *filter_spectra, unsigned parts, size_t batch);
int mufft_execute_conv_output_partitioned_batch_async(mufft_plan_conv *plan, void *output, const void *spectra,
                                                      const void *filter_spectra, unsigned parts, size_t batch);

/* ---- one large 2D transform over several devices of this process (slab decomposition) ----------------------
 * BASELINE-style: device g owns rows [g Ny/G, (g+1) Ny/G) of the [Ny][Nx] array.  Row pass local, with the
 * exchange fused into it: the row kernel stores element x of every row straight into the column slab of device
 * x / (Nx/G) over NVLink (peer access; no pack pass, no collective library).  Events order the devices, then each
 * transforms the columns of its slab in place.  Result: device g holds columns [g Nx/G, (g+1) Nx/G) of the
 * spectrum as a row-major [Ny][Nx/G] array ("transposed-out"); the _host call gathers the natural [Ny][Nx]
 * order.  R2C (forward, Nx reals per row): real row transforms first (half the exchanged volume); the Nx/2 lower
 * columns are split evenly and the Nyquist column Nx/2 rides with the last device.
 * ndev: a power of two (<= 16) dividing Ny and Nx (R2C: Nx/2); devices: CUDA ordinals, NULL = 0 .. ndev-1.  The
 * reference's function this distributes is mufft_execute_plan_2d (fft.c:852-989). */
typedef struct mufft_plan_2d_slab mufft_plan_2d_slab;
mufft_plan_2d_slab *mufft_cuda_create_plan_2d_slab_c2c(unsigned Nx, unsigned Ny, int direction, unsigned ndev, const int *devices, unsigned flags);
mufft_plan_2d_slab *mufft_cuda_create_plan_2d_slab_r2c(unsigned Nx, unsigned Ny, unsigned ndev, const int *devices, unsigned flags);
void mufft_cuda_free_plan_2d_slab(mufft_plan_2d_slab *plan);
unsigned mufft_cuda_slab_devices(const mufft_plan_2d_slab *plan);
/* Plan-owned device buffers of device index g: the input row slab ([Ny/G][Nx] complex, R2C: floats) and the
 * output column slab ([Ny][columns] complex; columns and the first column's index in the spectrum below). */
void *mufft_cuda_slab_input(mufft_plan_2d_slab *plan, unsigned g);
void *mufft_cuda_slab_output(mufft_plan_2d_slab *plan, unsigned g);
size_t mufft_cuda_slab_input_bytes(const mufft_plan_2d_slab *plan, unsigned g);
size_t mufft_cuda_slab_output_bytes(const mufft_plan_2d_slab *plan, unsigned g);
unsigned mufft_cuda_slab_output_columns(const mufft_plan_2d_slab *plan, unsigned g, unsigned *first_column);
/* Transforms the input slabs into the output slabs on all devices; returns when every device is done. */
int mufft_cuda_execute_plan_2d_slab(mufft_plan_2d_slab *plan);
/* Device time of the last execute in milliseconds: max over the devices, row pass start to column pass end. */
float mufft_cuda_slab_last_ms(const mufft_plan_2d_slab *plan);
/* ... and of its row pass + exchange alone (the column pass is the rest, less the wait for the slowest device). */
float mufft_cuda_slab_last_rows_ms(const mufft_plan_2d_slab *plan);
/* Whole arrays in host memory: input [Ny][Nx] (complex, R2C: floats), output [Ny][Nx] complex in natural order
 * (R2C: [Ny][Nx/2 + 1]), i.e. what mufft_execute_plan_2d computes, on ndev devices. */
int mufft_cuda_execute_plan_2d_slab_host(mufft_plan_2d_slab *plan, void *output, const void *input);

/* Plan-creation flags of this implementation, OR-ed into the `flags` argument of the fft.h constructors
 * (the reference's own flags use bits 0-2, 16 and 17, its internal ones bits 24-28; ref fft.h:52-65,
 * fft_internal.h:180-192).  They select the slower fallback paths, which are kept for measurement and
 * for tests; results are the same (within rounding) on every path.  There is no process-global switch:
 * what a plan does is fixed when it is created, and two plans with different flags can run side by side.
 *   MUFFT_CUDA_FLAG_SEPARATE_2D     2D plans: row kernel + column kernel through plan scratch instead of the
 *                                   fused persistent kernel (which exists for 256..2048 complex points per side).
 *   MUFFT_CUDA_FLAG_TWO_KERNEL      rows longer than 8192 points: the two-kernel four-step split through
 *                                   scratch in HBM instead of the single-launch kernels (16384 points on one
 *                                   CTA, 32768..1048576 points on the four-step kernel that keeps its
 *                                   intermediate in L2).
 *   MUFFT_CUDA_FLAG_GENERIC_PASSES  lengths beyond 8192 points: one generic radix-8/4/2 Stockham pass per
 *                                   kernel launch (the reference's own structure, fft.c:811-850).
 *   MUFFT_CUDA_FLAG_FOUR_STEP_16K   16384-point rows: the single-launch four-step kernel (16 x 1024) instead
 *                                   of the one-CTA engine kernel. */
#define MUFFT_CUDA_FLAG_SEPARATE_2D (1u << 20)
#define MUFFT_CUDA_FLAG_TWO_KERNEL (1u << 21)
#define MUFFT_CUDA_FLAG_GENERIC_PASSES (1u << 22)
#define MUFFT_CUDA_FLAG_FOUR_STEP_16K (1u << 23)

/* Introspection used by tests and the bench: which path a plan takes.
 * 1D: 1 = one kernel launch per batch (on-chip engine up to 16384 points, L2-resident four-step beyond),
 *     0 = several launches through scratch.
 * 2D: bit 0 = rows in one launch, bit 1 = columns on the shared-memory column kernel, bit 2 = the fused
 *     persistent kernel does the whole transform. */
int mufft_cuda_plan_1d_uses_engine(const mufft_plan_1d *plan);
int mufft_cuda_plan_2d_paths(const mufft_plan_2d *plan);

#ifdef __cplusplus
}
#endif
#endif
