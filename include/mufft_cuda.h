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
 * mono N <= 32768 with the default mufft_cuda_set_onchip_limit), otherwise nonzero is returned. */
int mufft_execute_conv_output_partitioned_batch(mufft_plan_conv *plan, void *output, const void *spectra,
                                                const void *filter_spectra, unsigned parts, size_t batch);
int mufft_execute_conv_output_partitioned_batch_async(mufft_plan_conv *plan, void *output, const void *spectra,
                                                      const void *filter_spectra, unsigned parts, size_t batch);

/* Row-transform kernel flavour: 0 = direct loads (default), 1 = persistent CTAs with bulk-copy
 * prefetch into shared memory when the input rows are 16-byte aligned.  Same results either way. */
void mufft_cuda_set_prefetch(int enable);
/* 2D C2C transforms of 256..2048 points per side.
 *   1 (default)  one persistent kernel does rows and columns of the whole batch with the intermediate
 *                kept in L2, in the flavour measured fastest for the shape;
 *   0            separate row and column kernels per group of images (what every other 2D shape uses);
 *   2..16        that kernel in one fixed flavour, for measurements and tests: 8 plain; 7 one-sided
 *                (release / acquire) fences in the ticket protocol; 3 = 7 + twiddle tables staged in
 *                shared memory; 6 = 7 + only the power-of-two twiddles loaded, the rest derived;
 *                5 = 3 + 6; 2 / 4 = tiles brought into shared memory by the TMA unit one tile ahead
 *                (bulk copy for row tiles, 2D tensor-map boxes for column tiles), plain / as 3 -
 *                measured slower than direct loads on B200, kept as an experiment; 9 / 10 = 3 / 6 with
 *                tickets drawn two tiles ahead and the dependency of a column ticket read early
 *                (not measured yet); 11 = 6 compiled for three CTAs per SM instead of two (80 registers,
 *                the excess spilled; not measured yet); 12 / 13 = 3 / 6 with row and column tickets
 *                alternating inside a round instead of all row tickets first (not measured yet);
 *                14 = 7 at three CTAs per SM, 15 = 11 + 13, 16 = 3 at three CTAs per SM (three exchange areas
 *                and three table copies fit the shared memory of an SM for 1024 x 1024 and the 256 / 512-point
 *                shapes only) (not measured yet).
 * Same results (within rounding of the derived twiddles) in all modes.  MUFFT_CUDA_FUSED_2D in the
 * environment sets the initial value. */
void mufft_cuda_set_fused_2d(int mode);
/* Fused 2D kernel, measurement knob: the column pass of an image starts this many images behind its row
 * pass (0 = default: chosen from the number of resident CTAs, 3 for 1024 x 1024 on a B200).  Same results
 * for every value.  MUFFT_CUDA_FUSED_2D_LAG sets the initial value. */
void mufft_cuda_set_fused_2d_lag(unsigned images);
/* Longest row transform done by ONE kernel, for plans created afterwards: 16384 (default: one
 * 512-thread CTA per 16384-point transform, read once and written once) or 8192 (16384 points take
 * the two-kernel split; measured 1.8x slower, kept for comparison).  Affects 1D plans of 16384
 * complex / 32768 real points and 2D plans with such rows.  Other values are rounded down to one of
 * the two.  Same results either way.  MUFFT_CUDA_ONCHIP_LIMIT in the environment sets the initial value. */
void mufft_cuda_set_onchip_limit(unsigned n);
unsigned mufft_cuda_get_onchip_limit(void);
/* 2D R2C plans whose row transforms have 256..2048 complex points (Nx = 512..4096 reals) and 256..2048
 * rows: 1 = the fused persistent kernel (real row transforms with the resolve epilogue straight into
 * the output, column transforms in place, intermediate kept in L2), 0 (default until measured on the
 * GPU) = row kernel + column kernel through plan scratch.  Same results either way.
 * MUFFT_CUDA_FUSED_2D_REAL sets the initial value. */
void mufft_cuda_set_fused_2d_real(int enable);
/* Latency of synchronous calls on small transforms (one launch + one synchronise per
 * mufft_execute_plan_1d call): 1 = the library busy-polls the stream instead of blocking in the driver
 * when a call touches at most 1 MiB; 0 (default until measured).  MUFFT_CUDA_SPIN_SYNC sets the initial value. */
void mufft_cuda_set_spin_sync(int enable);
/* Synchronous calls on CUDA managed memory (mufft_alloc) that the caller fills and reads on the host between
 * calls: 0 (default until measured) = the kernel's page faults migrate the pages, 1 = buffers of 256 KiB or more
 * are prefetched to the device in front of the kernel, 2 = and the output back to the host behind it (only
 * right when the host reads it next).  MUFFT_CUDA_MANAGED_PREFETCH sets the initial value. */
void mufft_cuda_set_managed_prefetch(int level);
/* Row transforms of 32768 and 65536 points, for plans created afterwards: 1 = ONE kernel on a
 * thread-block cluster of 2 / 4 CTAs (512 threads each, the transform in registers, exchange through
 * distributed shared memory: 1 = remote stores + release / acquire cluster barriers, 2 = st.async
 * remote stores counted on an mbarrier of the receiving CTA), 0 (default until measured on the GPU) =
 * the two-kernel split.  Plain and
 * zero-padded C2C and the scatter epilogue of the slab exchange; real transforms of these lengths keep
 * the two-kernel path.  Same results either way.  MUFFT_CUDA_CLUSTER_ROWS sets the initial value. */
void mufft_cuda_set_cluster_rows(int mode);
/* Transform lengths beyond the single-kernel limit, up to 2097152: 1 (default) = two kernels (N = N1*N2: column transforms, then
 * twiddle + row transforms with a transposed store), 0 = one generic radix pass per kernel. */
void mufft_cuda_set_two_kernel(int enable);

/* Introspection used by tests and the bench: which path a plan takes.
 * 1D: 1 = single-kernel shared-memory engine, 0 = multi-pass generic kernels.
 * 2D: bit 0 = rows on the engine, bit 1 = columns on the shared-memory column kernel. */
int mufft_cuda_plan_1d_uses_engine(const mufft_plan_1d *plan);
int mufft_cuda_plan_2d_paths(const mufft_plan_2d *plan);

#ifdef __cplusplus
}
#endif
#endif
