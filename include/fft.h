/* muFFT public C API, as implemented by the B200 (sm_100a) library in this repository.
 *
 * This header is the drop-in boundary.  Every declaration, constant and typedef below has
 * the same name, value, signature and calling convention as the corresponding item in the
 * reference's fft.h (cited per item as "ref fft.h:LINE"), so code written against the
 * reference header compiles and links against libmufft_b200.so unchanged.  The comments are
 * ours and describe what the GPU implementation does; the API surface is muFFT's
 * (MIT licence, (C) 2015 Hans-Kristian Arntzen).
 *
 * Semantics kept from the reference:
 *   - interleaved single-precision complex {float re, im}; natural-order input and output;
 *   - transforms are unnormalised (a forward+inverse pair scales by N, or Nx*Ny in 2D);
 *     only mufft_execute_conv_output() applies 1/N;
 *   - power-of-two sizes only; constructors return NULL on any invalid argument or failure;
 *   - execute calls are synchronous: on return `output` is complete and host-readable.
 *
 * GPU-specific behaviour (additive, never required):
 *   - any pointer may be pageable host, pinned host, managed or device memory; host memory is
 *     staged through plan-owned device buffers.  mufft_alloc() returns CUDA managed memory,
 *     which both the host and the kernels can touch;
 *   - MUFFT_FLAG_CPU_* bits are accepted and ignored (there is no CPU path);
 *   - batched / asynchronous / multi-GPU entry points live in mufft_cuda.h.
 */
#ifndef MUFFT_H__
#define MUFFT_H__

#ifdef __cplusplus
extern "C" {
#endif

#include <stddef.h>

/* ref fft.h:30-36 */
#if !defined(__cplusplus) && defined(__STDC_VERSION__) && __STDC_VERSION__ >= 199901L
#define MUFFT_RESTRICT restrict
#else
#define MUFFT_RESTRICT
#endif

/* Direction of a complex transform: sign of the exponent (ref fft.h:44,46). */
#define MUFFT_FORWARD (-1)
#define MUFFT_INVERSE  (1)

/* Planning flags (ref fft.h:52-65).  The CPU_* bits select x86 instruction sets in the
 * reference; here they are accepted and have no effect. */
#define MUFFT_FLAG_CPU_ANY (0)
#define MUFFT_FLAG_CPU_NO_SIMD ((1 << 16) - 1)
#define MUFFT_FLAG_CPU_NO_AVX (1 << 0)
#define MUFFT_FLAG_CPU_NO_SSE3 (1 << 1)
#define MUFFT_FLAG_CPU_NO_SSE (1 << 2)
/* R2C writes all N bins (the upper half being the conjugate mirror) instead of N/2+1. */
#define MUFFT_FLAG_FULL_R2C (1 << 16)
/* 1D only: input[N/2..N) is taken to be zero and is never read; the buffer may be N/2 long. */
#define MUFFT_FLAG_ZERO_PAD_UPPER_HALF (1 << 17)

/* ---- 1D transforms (ref fft.h:78-121) ------------------------------------------------ */

typedef struct mufft_plan_1d mufft_plan_1d;

/* N-point complex transform, N a power of two >= 2 (ref fft.h:86). */
mufft_plan_1d *mufft_create_plan_1d_c2c(unsigned N, int direction, unsigned flags);

/* N real samples -> N/2+1 complex bins (N bins with MUFFT_FLAG_FULL_R2C); N a power of two
 * >= 4 (ref fft.h:99). */
mufft_plan_1d *mufft_create_plan_1d_r2c(unsigned N, unsigned flags);

/* N/2+1 complex bins -> N real samples, unnormalised; N a power of two >= 4 (ref fft.h:111). */
mufft_plan_1d *mufft_create_plan_1d_c2r(unsigned N, unsigned flags);

/* Run one transform.  `input` is not modified; `output` must not alias it (ref fft.h:117). */
void mufft_execute_plan_1d(mufft_plan_1d *plan, void * MUFFT_RESTRICT output, const void * MUFFT_RESTRICT input);

/* NULL is accepted (ref fft.h:121). */
void mufft_free_plan_1d(mufft_plan_1d *plan);

/* ---- 1D FFT convolution (ref fft.h:131-207) ------------------------------------------- */

/* One real channel (*) one real channel. */
#define MUFFT_CONV_METHOD_FLAG_MONO_MONO 0
/* Interleaved stereo (complex) first block (*) real second block. */
#define MUFFT_CONV_METHOD_FLAG_STEREO_MONO 1
/* The first / second block is zero padded as MUFFT_FLAG_ZERO_PAD_UPPER_HALF describes. */
#define MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_FIRST (1 << 2)
#define MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_SECOND (1 << 3)

#define MUFFT_CONV_BLOCK_FIRST 0
#define MUFFT_CONV_BLOCK_SECOND 1

typedef struct mufft_plan_conv mufft_plan_conv;

/* Circular convolution of length N (power of two >= 4) through the frequency domain
 * (ref fft.h:168). */
mufft_plan_conv *mufft_create_plan_conv(unsigned N, unsigned flags, unsigned method);

/* Forward-transform block 0 or 1 into a buffer of mufft_conv_get_transformed_block_size()
 * bytes (ref fft.h:179). */
void mufft_execute_conv_input(mufft_plan_conv *plan, unsigned block, void *output, const void *input);

/* Bytes one transformed block occupies (ref fft.h:185). */
size_t mufft_conv_get_transformed_block_size(mufft_plan_conv *plan);

/* output = IFFT(first .* second) / N: N reals (mono) or N complex (stereo) (ref fft.h:195). */
void mufft_execute_conv_output(mufft_plan_conv *plan, void *output, const void *input_first, const void *input_second);

void mufft_free_plan_conv(mufft_plan_conv *plan);

/* output[i] = normalization * a[i] * b[i] for `samples` complex values (ref fft.h:201-202). */
typedef void (*mufft_convolve_func)(void *output, const void *a, const void *b,
                                    float normalization, unsigned samples);

/* Host-callable spectral multiply; runs the CUDA kernel and synchronises (ref fft.h:207). */
mufft_convolve_func mufft_get_convolve_func(unsigned flags);

/* ---- 2D transforms, row-major [Ny][Nx] (ref fft.h:221-276) ---------------------------- */

typedef struct mufft_plan_2d mufft_plan_2d;

/* Nx columns by Ny rows, both powers of two >= 2 (ref fft.h:232). */
mufft_plan_2d *mufft_create_plan_2d_c2c(unsigned Nx, unsigned Ny, int direction, unsigned flags);

/* Real [Ny][Nx] -> complex rows of pitch Nx; columns 0..Nx/2 are defined (all Nx with
 * MUFFT_FLAG_FULL_R2C).  Nx >= 4 (ref fft.h:249). */
mufft_plan_2d *mufft_create_plan_2d_r2c(unsigned Nx, unsigned Ny, unsigned flags);

/* Complex rows of pitch Nx (columns 0..Nx/2 used) -> real [Ny][Nx].  The output buffer must
 * hold 2*Nx*Ny floats; the result occupies the first Nx*Ny (ref fft.h:266). */
mufft_plan_2d *mufft_create_plan_2d_c2r(unsigned Nx, unsigned Ny, unsigned flags);

void mufft_execute_plan_2d(mufft_plan_2d *plan, void * MUFFT_RESTRICT output, const void * MUFFT_RESTRICT input);

void mufft_free_plan_2d(mufft_plan_2d *plan);

/* ---- memory (ref fft.h:286-297) ------------------------------------------------------- */

/* 64-byte (in fact 256-byte) aligned CUDA managed memory; NULL on failure. */
void *mufft_alloc(size_t size);
/* Same, zero filled. */
void *mufft_calloc(size_t size);
/* NULL is accepted. */
void mufft_free(void *ptr);

#ifdef __cplusplus
}
#endif

#endif
