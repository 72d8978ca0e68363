/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
 *
 * C-ABI of the CPU oracle: our own plain-C restatement of the reference muFFT algorithm
 * (scalar path), used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * to check the CUDA library.  Same call shapes as include/fft.h with an `oracle_` prefix so
 * both libraries can live in one process.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement (a) value-for-value
 * (numerically equal, every element) against the unmodified reference compiled into
 * oracle/_ref/libmufft_ref.so with MUFFT_FLAG_CPU_NO_SIMD, whenever that library is present,
 * and (b) against committed outputs of that same reference build in tests/golden/ (made by
 * tests/golden/make_golden.py), which is what runs where /root/reference does not exist.
 */
#ifndef MUFFT_ORACLE_H
#define MUFFT_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_plan_1d oracle_plan_1d;
typedef struct oracle_plan_2d oracle_plan_2d;
typedef struct oracle_plan_conv oracle_plan_conv;

oracle_plan_1d *oracle_create_plan_1d_c2c(unsigned N, int direction, unsigned flags);
oracle_plan_1d *oracle_create_plan_1d_r2c(unsigned N, unsigned flags);
oracle_plan_1d *oracle_create_plan_1d_c2r(unsigned N, unsigned flags);
void oracle_execute_plan_1d(oracle_plan_1d *plan, void *output, const void *input);
void oracle_free_plan_1d(oracle_plan_1d *plan);
/* Writes "radix@p" pairs of the plan into radix[]/p[] (up to cap) and returns the step count. */
unsigned oracle_plan_1d_steps(const oracle_plan_1d *plan, unsigned *radix, unsigned *p, unsigned cap);

oracle_plan_conv *oracle_create_plan_conv(unsigned N, unsigned flags, unsigned method);
void oracle_execute_conv_input(oracle_plan_conv *plan, unsigned block, void *output, const void *input);
size_t oracle_conv_get_transformed_block_size(oracle_plan_conv *plan);
void oracle_execute_conv_output(oracle_plan_conv *plan, void *output, const void *first, const void *second);
void oracle_free_plan_conv(oracle_plan_conv *plan);
void oracle_convolve(void *output, const void *a, const void *b, float normalization, unsigned samples);

oracle_plan_2d *oracle_create_plan_2d_c2c(unsigned Nx, unsigned Ny, int direction, unsigned flags);
oracle_plan_2d *oracle_create_plan_2d_r2c(unsigned Nx, unsigned Ny, unsigned flags);
oracle_plan_2d *oracle_create_plan_2d_c2r(unsigned Nx, unsigned Ny, unsigned flags);
void oracle_execute_plan_2d(oracle_plan_2d *plan, void *output, const void *input);
void oracle_free_plan_2d(oracle_plan_2d *plan);

/* Batched helper for bench.py's cpu_baseline leg: runs `count` independent 1D transforms
 * (element strides given in bytes) on the calling thread. */
void oracle_execute_plan_1d_many(oracle_plan_1d *plan, void *output, const void *input,
                                 size_t count, size_t out_stride_bytes, size_t in_stride_bytes);

#ifdef __cplusplus
}
#endif
#endif
