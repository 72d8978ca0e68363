/* Test infrastructure only: the slice of the FFTW3 single-precision API that the reference's own
 * test.c and bench.c use (test.c:29-372, bench.c:26-199), backed by a float64 radix-2 FFT
 * (fftw3_shim.c).  FFTW3f is an un-vendored, unpinned dependency of the reference's executables
 * (CMakeLists.txt:68) and is absent from this image; with this shim muFFT-test and muFFT-bench
 * compile UNCHANGED from /root/reference and link against libmufft_b200.so (oracle/Makefile:
 * `make consumers`).  Semantics follow the published FFTW3 manual: unnormalised transforms,
 * row-major n0 x n1, r2c/c2r half spectra of n1/2+1 bins per row. */
#ifndef MUFFT_B200_FFTW3_SHIM_H
#define MUFFT_B200_FFTW3_SHIM_H

#include <stddef.h>
#include <stdio.h> /* the real fftw3.h includes it, and test.c relies on that (test.c:503) */

#ifdef __cplusplus
extern "C" {
#endif

typedef float fftwf_complex[2];
typedef struct fftwf_plan_s *fftwf_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
fftwf_plan fftwf_plan_dft_2d(int n0, int n1, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned flags);
fftwf_plan fftwf_plan_dft_r2c_2d(int n0, int n1, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_2d(int n0, int n1, fftwf_complex *in, float *out, unsigned flags);
void fftwf_execute(const fftwf_plan plan);
void fftwf_destroy_plan(fftwf_plan plan);
void fftwf_cleanup(void);

#ifdef __cplusplus
}
#endif
#endif
