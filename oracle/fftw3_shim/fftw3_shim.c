/* Test infrastructure only (see fftw3.h in this directory): float64 stand-in for the FFTW3f calls
 * of the reference's test.c / bench.c.  Power-of-two sizes only (all the reference ever asks for). */
#include "fftw3.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { K_C2C, K_R2C, K_C2R };
struct fftwf_plan_s {
    int kind, n0, n1, sign; /* n0 rows (1 for 1D), n1 columns */
    void *in, *out;
    double *wr, *wi;        /* exp(2*pi*i*k/nmax), k < nmax/2 */
    int nmax;
    double *buf;            /* n0 * n1 complex doubles */
};

void *fftwf_malloc(size_t n)
{
    void *p = NULL;
    return posix_memalign(&p, 64, n ? n : 64) ? NULL : p;
}
void fftwf_free(void *p) { free(p); }
void fftwf_cleanup(void) {}

/* In-place decimation-in-time FFT of n complex doubles at stride `st` (in complex elements). */
static void fft_strided(const struct fftwf_plan_s *pl, double *x, int n, int st, int sign)
{
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            double tr = x[2 * i * st], ti = x[2 * i * st + 1];
            x[2 * i * st] = x[2 * j * st]; x[2 * i * st + 1] = x[2 * j * st + 1];
            x[2 * j * st] = tr; x[2 * j * st + 1] = ti;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1, step = pl->nmax / len;
        for (int base = 0; base < n; base += len) {
            for (int k = 0; k < half; k++) {
                const double wr = pl->wr[k * step], wi = sign * pl->wi[k * step];
                double *a = x + 2 * (size_t)(base + k) * st, *b = x + 2 * (size_t)(base + k + half) * st;
                const double tr = b[0] * wr - b[1] * wi, ti = b[0] * wi + b[1] * wr;
                b[0] = a[0] - tr; b[1] = a[1] - ti;
                a[0] += tr; a[1] += ti;
            }
        }
    }
}

static fftwf_plan make_plan(int kind, int n0, int n1, int sign, void *in, void *out)
{
    if (n0 < 1 || n1 < 1 || (n0 & (n0 - 1)) || (n1 & (n1 - 1))) return NULL;
    struct fftwf_plan_s *pl = calloc(1, sizeof(*pl));
    if (!pl) return NULL;
    pl->kind = kind; pl->n0 = n0; pl->n1 = n1; pl->sign = sign; pl->in = in; pl->out = out;
    pl->nmax = n0 > n1 ? n0 : n1;
    if (pl->nmax < 2) pl->nmax = 2;
    pl->wr = malloc(sizeof(double) * pl->nmax / 2);
    pl->wi = malloc(sizeof(double) * pl->nmax / 2);
    pl->buf = malloc(sizeof(double) * 2 * (size_t)n0 * n1);
    if (!pl->wr || !pl->wi || !pl->buf) { fftwf_destroy_plan(pl); return NULL; }
    for (int k = 0; k < pl->nmax / 2; k++) {
        pl->wr[k] = cos(2.0 * M_PI * k / pl->nmax);
        pl->wi[k] = sin(2.0 * M_PI * k / pl->nmax);
    }
    return pl;
}

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags)
{ (void)flags; return make_plan(K_C2C, 1, n, sign, in, out); }
fftwf_plan fftwf_plan_dft_2d(int n0, int n1, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags)
{ (void)flags; return make_plan(K_C2C, n0, n1, sign, in, out); }
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned flags)
{ (void)flags; return make_plan(K_R2C, 1, n, -1, in, out); }
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned flags)
{ (void)flags; return make_plan(K_C2R, 1, n, +1, in, out); }
fftwf_plan fftwf_plan_dft_r2c_2d(int n0, int n1, float *in, fftwf_complex *out, unsigned flags)
{ (void)flags; return make_plan(K_R2C, n0, n1, -1, in, out); }
fftwf_plan fftwf_plan_dft_c2r_2d(int n0, int n1, fftwf_complex *in, float *out, unsigned flags)
{ (void)flags; return make_plan(K_C2R, n0, n1, +1, in, out); }

void fftwf_destroy_plan(fftwf_plan pl)
{
    if (!pl) return;
    free(pl->wr); free(pl->wi); free(pl->buf); free(pl);
}

static void transform_2d(const struct fftwf_plan_s *pl, int sign)
{
    const int n0 = pl->n0, n1 = pl->n1;
    if (n1 > 1) for (int y = 0; y < n0; y++) fft_strided(pl, pl->buf + 2 * (size_t)y * n1, n1, 1, sign);
    if (n0 > 1) for (int x = 0; x < n1; x++) fft_strided(pl, pl->buf + 2 * (size_t)x, n0, n1, sign);
}

void fftwf_execute(const fftwf_plan pl)
{
    const int n0 = pl->n0, n1 = pl->n1, h = n1 / 2 + 1;
    double *b = pl->buf;
    if (pl->kind == K_C2C) {
        const float *in = pl->in; float *out = pl->out;
        for (size_t i = 0; i < 2 * (size_t)n0 * n1; i++) b[i] = in[i];
        transform_2d(pl, pl->sign);
        for (size_t i = 0; i < 2 * (size_t)n0 * n1; i++) out[i] = (float)b[i];
    } else if (pl->kind == K_R2C) {
        const float *in = pl->in; float *out = pl->out;
        for (size_t i = 0; i < (size_t)n0 * n1; i++) { b[2 * i] = in[i]; b[2 * i + 1] = 0.0; }
        transform_2d(pl, -1);
        for (int y = 0; y < n0; y++)
            for (int x = 0; x < h; x++) {
                out[2 * ((size_t)y * h + x)] = (float)b[2 * ((size_t)y * n1 + x)];
                out[2 * ((size_t)y * h + x) + 1] = (float)b[2 * ((size_t)y * n1 + x) + 1];
            }
    } else {
        /* Half spectrum [n0][n1/2+1] -> real [n0][n1]: inverse along the columns of the half
         * spectrum first, then every row is completed by Hermitian symmetry and inverted; the
         * real part is the answer (imaginary parts of the DC / Nyquist bins drop out, as in FFTW). */
        const float *in = pl->in; float *out = pl->out;
        for (int y = 0; y < n0; y++) {
            for (int x = 0; x < n1; x++) { b[2 * ((size_t)y * n1 + x)] = 0.0; b[2 * ((size_t)y * n1 + x) + 1] = 0.0; }
            for (int x = 0; x < h; x++) {
                b[2 * ((size_t)y * n1 + x)] = in[2 * ((size_t)y * h + x)];
                b[2 * ((size_t)y * n1 + x) + 1] = in[2 * ((size_t)y * h + x) + 1];
            }
        }
        if (n0 > 1) for (int x = 0; x < h; x++) fft_strided(pl, b + 2 * (size_t)x, n0, n1, +1);
        for (int y = 0; y < n0; y++) {
            double *row = b + 2 * (size_t)y * n1;
            for (int x = 1; x < n1 - x; x++) { row[2 * (n1 - x)] = row[2 * x]; row[2 * (n1 - x) + 1] = -row[2 * x + 1]; }
            row[1] = 0.0;
            if (n1 > 1) row[2 * (n1 / 2) + 1] = 0.0;
            if (n1 > 1) fft_strided(pl, row, n1, 1, +1);
            for (int x = 0; x < n1; x++) out[(size_t)y * n1 + x] = (float)row[2 * x];
        }
    }
}
