/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT (see mufft_oracle.h for the parity status).
 *
 * Plain-C restatement of the reference muFFT algorithm, scalar flavour: out-of-place Stockham
 * autosort with radix-8/4/2 decimation-in-time passes, the R2C/C2R resolve butterflies, the 2D
 * row/column driver and the FFT convolution.  Written from the reference's behaviour, not its
 * text: one strided pass routine serves both the horizontal and the vertical kernels, plans are
 * a fixed-size step array, and twiddle levels are addressed by level rather than by running
 * offset.  The floating-point operation ORDER of every butterfly follows the reference's
 * scalar kernels so that outputs are value-identical to `MUFFT_FLAG_CPU_NO_SIMD` plans; each
 * routine cites the reference lines it restates (paths relative to /root/reference).
 *
 * Build with -ffp-contract=off: the reference is compiled without FMA (CMakeLists.txt:6-7).
 */
#include "mufft_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* Flag values of fft.h:44-65,131-143 and the internal bits of fft_internal.h:180-192. */
#define FWD (-1)
#define INV (1)
#define F_FULL_R2C (1u << 16)
#define F_ZERO_PAD (1u << 17)
#define M_STEREO 1u
#define M_ZP_FIRST (1u << 2)
#define M_ZP_SECOND (1u << 3)
#define PAD_COMPLEX 4 /* fft_internal.h:203 */

typedef struct { float re, im; } cf;

/* Complex helpers: same operand order as fft_internal.h:84-106. */
static inline cf c_add(cf a, cf b) { cf r = { a.re + b.re, a.im + b.im }; return r; }
static inline cf c_sub(cf a, cf b) { cf r = { a.re - b.re, a.im - b.im }; return r; }
static inline cf c_mul(cf a, cf b)
{
    cf r = { a.re * b.re - a.im * b.im, a.re * b.im + b.re * a.im };
    return r;
}
static inline cf c_scale(float s, cf a) { cf r = { a.re * s, a.im * s }; return r; }
static inline cf c_conj(cf a) { cf r = { a.re, -a.im }; return r; }

static void *xalloc(size_t bytes)
{
    void *p = NULL;
    if (bytes == 0) bytes = 64;
    if (posix_memalign(&p, 64, bytes) != 0) return NULL;
    return p;
}

/* exp(i*pi*dir*k/p), evaluated in double and rounded once (fft.c:102-106). */
static cf unit_root(int dir, int k, int p)
{
    double phase = (M_PI * dir * k) / p;
    cf r = { (float)cos(phase), (float)sin(phase) };
    return r;
}

/* Where the twiddles of butterfly span `p` start inside the table.  The reference lays the
 * levels out as [p=1][p=2,2 entries][one pad][p=4]...[p=N/2] (fft.c:122-131), i.e. level 1 at
 * 0, level 2 at 1 and every level p>=4 at offset p. */
static inline unsigned level_at(unsigned p) { return p == 1 ? 0u : (p == 2 ? 1u : p); }

static cf *make_twiddles(unsigned n, int dir)
{
    cf *t = xalloc((size_t)(n < 4 ? 4 : n) * sizeof(cf));
    if (!t) return NULL;
    memset(t, 0, (size_t)(n < 4 ? 4 : n) * sizeof(cf));
    for (unsigned p = 1; p < n; p <<= 1)
        for (unsigned k = 0; k < p; k++)
            t[level_at(p) + k] = unit_root(dir, (int)k, (int)p);
    return t;
}

/* (0 + i*dir) * exp(i*pi*dir*k/n), k < n  (fft.c:441-455). */
static cf *make_resolve_twiddles(int dir, unsigned n)
{
    cf *t = xalloc((size_t)n * sizeof(cf));
    if (!t) return NULL;
    cf j = { 0.0f, (float)dir };
    for (unsigned k = 0; k < n; k++) t[k] = c_mul(j, unit_root(dir, (int)k, (int)n));
    return t;
}

/* ---- one Stockham pass -----------------------------------------------------------------
 * `lines` points along the transform axis, `width` independent transforms side by side,
 * `stride` elements between consecutive points of one transform.  1D: width 1, stride 1.
 * Restates kernel.c:84-453 (horizontal) and kernel.c:455-738 (vertical): butterfly i reads
 * in[i + m*lines/radix] and writes out[(i-k)*radix + k + q*p], k = i mod p.  When p == 1 the
 * first-level multiplies are skipped exactly as the *_p1 kernels skip them.  `half` = the
 * *_half_p1 kernels (kernel.c:104-122,200-226,333-387): the upper half of the input is zero
 * and is never read. */
static void stockham_pass(cf *out, const cf *in, const cf *table, unsigned radix, unsigned p,
                          unsigned lines, unsigned width, unsigned stride, int half)
{
    const unsigned q = lines / radix;
    const cf *t1 = table + level_at(p);
    const cf *t2 = table + level_at(2 * p);
    const cf *t4 = table + level_at(4 * p);
    const cf zero = { 0.0f, 0.0f };

    for (unsigned i = 0; i < q; i++) {
        const unsigned k = i & (p - 1);
        const size_t j = ((size_t)(i - k) * radix + k) * stride;
        const size_t ostep = (size_t)p * stride;
        const size_t istep = (size_t)q * stride;
        const cf *src = in + (size_t)i * stride;
        cf *dst = out + j;

        for (unsigned x = 0; x < width; x++) {
            cf v[8];
            for (unsigned m = 0; m < radix; m++)
                v[m] = (half && m >= radix / 2) ? zero : src[m * istep + x];

            if (radix == 2) {
                cf b = p == 1 ? v[1] : c_mul(t1[k], v[1]);
                dst[x] = c_add(v[0], b);
                dst[x + ostep] = c_sub(v[0], b);
            } else if (radix == 4) {
                cf a = v[0], b = v[1], c = v[2], d = v[3];
                if (p != 1) { c = c_mul(t1[k], c); d = c_mul(t1[k], d); }
                cf r0 = c_add(a, c), r1 = c_sub(a, c), r2 = c_add(b, d), r3 = c_sub(b, d);
                if (p != 1) r2 = c_mul(r2, t2[k]);
                r3 = c_mul(r3, t2[k + p]);
                dst[x] = c_add(r0, r2);
                dst[x + ostep] = c_add(r1, r3);
                dst[x + 2 * ostep] = c_sub(r0, r2);
                dst[x + 3 * ostep] = c_sub(r1, r3);
            } else {
                cf a = v[0], b = v[1], c = v[2], d = v[3], e = v[4], f = v[5], g = v[6], h = v[7];
                if (p != 1) {
                    e = c_mul(t1[k], e); f = c_mul(t1[k], f);
                    g = c_mul(t1[k], g); h = c_mul(t1[k], h);
                }
                cf r0 = c_add(a, e), r1 = c_sub(a, e), r2 = c_add(b, f), r3 = c_sub(b, f);
                cf r4 = c_add(c, g), r5 = c_sub(c, g), r6 = c_add(d, h), r7 = c_sub(d, h);
                if (p != 1) { r4 = c_mul(r4, t2[k]); r6 = c_mul(r6, t2[k]); }
                r5 = c_mul(r5, t2[k + p]);
                r7 = c_mul(r7, t2[k + p]);
                a = c_add(r0, r4); b = c_add(r1, r5); c = c_sub(r0, r4); d = c_sub(r1, r5);
                e = c_add(r2, r6); f = c_add(r3, r7); g = c_sub(r2, r6); h = c_sub(r3, r7);
                if (p != 1) e = c_mul(e, t4[k]);
                f = c_mul(f, t4[k + p]);
                g = c_mul(g, t4[k + 2 * p]);
                h = c_mul(h, t4[k + 3 * p]);
                dst[x] = c_add(a, e);
                dst[x + ostep] = c_add(b, f);
                dst[x + 2 * ostep] = c_add(c, g);
                dst[x + 3 * ostep] = c_add(d, h);
                dst[x + 4 * ostep] = c_sub(a, e);
                dst[x + 5 * ostep] = c_sub(b, f);
                dst[x + 6 * ostep] = c_sub(c, g);
                dst[x + 7 * ostep] = c_sub(d, h);
            }
        }
    }
}

/* X[k] from the half-length spectrum Z (kernel.c:47-82).  n = complex length. */
static void resolve_r2c(cf *out, const cf *in, const cf *tw, unsigned n, int full)
{
    cf fe0 = { in[0].re, 0.0f }, fo0 = { in[0].im, 0.0f };
    out[0] = c_add(fe0, fo0);
    out[n] = c_sub(fe0, fo0);
    for (unsigned i = 1; i < n; i++) {
        cf a = in[i], b = c_conj(in[n - i]);
        cf fe = c_add(a, b);
        cf fo = c_mul(tw[i], c_sub(a, b));
        out[i] = c_scale(0.5f, c_add(fe, fo));
        if (full) out[i + n] = c_scale(0.5f, c_sub(fe, fo));
    }
}

/* z[k] for the half-length inverse from X[0..n] (kernel.c:36-45). */
static void resolve_c2r(cf *out, const cf *in, const cf *tw, unsigned n)
{
    for (unsigned i = 0; i < n; i++) {
        cf a = in[i], b = c_conj(in[n - i]);
        out[i] = c_add(c_add(a, b), c_mul(c_sub(a, b), tw[i]));
    }
}

/* kernel.c:22-34 */
void oracle_convolve(void *output, const void *a_, const void *b_, float normalization, unsigned samples)
{
    cf *out = output;
    const cf *a = a_, *b = b_;
    for (unsigned i = 0; i < samples; i++) out[i] = c_scale(normalization, c_mul(a[i], b[i]));
}

/* ---- schedules ---------------------------------------------------------------------------
 * The reference picks passes by a first-fit scan of a per-ISA table (fft.c:217-287, 347-376,
 * 401-432).  For the scalar table the scan reduces to the rules below.  Note the quirk that
 * inverse plans start with radix-2 passes because the direction-agnostic radix2_p1 entry
 * precedes the inverse radix-8/4 p==1 entries (fft.c:223-224 vs 233-236, 264-269). */
#define MAX_STEPS 40
typedef struct { unsigned radix, p; } step_t;

static unsigned schedule_1d(step_t *steps, unsigned n, int dir, int zero_pad)
{
    unsigned count = 0, rem = n, p = 1;
    while (rem > 1) {
        unsigned r;
        if (p == 1) {
            if (dir == FWD) r = rem % 8 == 0 ? 8 : (rem % 4 == 0 ? 4 : 2);
            else r = 2; /* radix2_p1 / radix2_half_p1 come first for inverse plans */
            (void)zero_pad;
        } else if (p == 2) {
            r = 2; /* forward/inverse_radix2_p2: the only entries that accept p == 2 */
        } else if (p >= 8 && rem % 8 == 0) r = 8;
        else if (rem % 4 == 0) r = 4;
        else r = 2;
        steps[count].radix = r; steps[count].p = p; count++;
        rem /= r; p *= r;
    }
    return count;
}

static unsigned schedule_vertical(step_t *steps, unsigned n, int dir)
{
    unsigned count = 0, rem = n, p = 1;
    while (rem > 1) {
        unsigned r;
        if (p == 1) {
            if (dir == FWD) r = rem % 8 == 0 ? 8 : (rem % 4 == 0 ? 4 : 2);
            else r = 2;
        } else if (p >= 8 && rem % 8 == 0) r = 8;
        else if (p >= 4 && rem % 4 == 0) r = 4;
        else r = 2; /* radix2_generic_vert accepts p >= 2 (fft.c:274-275) */
        steps[count].radix = r; steps[count].p = p; count++;
        rem /= r; p *= r;
    }
    return count;
}

static int bad_size(unsigned n) { return (n & (n - 1)) != 0 || n == 1; }

/* ---- 1D ----------------------------------------------------------------------------------- */
enum { K_C2C = 0, K_R2C, K_R2C_FULL, K_C2R };

struct oracle_plan_1d {
    unsigned n; /* complex length */
    int kind, zero_pad;
    unsigned nsteps;
    step_t steps[MAX_STEPS];
    cf *tw, *rtw, *tmp;
};

void oracle_free_plan_1d(oracle_plan_1d *plan)
{
    if (!plan) return;
    free(plan->tw); free(plan->rtw); free(plan->tmp); free(plan);
}

/* fft.c:626-662 */
oracle_plan_1d *oracle_create_plan_1d_c2c(unsigned N, int direction, unsigned flags)
{
    if (bad_size(N) || N == 0) return NULL;
    if (direction != FWD && direction != INV) return NULL;
    oracle_plan_1d *plan = calloc(1, sizeof(*plan));
    if (!plan) return NULL;
    plan->n = N;
    plan->kind = K_C2C;
    plan->zero_pad = (flags & F_ZERO_PAD) != 0;
    plan->tw = make_twiddles(N, direction);
    plan->tmp = xalloc((size_t)N * sizeof(cf));
    plan->nsteps = schedule_1d(plan->steps, N, direction, plan->zero_pad);
    if (!plan->tw || !plan->tmp) { oracle_free_plan_1d(plan); return NULL; }
    return plan;
}

/* fft.c:476-513 */
oracle_plan_1d *oracle_create_plan_1d_r2c(unsigned N, unsigned flags)
{
    if (bad_size(N) || N == 0) return NULL;
    oracle_plan_1d *plan = oracle_create_plan_1d_c2c(N / 2, FWD, flags);
    if (!plan) return NULL;
    plan->kind = (flags & F_FULL_R2C) ? K_R2C_FULL : K_R2C;
    plan->rtw = make_resolve_twiddles(FWD, N / 2);
    if (!plan->rtw) { oracle_free_plan_1d(plan); return NULL; }
    return plan;
}

/* fft.c:515-547 */
oracle_plan_1d *oracle_create_plan_1d_c2r(unsigned N, unsigned flags)
{
    if (bad_size(N) || N == 0) return NULL;
    oracle_plan_1d *plan = oracle_create_plan_1d_c2c(N / 2, INV, flags);
    if (!plan) return NULL;
    plan->kind = K_C2R;
    plan->rtw = make_resolve_twiddles(INV, N / 2);
    if (!plan->rtw) { oracle_free_plan_1d(plan); return NULL; }
    return plan;
}

unsigned oracle_plan_1d_steps(const oracle_plan_1d *plan, unsigned *radix, unsigned *p, unsigned cap)
{
    for (unsigned i = 0; i < plan->nsteps && i < cap; i++) { radix[i] = plan->steps[i].radix; p[i] = plan->steps[i].p; }
    return plan->nsteps;
}

/* The ping-pong driver of fft.c:811-850: the pass count decides which buffer goes first so
 * that the last write lands in `output`; `output` doubles as scratch. */
void oracle_execute_plan_1d(oracle_plan_1d *plan, void *output, const void *input)
{
    const unsigned n = plan->n;
    const int r2c = plan->kind == K_R2C || plan->kind == K_R2C_FULL;
    cf *out = output, *in = plan->tmp;
    unsigned total = plan->nsteps + (r2c ? 1u : 0u);
    if (total & 1u) { cf *t = out; out = in; in = t; }

    if (plan->kind == K_C2R) {
        resolve_c2r(out, input, plan->rtw, n);
        stockham_pass(in, out, plan->tw, plan->steps[0].radix, 1, n, 1, 1, plan->zero_pad);
    } else {
        stockham_pass(in, input, plan->tw, plan->steps[0].radix, 1, n, 1, 1, plan->zero_pad);
    }
    for (unsigned s = 1; s < plan->nsteps; s++) {
        stockham_pass(out, in, plan->tw, plan->steps[s].radix, plan->steps[s].p, n, 1, 1, 0);
        cf *t = out; out = in; in = t;
    }
    if (r2c) resolve_r2c(out, in, plan->rtw, n, plan->kind == K_R2C_FULL);
}

void oracle_execute_plan_1d_many(oracle_plan_1d *plan, void *output, const void *input,
                                 size_t count, size_t out_stride_bytes, size_t in_stride_bytes)
{
    for (size_t b = 0; b < count; b++)
        oracle_execute_plan_1d(plan, (char *)output + b * out_stride_bytes,
                               (const char *)input + b * in_stride_bytes);
}

/* ---- convolution (fft.c:565-624, 794-809) -------------------------------------------------- */
struct oracle_plan_conv {
    oracle_plan_1d *in_plan[2], *out_plan;
    size_t block_bytes;
    cf *product;
    float norm;
    unsigned multiply_n;
};

void oracle_free_plan_conv(oracle_plan_conv *plan)
{
    if (!plan) return;
    oracle_free_plan_1d(plan->in_plan[0]);
    oracle_free_plan_1d(plan->in_plan[1]);
    oracle_free_plan_1d(plan->out_plan);
    free(plan->product);
    free(plan);
}

oracle_plan_conv *oracle_create_plan_conv(unsigned N, unsigned flags, unsigned method)
{
    if (bad_size(N) || N == 0) return NULL;
    oracle_plan_conv *c = calloc(1, sizeof(*c));
    if (!c) return NULL;
    unsigned zp0 = (method & M_ZP_FIRST) ? F_ZERO_PAD : 0, zp1 = (method & M_ZP_SECOND) ? F_ZERO_PAD : 0;
    if (method & M_STEREO) {
        c->block_bytes = (size_t)N * sizeof(cf);
        c->in_plan[0] = oracle_create_plan_1d_c2c(N, FWD, flags | zp0);
        c->in_plan[1] = oracle_create_plan_1d_r2c(N, flags | zp1 | F_FULL_R2C);
        c->out_plan = oracle_create_plan_1d_c2c(N, INV, flags);
        c->multiply_n = N;
    } else {
        c->block_bytes = ((size_t)N / 2 + PAD_COMPLEX) * sizeof(cf);
        c->in_plan[0] = oracle_create_plan_1d_r2c(N, flags | zp0);
        c->in_plan[1] = oracle_create_plan_1d_r2c(N, flags | zp1);
        c->out_plan = oracle_create_plan_1d_c2r(N, flags);
        c->multiply_n = N / 2 + 1;
    }
    c->norm = 1.0f / N;
    c->product = xalloc(c->block_bytes);
    if (!c->in_plan[0] || !c->in_plan[1] || !c->out_plan || !c->product) { oracle_free_plan_conv(c); return NULL; }
    memset(c->product, 0, c->block_bytes);
    return c;
}

void oracle_execute_conv_input(oracle_plan_conv *plan, unsigned block, void *output, const void *input)
{
    oracle_execute_plan_1d(plan->in_plan[block], output, input);
}

size_t oracle_conv_get_transformed_block_size(oracle_plan_conv *plan) { return plan->block_bytes; }

void oracle_execute_conv_output(oracle_plan_conv *plan, void *output, const void *first, const void *second)
{
    oracle_convolve(plan->product, first, second, plan->norm, plan->multiply_n);
    oracle_execute_plan_1d(plan->out_plan, output, plan->product);
}

/* ---- 2D (fft.c:664-792, 852-989) ------------------------------------------------------------- */
struct oracle_plan_2d {
    unsigned nx, ny; /* nx = complex row length (Nx/2 for R2C/C2R) */
    int kind;
    unsigned vertical_nx;
    unsigned nsteps_x, nsteps_y;
    step_t steps_x[MAX_STEPS], steps_y[MAX_STEPS];
    cf *tw_x, *tw_y, *rtw, *tmp;
};

void oracle_free_plan_2d(oracle_plan_2d *plan)
{
    if (!plan) return;
    free(plan->tw_x); free(plan->tw_y); free(plan->rtw); free(plan->tmp); free(plan);
}

static oracle_plan_2d *plan_2d_common(unsigned nx, unsigned ny, int direction, int real_rows)
{
    oracle_plan_2d *plan = calloc(1, sizeof(*plan));
    if (!plan) return NULL;
    plan->nx = nx; plan->ny = ny; plan->vertical_nx = nx; plan->kind = K_C2C;
    plan->tw_x = make_twiddles(nx, direction);
    plan->tw_y = make_twiddles(ny, direction);
    plan->tmp = xalloc((size_t)(real_rows ? 2 : 1) * nx * ny * sizeof(cf)); /* fft.c:684-691 */
    plan->nsteps_x = schedule_1d(plan->steps_x, nx, direction, 0);
    plan->nsteps_y = schedule_vertical(plan->steps_y, ny, direction);
    if (!plan->tw_x || !plan->tw_y || !plan->tmp) { oracle_free_plan_2d(plan); return NULL; }
    return plan;
}

oracle_plan_2d *oracle_create_plan_2d_c2c(unsigned Nx, unsigned Ny, int direction, unsigned flags)
{
    (void)flags;
    if (bad_size(Nx) || bad_size(Ny) || Nx == 0 || Ny == 0) return NULL;
    if (direction != FWD && direction != INV) return NULL;
    return plan_2d_common(Nx, Ny, direction, 0);
}

oracle_plan_2d *oracle_create_plan_2d_r2c(unsigned Nx, unsigned Ny, unsigned flags)
{
    if (bad_size(Nx) || bad_size(Ny) || Nx == 0 || Ny == 0 || bad_size(Nx / 2)) return NULL;
    oracle_plan_2d *plan = plan_2d_common(Nx / 2, Ny, FWD, 1);
    if (!plan) return NULL;
    plan->kind = (flags & F_FULL_R2C) ? K_R2C_FULL : K_R2C;
    plan->vertical_nx = (flags & F_FULL_R2C) ? Nx : Nx / 2 + 1; /* fft.c:738-746 */
    plan->rtw = make_resolve_twiddles(FWD, Nx / 2);
    if (!plan->rtw) { oracle_free_plan_2d(plan); return NULL; }
    return plan;
}

oracle_plan_2d *oracle_create_plan_2d_c2r(unsigned Nx, unsigned Ny, unsigned flags)
{
    (void)flags;
    if (bad_size(Nx) || bad_size(Ny) || Nx == 0 || Ny == 0 || bad_size(Nx / 2)) return NULL;
    oracle_plan_2d *plan = plan_2d_common(Nx / 2, Ny, INV, 1);
    if (!plan) return NULL;
    plan->kind = K_C2R;
    plan->rtw = make_resolve_twiddles(INV, Nx / 2);
    if (!plan->rtw) { oracle_free_plan_2d(plan); return NULL; }
    return plan;
}

#define SWAP_CF(a, b) do { cf *t_ = (a); (a) = (b); (b) = t_; } while (0)

void oracle_execute_plan_2d(oracle_plan_2d *plan, void *output, const void *input_)
{
    const cf *input = input_;
    const unsigned nx = plan->nx, ny = plan->ny;

    if (plan->kind == K_C2R) {
        /* Columns first, on nx+1 columns of rows pitched 2*nx; then the per-row resolve; then
         * the row transforms (fft.c:863-916). */
        cf *hout = output, *hin = plan->tmp;
        if (plan->nsteps_x & 1u) SWAP_CF(hout, hin);
        cf *out = hout, *in = hin;
        if ((plan->nsteps_y & 1u) == 0) SWAP_CF(out, in);

        stockham_pass(in, input, plan->tw_y, plan->steps_y[0].radix, 1, ny, nx + 1, 2 * nx, 0);
        for (unsigned s = 1; s < plan->nsteps_y; s++) {
            stockham_pass(out, in, plan->tw_y, plan->steps_y[s].radix, plan->steps_y[s].p, ny, nx + 1, 2 * nx, 0);
            SWAP_CF(out, in);
        }
        for (unsigned y = 0; y < ny; y++)
            resolve_c2r(hout + (size_t)y * nx, hin + (size_t)2 * y * nx, plan->rtw, nx);
        for (unsigned y = 0; y < ny; y++) {
            cf *tin = hin, *tout = hout;
            stockham_pass(tin + (size_t)y * nx, tout + (size_t)y * nx, plan->tw_x, plan->steps_x[0].radix, 1, nx, 1, 1, 0);
            for (unsigned s = 1; s < plan->nsteps_x; s++) {
                stockham_pass(tout + (size_t)y * nx, tin + (size_t)y * nx, plan->tw_x,
                              plan->steps_x[s].radix, plan->steps_x[s].p, nx, 1, 1, 0);
                SWAP_CF(tout, tin);
            }
        }
        return;
    }

    /* C2C / R2C: rows, (resolve into doubled pitch), columns (fft.c:917-988). */
    const int r2c = plan->kind != K_C2C;
    const unsigned pitch_y = r2c ? 2 * nx : nx;
    cf *hout = output, *hin = plan->tmp;
    if ((plan->nsteps_y & 1u) == 0) SWAP_CF(hout, hin);
    cf *out = hin, *in = hout;
    unsigned row_steps = plan->nsteps_x + (r2c ? 1u : 0u);
    if (row_steps & 1u) SWAP_CF(out, in);

    for (unsigned y = 0; y < ny; y++) {
        cf *tin = in, *tout = out;
        stockham_pass(tin + (size_t)y * nx, input + (size_t)y * nx, plan->tw_x, plan->steps_x[0].radix, 1, nx, 1, 1, 0);
        for (unsigned s = 1; s < plan->nsteps_x; s++) {
            stockham_pass(tout + (size_t)y * nx, tin + (size_t)y * nx, plan->tw_x,
                          plan->steps_x[s].radix, plan->steps_x[s].p, nx, 1, 1, 0);
            SWAP_CF(tout, tin);
        }
    }
    if (r2c)
        for (unsigned y = 0; y < ny; y++)
            resolve_r2c(hin + (size_t)2 * y * nx, hout + (size_t)y * nx, plan->rtw, nx, plan->kind == K_R2C_FULL);

    const unsigned cols = plan->vertical_nx;
    stockham_pass(hout, hin, plan->tw_y, plan->steps_y[0].radix, 1, ny, cols, pitch_y, 0);
    SWAP_CF(hout, hin);
    for (unsigned s = 1; s < plan->nsteps_y; s++) {
        stockham_pass(hout, hin, plan->tw_y, plan->steps_y[s].radix, plan->steps_y[s].p, ny, cols, pitch_y, 0);
        SWAP_CF(hout, hin);
    }
}
