/* TEST / BASELINE INFRASTRUCTURE — NOT PART OF THE PRODUCT.
 *
 * CPU baseline driver: an FFTW-free replica of the muFFT halves of the reference's bench.c
 * loops (bench.c:201-228 1D C2C, 270-310 stereo convolution, 312-344 R2C+C2R pair, 380-407
 * 2D C2C), linked against the UNMODIFIED reference objects built by oracle/Makefile into
 * oracle/_ref/.  The reference library is single threaded; to use all host cores this driver
 * runs one independent plan + buffers per thread over disjoint shards of the batch, which is
 * how a caller of the reference would parallelise a batched job.
 *
 * usage: ref_bench MODE N NY BATCH THREADS REPEATS [hot]
 *   MODE    c2c | r2c_c2r | conv_stereo | c2c2d
 *   N       transform length (Nx for c2c2d); NY rows for c2c2d (ignored otherwise)
 *   BATCH   independent transforms (total over all threads), each with its own memory
 *   REPEATS sweeps over the batch inside the timed region
 *   hot     optional: every transform reuses one buffer (what muFFT-bench itself measures)
 * Prints one JSON line with the wall time of the slowest thread and the derived throughput.
 * Input generator: srand(0), (float)rand()/RAND_MAX - 0.5f, as test.c:38-44 / bench.c.
 */
#define _GNU_SOURCE
#include "fft.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { float re, im; } cf;

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

enum { MODE_C2C, MODE_R2C_C2R, MODE_CONV_STEREO, MODE_C2C2D };

typedef struct {
    int mode, hot;
    unsigned n, ny;
    size_t count; /* transforms in this thread's shard */
    unsigned repeats;
    unsigned seed;
    pthread_barrier_t *barrier;
    double seconds;
    double checksum;
} job_t;

static void fill(float *p, size_t n, unsigned *seed)
{
    for (size_t i = 0; i < n; i++) p[i] = (float)rand_r(seed) / RAND_MAX - 0.5f;
}

static void *worker(void *arg)
{
    job_t *job = arg;
    const unsigned n = job->n, ny = job->ny;
    const size_t count = job->count ? job->count : 1;
    const size_t slots = job->hot ? 1 : count;
    double t0 = 0, t1 = 0, sum = 0;

    if (job->mode == MODE_C2C) {
        mufft_plan_1d *plan = mufft_create_plan_1d_c2c(n, MUFFT_FORWARD, 0);
        cf *in = mufft_alloc(slots * n * sizeof(cf)), *out = mufft_alloc(slots * n * sizeof(cf));
        fill((float *)in, slots * n * 2, &job->seed);
        memset(out, 0, slots * n * sizeof(cf));
        pthread_barrier_wait(job->barrier);
        t0 = now();
        for (unsigned r = 0; r < job->repeats; r++)
            for (size_t b = 0; b < job->count; b++) {
                size_t s = job->hot ? 0 : b;
                mufft_execute_plan_1d(plan, out + s * n, in + s * n);
            }
        t1 = now();
        sum = out[0].re + out[(slots - 1) * n + n / 2].im;
        mufft_free(in); mufft_free(out); mufft_free_plan_1d(plan);
    } else if (job->mode == MODE_R2C_C2R) {
        mufft_plan_1d *fwd = mufft_create_plan_1d_r2c(n, 0), *inv = mufft_create_plan_1d_c2r(n, 0);
        const size_t bins = n / 2 + 4; /* padded pitch */
        float *in = mufft_alloc(slots * n * sizeof(float)), *back = mufft_alloc(slots * n * sizeof(float));
        cf *spec = mufft_alloc(slots * bins * sizeof(cf));
        fill(in, slots * n, &job->seed);
        memset(back, 0, slots * n * sizeof(float));
        memset(spec, 0, slots * bins * sizeof(cf));
        pthread_barrier_wait(job->barrier);
        t0 = now();
        for (unsigned r = 0; r < job->repeats; r++)
            for (size_t b = 0; b < job->count; b++) {
                size_t s = job->hot ? 0 : b;
                mufft_execute_plan_1d(fwd, spec + s * bins, in + s * n);
                mufft_execute_plan_1d(inv, back + s * n, spec + s * bins);
            }
        t1 = now();
        sum = back[0] + back[(slots - 1) * n + 1];
        mufft_free(in); mufft_free(back); mufft_free(spec);
        mufft_free_plan_1d(fwd); mufft_free_plan_1d(inv);
    } else if (job->mode == MODE_CONV_STEREO) {
        /* FFT length n, blocks of n/2 stereo samples, both zero-pad flags (bench.c:270-310). */
        mufft_plan_conv *plan = mufft_create_plan_conv(n, 0,
                MUFFT_CONV_METHOD_FLAG_STEREO_MONO | MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_FIRST |
                MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_SECOND);
        const size_t blk = mufft_conv_get_transformed_block_size(plan) / sizeof(cf);
        cf *in = mufft_alloc(slots * (n / 2) * sizeof(cf));
        float *filt = mufft_alloc((n / 2) * sizeof(float));
        cf *spec = mufft_alloc(slots * blk * sizeof(cf)), *fspec = mufft_alloc(blk * sizeof(cf));
        cf *out = mufft_alloc(slots * n * sizeof(cf));
        fill((float *)in, slots * n, &job->seed);
        fill(filt, n / 2, &job->seed);
        memset(spec, 0, slots * blk * sizeof(cf));
        memset(out, 0, slots * n * sizeof(cf));
        mufft_execute_conv_input(plan, MUFFT_CONV_BLOCK_SECOND, fspec, filt);
        pthread_barrier_wait(job->barrier);
        t0 = now();
        for (unsigned r = 0; r < job->repeats; r++)
            for (size_t b = 0; b < job->count; b++) {
                size_t s = job->hot ? 0 : b;
                mufft_execute_conv_input(plan, MUFFT_CONV_BLOCK_FIRST, spec + s * blk, in + s * (n / 2));
                mufft_execute_conv_output(plan, out + s * n, spec + s * blk, fspec);
            }
        t1 = now();
        sum = out[0].re + out[(slots - 1) * n + 1].im;
        mufft_free(in); mufft_free(filt); mufft_free(spec); mufft_free(fspec); mufft_free(out);
        mufft_free_plan_conv(plan);
    } else {
        mufft_plan_2d *plan = mufft_create_plan_2d_c2c(n, ny, MUFFT_FORWARD, 0);
        const size_t img = (size_t)n * ny;
        cf *in = mufft_alloc(slots * img * sizeof(cf)), *out = mufft_alloc(slots * img * sizeof(cf));
        fill((float *)in, slots * img * 2, &job->seed);
        memset(out, 0, slots * img * sizeof(cf));
        pthread_barrier_wait(job->barrier);
        t0 = now();
        for (unsigned r = 0; r < job->repeats; r++)
            for (size_t b = 0; b < job->count; b++) {
                size_t s = job->hot ? 0 : b;
                mufft_execute_plan_2d(plan, out + s * img, in + s * img);
            }
        t1 = now();
        sum = out[0].re + out[(slots - 1) * img + 1].im;
        mufft_free(in); mufft_free(out); mufft_free_plan_2d(plan);
    }
    (void)count;
    job->seconds = t1 - t0;
    job->checksum = sum;
    return NULL;
}

int main(int argc, char **argv)
{
    if (argc < 7) {
        fprintf(stderr, "usage: %s c2c|r2c_c2r|conv_stereo|c2c2d N NY BATCH THREADS REPEATS [hot]\n", argv[0]);
        return 2;
    }
    int mode = !strcmp(argv[1], "c2c") ? MODE_C2C : !strcmp(argv[1], "r2c_c2r") ? MODE_R2C_C2R :
               !strcmp(argv[1], "conv_stereo") ? MODE_CONV_STEREO : !strcmp(argv[1], "c2c2d") ? MODE_C2C2D : -1;
    if (mode < 0) { fprintf(stderr, "unknown mode %s\n", argv[1]); return 2; }
    unsigned n = (unsigned)strtoul(argv[2], NULL, 10), ny = (unsigned)strtoul(argv[3], NULL, 10);
    size_t batch = strtoull(argv[4], NULL, 10);
    int threads = atoi(argv[5]);
    unsigned repeats = (unsigned)strtoul(argv[6], NULL, 10);
    int hot = argc > 7 && !strcmp(argv[7], "hot");
    if (threads < 1) threads = 1;
    if (repeats < 1) repeats = 1;

    pthread_barrier_t barrier;
    pthread_barrier_init(&barrier, NULL, (unsigned)threads);
    pthread_t *tid = calloc((size_t)threads, sizeof(*tid));
    job_t *jobs = calloc((size_t)threads, sizeof(*jobs));
    for (int t = 0; t < threads; t++) {
        size_t lo = batch * (size_t)t / (size_t)threads, hi = batch * (size_t)(t + 1) / (size_t)threads;
        jobs[t] = (job_t){ .mode = mode, .hot = hot, .n = n, .ny = ny, .count = hi - lo, .repeats = repeats,
                           .seed = (unsigned)t * 2654435761u + 1u, .barrier = &barrier };
        pthread_create(&tid[t], NULL, worker, &jobs[t]);
    }
    double worst = 0, sum = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(tid[t], NULL);
        if (jobs[t].seconds > worst) worst = jobs[t].seconds;
        sum += jobs[t].checksum;
    }
    double transforms = (double)batch * repeats;
    double lg = log2((double)n), flops;
    if (mode == MODE_C2C2D) flops = 5.0 * n * ny * (lg + log2((double)ny));
    else if (mode == MODE_CONV_STEREO) flops = 2.0 * 5.0 * n * lg + 6.0 * n;
    else flops = 5.0 * n * lg; /* an R2C+C2R pair counts as one 5NlogN (bench.c:500) */
    printf("{\"mode\": \"%s\", \"n\": %u, \"ny\": %u, \"batch\": %zu, \"threads\": %d, \"repeats\": %u, \"hot\": %d, "
           "\"seconds\": %.6f, \"transforms_per_s\": %.3f, \"gflops\": %.3f, \"us_per_transform_per_thread\": %.4f, "
           "\"checksum\": %.6g}\n",
           argv[1], n, ny, batch, threads, repeats, hot, worst, transforms / worst,
           transforms * flops / worst * 1e-9, worst * 1e6 * threads / transforms, sum);
    free(tid); free(jobs);
    return 0;
}
