// Host side of the engine: twiddle-table generation for every shape of engine_shape().  Shared by
// dispatch.cu (the product) and tests/emu (the CPU emulation of the CUDA side), so that both hand
// the kernels exactly the same tables.
#pragma once

#include "abi.h"
#include "fft_engine.cuh"
#include "rows_kernel.cuh"

#include <math.h>
#include <stdlib.h>

namespace mufft {

template <int N, int V>
static void fill_table(float *table, int direction)
{
    constexpr int passes = engine_num_passes(N, V);
    for (int j = 1; j < passes; j++) {
        const int r = engine_radix(N, j, V), p = engine_span(N, j, V);
        float *block = table + 2 * (size_t)engine_tw_offset(N, j, V);
        for (int m = 1; m < r; m++)
            for (int k = 0; k < p; k++) {
                // exp(dir * 2*pi*i * k*m / (r*p)), evaluated in double and rounded once, the way
                // the reference builds its tables (ref fft.c:102-106).
                const double phase = (M_PI * direction * 2.0 * (double)(k * m)) / (double)(r * p);
                block[2 * ((size_t)(m - 1) * p + k) + 0] = (float)cos(phase);
                block[2 * ((size_t)(m - 1) * p + k) + 1] = (float)sin(phase);
            }
    }
}

template <int N>
static void fill_table_variant(float *table, int direction, int variant)
{
    if (variant == 1) { if constexpr (engine_variant_exists(N, 1)) fill_table<N, 1>(table, direction); }
    else if (variant == 2) { if constexpr (engine_variant_exists(N, 2)) fill_table<N, 2>(table, direction); }
    else fill_table<N, 0>(table, direction);
}

#define MUFFT_FOR_CONV_SIZES(X) \
    X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192)
// row kernels also exist for MCU_ENGINE_BIG_N
#define MUFFT_FOR_ENGINE_SIZES(X) MUFFT_FOR_CONV_SIZES(X) X(16384)
#define MUFFT_FOR_TABLE_SIZES(X) MUFFT_FOR_ENGINE_SIZES(X)


static inline int tables_rows_variant(unsigned n, int in_mode, int out_mode)
{
    // the kernels' own choice (rows_kernel.cuh), so that the table handed to a kernel is always the
    // one of the shape it was compiled for
    switch (n) {
#define X(v) case v: return rows_variant(v, in_mode, out_mode);
        MUFFT_FOR_ENGINE_SIZES(X)
#undef X
    default: return 0;
    }
}

static inline size_t tables_size(unsigned n, int variant)
{
    switch (n) {
#define X(v) case v: return (size_t)(variant == 1 ? engine_tw_total(v, 1) : variant == 2 ? engine_tw_total(v, 2) : engine_tw_total(v, 0)) + 1;
        MUFFT_FOR_TABLE_SIZES(X)
#undef X
    default: return 0;
    }
}

static inline void tables_fill(float *table, unsigned n, int direction, int variant)
{
    switch (n) {
#define X(v) case v: fill_table_variant<v>(table, direction, variant); break;
        MUFFT_FOR_TABLE_SIZES(X)
#undef X
    default: break;
    }
}

// n = n1 * n2 with n1 <= 2048 (column kernel) and n2 in {256, 512, 1024} (row kernel B)
static inline unsigned tables_split_n2(unsigned n, int columns)
{
    if (n <= MCU_COLS_MAX_N || (n & (n - 1)) != 0) return 0;
    // MUFFT_CUDA_SPLIT_N2=<256|512|1024>: measurement knob, overrides the choice where it is legal
    static const unsigned forced = [] { const char *e = getenv("MUFFT_CUDA_SPLIT_N2"); return e ? (unsigned)atoi(e) : 0u; }();
    if ((forced == 256 || forced == 512 || forced == 1024) && n / forced >= 2 && n / forced <= MCU_COLS_MAX_N) return forced;
    // Row transforms: the widest second factor wins from 32768 points on (measured on B200,
    // profiles/README.md section 3: N = 32768 4.06 -> 3.52 ms, N = 2^20 6.00 -> 4.27 ms with 1024 instead
    // of 256 / 512).  Column transforms keep the narrow split they were measured with.
    if (!columns && n >= 32768u && n <= 2048u * 1024u) return 1024;
    if (n <= 2048u * 256u) return 256;
    if (n <= 2048u * 512u) return 512;
    if (n <= 2048u * 1024u) return 1024;
    return 0;
}


} // namespace mufft
