// One translation unit per image shape: compiled with -DF2D_NX=<nx> -DF2D_NY=<ny>.  Exposes
// f2d_launch_<nx>x<ny>(args, stream) (every kind, both directions of C2C) to dispatch.cu.
#include "fft2d_kernel.cuh"
#include "launch_util.cuh"

#if !defined(F2D_NX) || !defined(F2D_NY)
#error "compile with -DF2D_NX=<nx> -DF2D_NY=<ny>"
#endif

#define MUFFT_CAT4_(a, b, c, d) a##b##c##d
#define MUFFT_CAT4(a, b, c, d) MUFFT_CAT4_(a, b, c, d)

namespace mufft {

// lag L: the last producer ticket of image i and its first consumer ticket are L*(NP+NQ)-NP tickets
// apart; keep that above ~1.5x the tickets in flight (one per CTA).  Measured for 1024 x 1024 x 256 on 148
// SMs (lag 1 / 2 / 3 / 4 / 6: 1.395 / 1.247 / 1.196 / 1.217 / 1.382 ms): the rule's 3 is the optimum.
static unsigned default_lag(unsigned grid, unsigned per_image) { return 1 + (3 * grid / 2 + per_image - 1) / per_image; }

// Shapes the fused kernel cannot take with the configured CTA size report "unsupported" to the
// host layer (which then uses the row kernel + column kernel path) instead of failing the build.
template <int DIR, int KIND>
static cudaError_t launch_kind(const mcu_2d_args &a, cudaStream_t stream, unsigned *lag_only)
{
  if constexpr (!f2d_supported(F2D_NX, F2D_NY)) {
    (void)a; (void)stream; (void)lag_only;
    return cudaErrorInvalidValue;
  } else {
    static launch_cfg cfg;
    auto kernel = fft2d_fused_kernel<F2D_NX, F2D_NY, DIR, KIND>;
    constexpr size_t smem = f2d_smem_bytes(F2D_NX, F2D_NY, KIND);
    constexpr unsigned per_image = f2d_tickets_per_image(F2D_NX, F2D_NY, KIND);
    int ctas = 0;
    cudaError_t err = launch_prepare(cfg, kernel, smem, MUFFT_F2D_THREADS, &ctas);
    if (err != cudaSuccess) return err;
    if (lag_only != nullptr) {
        *lag_only = default_lag((unsigned)ctas, per_image);
        return cudaSuccess;
    }
    if (a.batch == 0) return cudaSuccess;
    const unsigned long long tickets = (unsigned long long)a.batch * per_image;
    if (tickets + (unsigned long long)ctas >= 0xffffffffull) return cudaErrorInvalidConfiguration; // every CTA draws one ticket past the end
    const unsigned grid = tickets < (unsigned long long)ctas ? (unsigned)tickets : (unsigned)ctas;
    mcu_2d_args b = a;
    if (b.lag == 0) b.lag = default_lag(grid, per_image);
    if (KIND == F2D_C2R) {
        // the ring must hold more images than the consumers run behind (or simply every image of the batch)
        const size_t lag_eff = b.lag < b.batch ? b.lag : b.batch;
        if (b.mid == nullptr || b.ring == 0 || !(b.ring > lag_eff || b.ring >= b.batch)) return cudaErrorInvalidValue;
    }
    kernel<<<grid, MUFFT_F2D_THREADS, smem, stream>>>(b);
    return cudaGetLastError();
  }
}

template <typename F>
static cudaError_t dispatch_kind(const mcu_2d_args &a, F &&f)
{
    switch (a.kind) {
    case MCU_2D_C2C: return a.direction < 0 ? f(std::integral_constant<int, -1>{}, std::integral_constant<int, F2D_C2C>{})
                                            : f(std::integral_constant<int, 1>{}, std::integral_constant<int, F2D_C2C>{});
    case MCU_2D_R2C: return a.direction < 0 ? f(std::integral_constant<int, -1>{}, std::integral_constant<int, F2D_R2C>{}) : cudaErrorInvalidValue;
    case MCU_2D_R2C_FULL: return a.direction < 0 ? f(std::integral_constant<int, -1>{}, std::integral_constant<int, F2D_R2C_FULL>{}) : cudaErrorInvalidValue;
    case MCU_2D_C2R: return a.direction > 0 ? f(std::integral_constant<int, 1>{}, std::integral_constant<int, F2D_C2R>{}) : cudaErrorInvalidValue;
    default: return cudaErrorInvalidValue;
    }
}

bool MUFFT_CAT4(f2d_supported_, F2D_NX, x, F2D_NY)() { return f2d_supported(F2D_NX, F2D_NY); }
int MUFFT_CAT4(f2d_rows_variant_, F2D_NX, x, F2D_NY)(int kind) { return f2d_rows_variant(F2D_NX, kind); }

unsigned MUFFT_CAT4(f2d_lag_, F2D_NX, x, F2D_NY)(const mcu_2d_args &a)
{
    unsigned lag = 0;
    const cudaError_t err = dispatch_kind(a, [&](auto dir, auto kind) { return launch_kind<decltype(dir)::value, decltype(kind)::value>(a, nullptr, &lag); });
    if (err != cudaSuccess) (void)cudaGetLastError();
    return err == cudaSuccess ? lag : 0;
}

cudaError_t MUFFT_CAT4(f2d_launch_, F2D_NX, x, F2D_NY)(const mcu_2d_args &a, cudaStream_t stream)
{
    return dispatch_kind(a, [&](auto dir, auto kind) { return launch_kind<decltype(dir)::value, decltype(kind)::value>(a, stream, nullptr); });
}

} // namespace mufft
