"""Parity of the CUDA library against the CPU oracle, through the C ABI of include/fft.h.

Mirrors the reference's test.c: same size sweeps (test.c:492-575), same input distribution
(uniform [-0.5, 0.5), test.c:38-44), and both acceptance criteria:
  * north star: relative L2 error vs the reference algorithm <= 1e-6 * log2(N);
  * test.c:58,195,280: every element within 1e-6 * sqrt(N) (2e-6 * sqrt(N) for convolution,
    test.c:431,474) - checked here against the oracle AND against float64 numpy truth, so a
    bias shared with the oracle cannot hide.
"""
import math

import numpy as np
import pytest

from cpu_libs import (CONV_MONO, CONV_STEREO, CONV_ZP_FIRST, CONV_ZP_SECOND, FLAG_FULL_R2C, FLAG_ZERO_PAD,
                      FORWARD, INVERSE, c_rand, rel_l2)

pytestmark = pytest.mark.gpu

SIZES_1D = [1 << k for k in range(1, 18)]  # 2 .. 131072 (test.c sweeps to 65536; bench.c to 131072)


def tol_rel(n):
    return 1e-6 * max(1.0, math.log2(n))


def check(got, want, n, truth=None, factor=1.0):
    assert got.shape == want.shape
    assert np.all(np.isfinite(got))
    assert rel_l2(got, want) <= tol_rel(n), f"rel L2 {rel_l2(got, want):.3e} > {tol_rel(n):.3e}"
    assert np.max(np.abs(got - want)) < factor * 1e-6 * math.sqrt(n) + 1e-7
    if truth is not None:
        assert rel_l2(got, truth) <= tol_rel(n)


def cplx(n, seed=0):
    r = c_rand(2 * n, seed) if n <= 4096 else (np.random.default_rng(seed).random(2 * n, np.float32) - np.float32(0.5))
    return (r[0::2] + 1j * r[1::2]).astype(np.complex64)


def real(n, seed=0):
    return c_rand(n, seed) if n <= 8192 else (np.random.default_rng(seed).random(n, np.float32) - np.float32(0.5))


@pytest.mark.parametrize("n", SIZES_1D)
@pytest.mark.parametrize("direction", [FORWARD, INVERSE])
def test_c2c_1d(gpu, oracle, n, direction):
    x = cplx(n)
    want = oracle.c2c(x, direction)
    got = gpu.c2c(x, direction)
    x64 = x.astype(np.complex128)
    truth = np.fft.fft(x64) if direction == FORWARD else np.fft.ifft(x64) * n
    check(got, want, n, truth)


@pytest.mark.parametrize("n", SIZES_1D)
def test_c2c_1d_zero_pad(gpu, oracle, n):
    x = cplx(n // 2, seed=1)
    want = oracle.c2c(x, FORWARD, FLAG_ZERO_PAD, n=n)
    got = gpu.c2c(x, FORWARD, FLAG_ZERO_PAD, n=n)  # the input buffer really is only N/2 long
    check(got, want, n)


@pytest.mark.parametrize("n", SIZES_1D[1:])
@pytest.mark.parametrize("flags", [0, FLAG_FULL_R2C, FLAG_ZERO_PAD, FLAG_FULL_R2C | FLAG_ZERO_PAD])
def test_r2c_1d(gpu, oracle, n, flags):
    x = real(n // 2 if flags & FLAG_ZERO_PAD else n, seed=2)
    want = oracle.r2c(x, flags, n=n)
    got = gpu.r2c(x, flags, n=n)
    padded = np.zeros(n, np.float64)
    padded[: x.size] = x
    truth = np.fft.fft(padded)[: want.size]
    check(got, want, n, truth)
    if flags & FLAG_FULL_R2C:  # conjugate symmetry, test.c:295-302
        assert np.max(np.abs(got[1:] - np.conj(got[:0:-1]))) < 1e-6 * math.sqrt(n)


@pytest.mark.parametrize("n", SIZES_1D[1:])
def test_c2r_1d(gpu, oracle, n):
    spec = np.fft.rfft(real(n, seed=3).astype(np.float64)).astype(np.complex64)
    want = oracle.c2r(spec, n)
    got = gpu.c2r(spec, n)
    truth = np.fft.irfft(spec.astype(np.complex128), n) * n
    check(got, want, n * n, None)  # outputs are O(N) large: scale the absolute criterion
    assert rel_l2(got, want) <= tol_rel(n)
    assert rel_l2(got, truth) <= tol_rel(n)


@pytest.mark.parametrize("n", [4, 8, 16, 64, 256, 1024, 4096, 8192, 16384, 32768])
@pytest.mark.parametrize("method", [CONV_MONO, CONV_STEREO,
                                    CONV_MONO | CONV_ZP_FIRST | CONV_ZP_SECOND,
                                    CONV_STEREO | CONV_ZP_FIRST | CONV_ZP_SECOND,
                                    CONV_STEREO | CONV_ZP_FIRST, CONV_MONO | CONV_ZP_SECOND])
def test_conv(gpu, oracle, n, method):
    la = n // 2 if method & CONV_ZP_FIRST else n
    lb = n // 2 if method & CONV_ZP_SECOND else n
    a = cplx(la, seed=4) if method & 1 else real(la, seed=4)
    b = real(lb, seed=5)
    want, w0, w1 = oracle.conv(a, b, n, method)
    got, g0, g1 = gpu.conv(a, b, n, method)
    used = n if method & 1 else n // 2 + 1
    assert rel_l2(g0[:used], w0[:used]) <= tol_rel(n)
    assert rel_l2(g1[:used], w1[:used]) <= tol_rel(n)
    assert rel_l2(got, want) <= 2 * tol_rel(n)
    assert np.max(np.abs(got - want)) < 2e-6 * math.sqrt(n)
    A = np.zeros(n, np.complex128)
    A[:la] = a
    B = np.zeros(n, np.complex128)
    B[:lb] = b
    truth = np.fft.ifft(np.fft.fft(A) * np.fft.fft(B))
    assert rel_l2(got, truth if method & 1 else truth.real) <= 2 * tol_rel(n)


SHAPES_2D = [(2, 2), (4, 2), (2, 4), (8, 8), (16, 4), (4, 32), (64, 64), (128, 32), (32, 256), (512, 512),
             (1024, 16), (16, 1024), (256, 2048), (8, 4096), (16384, 4)]


@pytest.mark.parametrize("nx,ny", SHAPES_2D)
@pytest.mark.parametrize("direction", [FORWARD, INVERSE])
def test_c2c_2d(gpu, oracle, nx, ny, direction):
    x = cplx(nx * ny, seed=6).reshape(ny, nx)
    want = oracle.c2c_2d(x, direction)
    got = gpu.c2c_2d(x, direction)
    check(got, want, nx * ny)


@pytest.mark.parametrize("nx,ny", [s for s in SHAPES_2D if s[0] >= 4])
@pytest.mark.parametrize("flags", [0, FLAG_FULL_R2C])
def test_r2c_2d(gpu, oracle, nx, ny, flags):
    x = real(nx * ny, seed=7).reshape(ny, nx)
    want = oracle.r2c_2d(x, flags)
    got = gpu.r2c_2d(x, flags)
    w = nx if flags else nx // 2 + 1  # only these columns are defined (ref fft.h:238-243, test.c:99-106)
    check(got[:, :w], want[:, :w], nx * ny, np.fft.fft2(x.astype(np.float64))[:, :w])


@pytest.mark.parametrize("nx,ny", [s for s in SHAPES_2D if s[0] >= 4])
def test_c2r_2d(gpu, oracle, nx, ny):
    x = real(nx * ny, seed=8).reshape(ny, nx)
    spec = np.zeros((ny, nx), np.complex64)
    spec[:, : nx // 2 + 1] = np.fft.rfft2(x.astype(np.float64)).astype(np.complex64)
    want = oracle.c2r_2d(spec)
    got = gpu.c2r_2d(spec)
    assert rel_l2(got, want) <= tol_rel(nx * ny)
    assert rel_l2(got, x.astype(np.float64) * nx * ny) <= tol_rel(nx * ny)


def test_constructor_errors(gpu, oracle):
    """NULL on the same arguments as the reference (ref fft.c:478-481,517-520,567-570,628-631,666-669)."""
    for n in (0, 1, 3, 6, 12, 100, 1000):
        for kind in ("c2c", "r2c", "c2r"):
            assert gpu.plan_1d_is_null(kind, n)
    for kind in ("r2c", "c2r"):
        assert gpu.plan_1d_is_null(kind, 2)          # complex half would be 1 point
        assert not gpu.plan_1d_is_null(kind, 4)
    assert not gpu.plan_1d_is_null("c2c", 2)
    assert not gpu._create_plan_conv(2, 0, 0) and not gpu._create_plan_conv(24, 0, 1)
    assert not gpu._create_plan_2d_c2c(1, 8, -1, 0) and not gpu._create_plan_2d_c2c(8, 3, -1, 0)
    assert not gpu._create_plan_2d_r2c(2, 8, 0) and not gpu._create_plan_2d_c2r(2, 8, 0)
    gpu._free_plan_1d(None)
    gpu._free_plan_2d(None)
    gpu._free_plan_conv(None)


def test_flags_cpu_bits_ignored(gpu, oracle):
    x = cplx(256, seed=9)
    base = gpu.c2c(x, FORWARD, 0)
    for flags in range(8):  # test.c sweeps all NO_AVX|NO_SSE3|NO_SSE subsets
        assert np.array_equal(gpu.c2c(x, FORWARD, flags), base)


@pytest.mark.parametrize("samples", [5, 2049, 8192])
@pytest.mark.parametrize("where", ["pageable", "managed", "device"])
def test_get_convolve_func(oracle, samples, where):
    """mufft_get_convolve_func (ref fft.h:201-207): the returned function computes out = norm * a .* b
    (ref kernel.c:22-34) on whatever memory it is handed and is complete on return."""
    import ctypes as C

    import mufft_b200 as m

    if m.mufft_cuda_device_count() < 1:
        pytest.fail("no CUDA device visible")
    fn = m.mufft_get_convolve_func(0)
    assert fn
    a, b = cplx(samples, seed=10), cplx(samples, seed=11)
    norm = 1.0 / 4096.0
    want = np.zeros(samples, np.complex64)
    oracle.lib.oracle_convolve.argtypes = [C.c_void_p] * 3 + [C.c_float, C.c_uint]
    oracle.lib.oracle_convolve.restype = None
    oracle.lib.oracle_convolve(want.ctypes.data, a.ctypes.data, b.ctypes.data, norm, samples)
    nbytes = samples * 8
    if where == "pageable":
        got = np.zeros(samples, np.complex64)
        fn(got.ctypes.data, a.ctypes.data, b.ctypes.data, norm, samples)
    elif where == "managed":
        ptrs = [m.mufft_alloc(nbytes) for _ in range(3)]
        assert all(ptrs)
        views = [np.frombuffer((C.c_char * nbytes).from_address(p), np.complex64) for p in ptrs]
        views[1][:] = a
        views[2][:] = b
        fn(ptrs[0], ptrs[1], ptrs[2], norm, samples)
        got = views[0].copy()
        del views
        for p in ptrs:
            m.mufft_free(p)
    else:
        import torch
        da = torch.from_numpy(a.view(np.float32)).cuda()
        db = torch.from_numpy(b.view(np.float32)).cuda()
        do = torch.zeros_like(da)
        torch.cuda.synchronize()
        fn(do.data_ptr(), da.data_ptr(), db.data_ptr(), norm, samples)
        got = do.cpu().numpy().view(np.complex64)
    assert rel_l2(got, want) <= 1e-6
    assert np.max(np.abs(got - want)) <= 1e-7
    truth = norm * a.astype(np.complex128) * b.astype(np.complex128)
    assert rel_l2(got, truth) <= 1e-6
