"""CPU tests (no GPU): the oracle (oracle/mufft_oracle.c, our C restatement) is pinned
  (a) bit-for-bit to committed outputs of the unmodified reference (tests/golden/, made by
      tests/golden/make_golden.py), which is what runs where /root/reference does not exist;
  (b) bit-for-bit to the reference library itself (oracle/_ref) on a wider sweep, when present;
  (c) to float64 numpy truth within the reference's own acceptance threshold (test.c:58,195,280).
"""
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden  # noqa: E402
from cpu_libs import (CONV_MONO, CONV_STEREO, CONV_ZP_FIRST, CONV_ZP_SECOND, FLAG_FULL_R2C, FLAG_NO_SIMD,  # noqa: E402
                      FLAG_ZERO_PAD, FORWARD, INVERSE, c_rand, rel_l2)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz")
CASES = list(make_golden.cases())


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def golden_input(golden, key):
    if key + "/in" in golden:
        return golden[key + "/in"]
    return golden[key + "/in0"], golden[key + "/in1"]


def golden_outputs(golden, key):
    outs = []
    while f"{key}/out{len(outs)}" in golden:
        outs.append(golden[f"{key}/out{len(outs)}"])
    return outs


def test_golden_file_covers_every_case(golden):
    for key, _, _ in CASES:
        assert f"{key}/out0" in golden, key


@pytest.mark.parametrize("key,kind,p", CASES, ids=[c[0] for c in CASES])
def test_oracle_equals_golden(oracle, golden, key, kind, p):
    x = golden_input(golden, key)
    got = make_golden.run(oracle, kind, p, x)
    got = got if isinstance(got, tuple) else (got,)
    want = golden_outputs(golden, key)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(g, w), f"{key}: oracle differs from the reference output"


def test_golden_inputs_are_the_reference_generator(golden):
    """The stored inputs are the srand(0)/rand() sequence of test.c:38-44."""
    r = c_rand(2 * 64)
    assert np.array_equal(golden["c2c_64_f/in"], (r[0::2] + 1j * r[1::2]).astype(np.complex64))


# ---- (b) against the reference library itself ---------------------------------------------------
def _cplx(n, seed):
    r = np.random.default_rng(seed).random(2 * n, np.float32) - np.float32(0.5)
    return (r[0::2] + 1j * r[1::2]).astype(np.complex64)


@pytest.mark.parametrize("n", [1 << k for k in range(1, 18)])
def test_oracle_equals_reference_1d(oracle, ref, n):
    x = _cplx(n, n)
    for d in (FORWARD, INVERSE):
        want = ref.c2c(x, d, FLAG_NO_SIMD)
        assert np.array_equal(oracle.c2c(x, d), want)
        assert np.array_equal(ref.c2c(x, d, 0), want)  # the AVX path gives the same bits
    if n >= 4:
        xr = x.view(np.float32)[:n].copy()
        for flags in (0, FLAG_FULL_R2C, FLAG_ZERO_PAD, FLAG_FULL_R2C | FLAG_ZERO_PAD):
            xin = xr[: n // 2] if flags & FLAG_ZERO_PAD else xr
            assert np.array_equal(oracle.r2c(xin, flags, n=n), ref.r2c(xin, flags | FLAG_NO_SIMD, n=n))
        spec = np.fft.rfft(xr.astype(np.float64)).astype(np.complex64)
        assert np.array_equal(oracle.c2r(spec, n), ref.c2r(spec, n, FLAG_NO_SIMD))


@pytest.mark.parametrize("n", [4, 8, 32, 512, 8192])
@pytest.mark.parametrize("method", [CONV_MONO, CONV_STEREO, CONV_MONO | CONV_ZP_FIRST | CONV_ZP_SECOND,
                                    CONV_STEREO | CONV_ZP_FIRST | CONV_ZP_SECOND, CONV_STEREO | CONV_ZP_FIRST])
def test_oracle_equals_reference_conv(oracle, ref, n, method):
    la = n // 2 if method & CONV_ZP_FIRST else n
    lb = n // 2 if method & CONV_ZP_SECOND else n
    a = _cplx(la, 1) if method & 1 else _cplx(la, 1).real.copy()
    b = _cplx(lb, 2).real.copy()
    got, want = oracle.conv(a, b, n, method), ref.conv(a, b, n, method, FLAG_NO_SIMD)
    used = n if method & 1 else n // 2 + 1
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1][:used], want[1][:used]) and np.array_equal(got[2][:used], want[2][:used])


@pytest.mark.parametrize("nx,ny", [(2, 2), (2, 8), (8, 2), (16, 16), (64, 8), (8, 128), (256, 256), (1024, 4)])
def test_oracle_equals_reference_2d(oracle, ref, nx, ny):
    x = _cplx(nx * ny, 3).reshape(ny, nx)
    for d in (FORWARD, INVERSE):
        assert np.array_equal(oracle.c2c_2d(x, d), ref.c2c_2d(x, d, FLAG_NO_SIMD))
    if nx >= 4:
        xr = x.real.copy()
        w = nx // 2 + 1
        assert np.array_equal(oracle.r2c_2d(xr)[:, :w], ref.r2c_2d(xr, FLAG_NO_SIMD)[:, :w])
        assert np.array_equal(oracle.r2c_2d(xr, FLAG_FULL_R2C), ref.r2c_2d(xr, FLAG_FULL_R2C | FLAG_NO_SIMD))
        spec = np.zeros((ny, nx), np.complex64)
        spec[:, :w] = np.fft.rfft2(xr.astype(np.float64)).astype(np.complex64)
        assert np.array_equal(oracle.c2r_2d(spec), ref.c2r_2d(spec, FLAG_NO_SIMD))


# ---- (c) against float64 truth, with the reference's own thresholds -------------------------------
@pytest.mark.parametrize("n", [2, 8, 64, 1024, 4096, 65536])
def test_oracle_vs_float64(oracle, n):
    x = _cplx(n, 7)
    for d, truth in ((FORWARD, np.fft.fft(x.astype(np.complex128))), (INVERSE, np.fft.ifft(x.astype(np.complex128)) * n)):
        got = oracle.c2c(x, d)
        assert np.max(np.abs(got - truth)) < 1e-6 * math.sqrt(n)  # test.c:195
        assert rel_l2(got, truth) < 1e-6 * math.log2(n)


def test_oracle_conv_vs_naive(oracle):
    """test.c:375-403 checks the convolution against an O(N^2) sum."""
    n = 64
    a, b = c_rand(n // 2, 1), c_rand(n // 2, 2)
    got, _, _ = oracle.conv(a, b, n, CONV_MONO | CONV_ZP_FIRST | CONV_ZP_SECOND)
    want = np.convolve(a.astype(np.float64), b.astype(np.float64))
    assert np.max(np.abs(got[: n - 1] - want)) < 2e-6 * math.sqrt(n)  # test.c:431


def test_oracle_plan_shapes(oracle):
    """NULL on the same arguments as the reference (fft.c:478-481,517-520,628-631)."""
    for n in (1, 3, 12, 100):
        for kind in ("c2c", "r2c", "c2r"):
            assert oracle.plan_1d_is_null(kind, n)
    assert oracle.plan_1d_is_null("r2c", 2) and oracle.plan_1d_is_null("c2r", 2)
    assert not oracle.plan_1d_is_null("c2c", 2)
