"""Generates tests/golden/reference_vectors.npz: inputs and the outputs of the UNMODIFIED reference
(oracle/_ref/libmufft_ref.so, built from /root/reference by oracle/Makefile) for a fixed list of
cases.  Run in the build container (where /root/reference exists):

    make -C oracle ref && python tests/golden/make_golden.py

The reference has no golden vectors of its own (its test.c compares against FFTW on the fly), so
these files are what pins the oracle - and through it the CUDA library - where the reference
sources are not available (the GPU box).  Every ISA path of the reference gives bit-identical
results on these cases (checked below), so one output per case is stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cpu_libs  # noqa: E402
from cpu_libs import (CONV_MONO, CONV_STEREO, CONV_ZP_FIRST, CONV_ZP_SECOND, FLAG_FULL_R2C, FLAG_NO_SIMD,  # noqa: E402
                      FLAG_ZERO_PAD, FORWARD, INVERSE)

C2C_SIZES = [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096]
REAL_SIZES = [4, 8, 16, 64, 256, 1024, 4096]
CONV_SIZES = [4, 16, 256, 2048]
SHAPES_2D = [(2, 2), (4, 2), (8, 4), (4, 16), (32, 16), (64, 64)]  # (Nx, Ny)


def cases():
    """(key, kind, params) for every golden case; inputs come from the glibc srand(0)/rand()
    generator of the reference's own tests (test.c:38-44)."""
    for n in C2C_SIZES:
        for d in (FORWARD, INVERSE):
            yield f"c2c_{n}_{'f' if d < 0 else 'i'}", "c2c", dict(n=n, direction=d, flags=0)
        yield f"c2c_{n}_zp", "c2c", dict(n=n, direction=FORWARD, flags=FLAG_ZERO_PAD)
    for n in REAL_SIZES:
        for flags in (0, FLAG_FULL_R2C, FLAG_ZERO_PAD, FLAG_FULL_R2C | FLAG_ZERO_PAD):
            yield f"r2c_{n}_{flags >> 16}", "r2c", dict(n=n, flags=flags)
        yield f"c2r_{n}", "c2r", dict(n=n)
    for n in CONV_SIZES:
        for method in (CONV_MONO, CONV_STEREO, CONV_MONO | CONV_ZP_FIRST | CONV_ZP_SECOND,
                       CONV_STEREO | CONV_ZP_FIRST | CONV_ZP_SECOND):
            yield f"conv_{n}_{method}", "conv", dict(n=n, method=method)
    for nx, ny in SHAPES_2D:
        for d in (FORWARD, INVERSE):
            yield f"c2c2d_{nx}x{ny}_{'f' if d < 0 else 'i'}", "c2c2d", dict(nx=nx, ny=ny, direction=d)
        if nx >= 4:
            for flags in (0, FLAG_FULL_R2C):
                yield f"r2c2d_{nx}x{ny}_{flags >> 16}", "r2c2d", dict(nx=nx, ny=ny, flags=flags)
            yield f"c2r2d_{nx}x{ny}", "c2r2d", dict(nx=nx, ny=ny)


def make_input(kind, p):
    if kind == "c2c":
        m = p["n"] // 2 if p["flags"] & FLAG_ZERO_PAD else p["n"]
        r = cpu_libs.c_rand(2 * m)
        return (r[0::2] + 1j * r[1::2]).astype(np.complex64)
    if kind == "r2c":
        return cpu_libs.c_rand(p["n"] // 2 if p["flags"] & FLAG_ZERO_PAD else p["n"])
    if kind == "c2r":
        # Hermitian input with real DC and Nyquist bins, as test.c:210-251 builds it
        r = cpu_libs.c_rand(p["n"] + 2)
        x = (r[0::2] + 1j * r[1::2]).astype(np.complex64)
        x[0] = x[0].real
        x[-1] = x[-1].real
        return x
    if kind == "conv":
        n, method = p["n"], p["method"]
        la = n // 2 if method & CONV_ZP_FIRST else n
        lb = n // 2 if method & CONV_ZP_SECOND else n
        r = cpu_libs.c_rand(2 * la + lb)
        a = (r[0:2 * la:2] + 1j * r[1:2 * la:2]).astype(np.complex64) if method & 1 else r[:la].copy()
        return a, r[2 * la:].copy()
    nx, ny = p["nx"], p["ny"]
    if kind == "c2c2d":
        r = cpu_libs.c_rand(2 * nx * ny)
        return (r[0::2] + 1j * r[1::2]).astype(np.complex64).reshape(ny, nx)
    if kind == "r2c2d":
        return cpu_libs.c_rand(nx * ny).reshape(ny, nx)
    if kind == "c2r2d":
        # pitch-Nx complex rows, columns 0..Nx/2 carry the spectrum of a real image (fft.h:255-257)
        img = cpu_libs.c_rand(nx * ny).reshape(ny, nx).astype(np.float64)
        spec = np.zeros((ny, nx), np.complex64)
        spec[:, : nx // 2 + 1] = np.fft.rfft2(img).astype(np.complex64)
        return spec
    raise ValueError(kind)


def run(lib, kind, p, x, extra_flags=0):
    if kind == "c2c":
        return lib.c2c(x, p["direction"], p["flags"] | extra_flags, n=p["n"])
    if kind == "r2c":
        return lib.r2c(x, p["flags"] | extra_flags, n=p["n"])
    if kind == "c2r":
        return lib.c2r(x, p["n"], extra_flags)
    if kind == "conv":
        out, s0, s1 = lib.conv(x[0], x[1], p["n"], p["method"], extra_flags)
        used = p["n"] if p["method"] & 1 else p["n"] // 2 + 1
        return out, s0[:used].copy(), s1[:used].copy()
    if kind == "c2c2d":
        return lib.c2c_2d(x, p["direction"], extra_flags)
    if kind == "r2c2d":
        out = lib.r2c_2d(x, p["flags"] | extra_flags)
        w = p["nx"] if p["flags"] & FLAG_FULL_R2C else p["nx"] // 2 + 1
        return out[:, :w].copy()  # the remaining columns are undefined (fft.h:238-243)
    if kind == "c2r2d":
        return lib.c2r_2d(x, extra_flags)
    raise ValueError(kind)


def main():
    ref = cpu_libs.load_ref()
    assert ref is not None, "needs oracle/_ref/libmufft_ref.so (make -C oracle ref)"
    store = {}
    for key, kind, p in cases():
        x = make_input(kind, p)
        simd = run(ref, kind, p, x)
        scalar = run(ref, kind, p, x, FLAG_NO_SIMD)
        parts_a = simd if isinstance(simd, tuple) else (simd,)
        parts_b = scalar if isinstance(scalar, tuple) else (scalar,)
        for a, b in zip(parts_a, parts_b):
            assert np.array_equal(a, b), f"{key}: AVX and scalar reference paths differ"
        if isinstance(x, tuple):
            store[key + "/in0"], store[key + "/in1"] = x
        else:
            store[key + "/in"] = x
        for i, a in enumerate(parts_a):
            store[f"{key}/out{i}"] = a
    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **store)
    print(f"wrote {path}: {len(store)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
