"""The CUDA kernel templates run on the CPU emulation of tests/emu (one OS thread per CUDA thread,
real barriers, per-CTA shared memory) and checked against the oracle.

This is a test of the kernels' LOGIC - index maps of every engine shape, prologue / epilogue
modes, barrier placement, the ticket protocol of the fused 2D kernel - in the suite that runs
where no GPU exists.  It is not a product path: libmufft_emu.so lives under tests/emu/_build, is
loaded by this file only, and the twiddle tables it is fed come from the real library's host code.
Kernels already verified on the GPU double as the check of the emulation itself.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

import cpu_libs

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
IN_COMPLEX, IN_HALF, IN_C2R, IN_MUL, IN_MUL_C2R = 0, 1, 2, 3, 4
OUT_COMPLEX, OUT_R2C, OUT_R2C_FULL = 0, 1, 2
vp, sz = C.c_void_p, C.c_size_t


@pytest.fixture(scope="module")
def emu():
    """The kernels as the product builds them; MUFFT_EMU_EXTRA="-D..." runs the same tests on an A/B variant of the kernels
    (built into tests/emu/_build_variant) before GPU time is spent on it."""
    extra = os.environ.get("MUFFT_EMU_EXTRA", "")
    build = "_build_variant" if extra else "_build"
    subprocess.run(["make", "-s", "-C", EMU_DIR, "-j", str(min(8, os.cpu_count() or 1)), f"BUILD={build}", f"EXTRA={extra}"], check=True)
    return C.CDLL(os.path.join(EMU_DIR, build, "libmufft_emu.so"), mode=os.RTLD_LOCAL)


@pytest.fixture(scope="module", params=["table", "derived"])
def emu_rows(request, emu):
    """The row kernels twice: twiddles from the table (product default) and with -DMUFFT_TW_ALL=2, every
    kernel deriving all but the power-of-two twiddles (an A/B build of tools/ab_build.sh)."""
    if request.param == "table":
        return emu
    subprocess.run(["make", "-s", "-C", EMU_DIR, "-j", str(min(8, os.cpu_count() or 1)), "BUILD=_build_derived", "EXTRA=-DMUFFT_TW_ALL=2"], check=True)
    return C.CDLL(os.path.join(EMU_DIR, "_build_derived", "libmufft_emu.so"), mode=os.RTLD_LOCAL)


@pytest.fixture(scope="module")
def host():
    """Host-side helpers of the real library (pure CPU code): engine twiddle tables and the variant choice."""
    import mufft_b200
    lib = mufft_b200.lib
    lib.mcu_engine_table_size.restype, lib.mcu_engine_table_size.argtypes = sz, [C.c_uint, C.c_int]
    lib.mcu_engine_table_fill.restype, lib.mcu_engine_table_fill.argtypes = None, [vp, C.c_uint, C.c_int, C.c_int]
    lib.mcu_rows_variant.restype, lib.mcu_rows_variant.argtypes = C.c_int, [C.c_uint, C.c_int, C.c_int]
    return lib


def engine_table(host, n, direction, variant=0):
    entries = host.mcu_engine_table_size(n, variant)
    assert entries > 0, f"no engine shape for n={n} variant={variant}"
    table = np.zeros(max(entries, 1), np.complex64)
    host.mcu_engine_table_fill(table.ctypes.data, n, direction, variant)
    return table


def resolve_twiddles(direction, n):
    """(0 + i*dir) * exp(i*pi*dir*k/n), k < n, in double then rounded (mufft_host.c: make_resolve_twiddles, ref fft.c:441-455)."""
    k = np.arange(n, dtype=np.float64)
    c, s = np.cos(np.pi * direction * k / n).astype(np.float32), np.sin(np.pi * direction * k / n).astype(np.float32)
    return (-np.float32(direction) * s + 1j * np.float32(direction) * c).astype(np.complex64)


def rows(emu, host, n, direction, in_mode, out_mode, x, out_elems, in2=None, norm=1.0, rtw=None):
    """Runs the emulated row kernel of complex length n on the rows of x; returns [batch][out_elems] complex."""
    fn = getattr(emu, f"emu_rows_{n}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, sz, sz, sz, sz, C.c_float]
    variant = host.mcu_rows_variant(n, in_mode, out_mode)
    tw = engine_table(host, n, direction, variant)
    x = np.ascontiguousarray(x, np.complex64)
    batch = x.shape[0]
    out = np.full((batch, out_elems), np.nan + 1j * np.nan, np.complex64)
    rc = fn(direction, in_mode, out_mode, out.ctypes.data, x.ctypes.data, None if in2 is None else in2.ctypes.data,
            tw.ctypes.data, None if rtw is None else rtw.ctypes.data, out_elems, x.shape[1], 0, batch, norm)
    assert rc == 0, f"emu_rows_{n} returned {rc}"
    return out


def rnd(rng, *shape):
    return rng.random(shape, np.float32) - np.float32(0.5)


def crnd(rng, *shape):
    return (rnd(rng, *shape) + 1j * rnd(rng, *shape)).astype(np.complex64)


def tol(n):
    return 1e-6 * math.log2(n)


# 64: several transforms per CTA in sub-warp groups; 512: one warp per transform + mirrored shapes;
# 2048 / 4096: CTA-wide exchanges, mirrored shapes; 16384: the 512-thread one-CTA-per-SM shape
# (MCU_ENGINE_BIG_N) that has not run on a GPU yet.
ROWS_SIZES = [64, 512, 2048, 4096, 16384]
ALL_ROWS_SIZES = [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]  # every engine shape


@pytest.mark.parametrize("n", ROWS_SIZES)
@pytest.mark.parametrize("direction", [-1, 1])
def test_rows_c2c(emu_rows, host, oracle, n, direction):
    emu = emu_rows
    rng = np.random.default_rng(n + direction)
    batch = 3 if n <= 4096 else 2  # 3 leaves a partly filled CTA at n = 64
    x = crnd(rng, batch, n)
    got = rows(emu, host, n, direction, IN_COMPLEX, OUT_COMPLEX, x, n)
    for b in range(batch):
        assert cpu_libs.rel_l2(got[b], oracle.c2c(x[b], direction)) <= tol(n)


@pytest.mark.parametrize("n", ROWS_SIZES)
def test_rows_zero_padded(emu_rows, host, oracle, n):
    emu = emu_rows
    rng = np.random.default_rng(n)
    x = crnd(rng, 2, n // 2)
    got = rows(emu, host, n, -1, IN_HALF, OUT_COMPLEX, x, n)
    for b in range(2):
        assert cpu_libs.rel_l2(got[b], oracle.c2c(x[b], -1, cpu_libs.FLAG_ZERO_PAD, n=n)) <= tol(n)


@pytest.mark.parametrize("n", ROWS_SIZES)
@pytest.mark.parametrize("full", [False, True])
def test_rows_r2c(emu_rows, host, oracle, n, full):
    """Real transform of 2n points on the complex engine of length n (mirrored last pass where the shape has one)."""
    emu = emu_rows
    rng = np.random.default_rng(2 * n + full)
    r = rnd(rng, 2, 2 * n)
    elems = 2 * n if full else n + 1
    got = rows(emu, host, n, -1, IN_COMPLEX, OUT_R2C_FULL if full else OUT_R2C, r.view(np.complex64), elems, rtw=resolve_twiddles(-1, n))
    for b in range(2):
        want = oracle.r2c(r[b], cpu_libs.FLAG_FULL_R2C if full else 0)
        assert cpu_libs.rel_l2(got[b], want) <= tol(2 * n)


@pytest.mark.parametrize("n", [512, 1024, 16384])
def test_rows_spectral_multiply_c2r(emu_rows, host, oracle, n):
    """Load stage of the mono mufft_execute_conv_output: product of two half spectra, C2R resolve, inverse
    transform (MCU_IN_MUL_C2R; 512 and 16384 feed the mirrored first pass, 1024 the general shape)."""
    emu = emu_rows
    rng = np.random.default_rng(9 * n)
    a = np.stack([np.fft.rfft(rnd(rng, 2 * n).astype(np.float64)) for _ in range(2)]).astype(np.complex64)  # [2][n+1]
    f = np.fft.rfft(rnd(rng, 2 * n).astype(np.float64)).astype(np.complex64)
    got = rows(emu, host, n, 1, IN_MUL_C2R, OUT_COMPLEX, a, n, in2=f, norm=1.0 / (2 * n), rtw=resolve_twiddles(1, n))
    for b in range(2):
        want = np.fft.irfft(a[b].astype(np.complex128) * f.astype(np.complex128), 2 * n)
        assert cpu_libs.rel_l2(got[b].view(np.float32), want) <= 2 * tol(2 * n)


@pytest.mark.parametrize("n", [512, 16384])
@pytest.mark.parametrize("full", [False, True])
def test_rows_zero_padded_r2c(emu_rows, host, oracle, n, full):
    """Zero-padded load stage and R2C epilogue in the same kernel (2n-point real transform of n reals)."""
    emu = emu_rows
    rng = np.random.default_rng(7 * n + full)
    r = rnd(rng, 2, n)
    elems = 2 * n if full else n + 1
    got = rows(emu, host, n, -1, IN_HALF, OUT_R2C_FULL if full else OUT_R2C, r.view(np.complex64), elems, rtw=resolve_twiddles(-1, n))
    flags = cpu_libs.FLAG_ZERO_PAD | (cpu_libs.FLAG_FULL_R2C if full else 0)
    for b in range(2):
        assert cpu_libs.rel_l2(got[b], oracle.r2c(r[b], flags, n=2 * n)) <= tol(2 * n)


@pytest.mark.parametrize("n", [512, 16384])
@pytest.mark.parametrize("direction", [-1, 1])
def test_rows_scatter_to_peer_slabs(emu, host, n, direction):
    """MCU_OUT_PEERS: the row kernel's stores are the exchange step of a slab 2D transform - element x
    of row b lands in the column slab of device x / (n / peers), at row row0 + b."""
    peers, batch, row0, rows_total = 4, 3, 2, 6
    width = n // peers
    rng = np.random.default_rng(n + 99)
    x = crnd(rng, batch, n)
    slabs = [np.full((rows_total, width), np.nan, np.complex64) for _ in range(peers)]
    bases = (C.c_void_p * peers)(*[s_.ctypes.data for s_ in slabs])
    fn = getattr(emu, f"emu_rows_peers_{n}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_uint, C.c_uint, sz, sz, vp, vp, sz, sz]
    tw = engine_table(host, n, direction)
    assert fn(direction, bases, peers, int(math.log2(width)), row0, width, x.ctypes.data, tw.ctypes.data, n, batch) == 0
    want = np.fft.fft(x.astype(np.complex128), axis=1) if direction < 0 else np.fft.ifft(x.astype(np.complex128), axis=1, norm="forward")
    for p_ in range(peers):
        assert np.all(np.isnan(slabs[p_][:row0].real)) and np.all(np.isnan(slabs[p_][row0 + batch:].real)), "wrote outside its rows"
        for b in range(batch):
            assert cpu_libs.rel_l2(slabs[p_][row0 + b], want[b, p_ * width:(p_ + 1) * width]) <= 2 * tol(n)


@pytest.mark.parametrize("n", [64, 512, 1024, 16384])
def test_rows_r2c_scatter_to_peer_slabs(emu, host, oracle, n):
    """MCU_OUT_R2C_PEERS: the exchange step of a slab R2C transform in the epilogue of the real row transform -
    bin k of row b lands in the slab of device min(k / (n / peers), peers - 1); the last slab is one column
    wider and takes the Nyquist bin (512 / 16384: mirrored last pass, 64 / 1024: general shape)."""
    peers, batch, row0, rows_total = 4, 3, 1, 5
    width = n // peers
    rng = np.random.default_rng(n + 123)
    r = rnd(rng, batch, 2 * n)
    slabs = [np.full((rows_total, width + (1 if p_ == peers - 1 else 0)), np.nan + 1j * np.nan, np.complex64) for p_ in range(peers)]
    bases = (C.c_void_p * peers)(*[s_.ctypes.data for s_ in slabs])
    fn = getattr(emu, f"emu_rows_r2c_peers_{n}")
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(C.c_void_p), C.c_uint, C.c_uint, sz, sz, vp, vp, vp, sz, sz]
    variant = host.mcu_rows_variant(n, IN_COMPLEX, 4)
    assert variant == host.mcu_rows_variant(n, IN_COMPLEX, OUT_R2C)
    tw = engine_table(host, n, -1, variant)
    assert fn(bases, peers, int(math.log2(width)), row0, width, r.ctypes.data, tw.ctypes.data, resolve_twiddles(-1, n).ctypes.data, n, batch) == 0
    for b in range(batch):
        want = oracle.r2c(r[b])
        got = np.concatenate([s_[row0 + b] for s_ in slabs])
        assert got.shape == want.shape
        assert cpu_libs.rel_l2(got, want) <= max(tol(2 * n), 2e-7)
    for s_ in slabs:
        assert np.all(np.isnan(s_[:row0].real)) and np.all(np.isnan(s_[row0 + batch:].real)), "wrote outside its rows"


@pytest.mark.parametrize("n", [512, 16384])
@pytest.mark.parametrize("mono", [False, True])
def test_rows_partitioned_multiply(emu_rows, host, oracle, n, mono):
    """MCU_IN_MULACC / MCU_IN_MULACC_C2R: output block b = IFFT(sum_p S[b - p] .* F[p]) / N, the load stage of
    mufft_execute_conv_output_partitioned_batch (stereo: n-point complex; mono: C2R resolve of n + 1 bins)."""
    emu = emu_rows
    parts, batch = 3, 2
    rng = np.random.default_rng(n + 5 + mono)
    bins = n + 1 if mono else n
    if mono:  # spectra of real signals, so that the C2R result is the real convolution
        spectra = np.stack([np.fft.rfft(rnd(rng, 2 * n).astype(np.float64)) for _ in range(batch + parts - 1)]).astype(np.complex64)
        filt = np.stack([np.fft.rfft(rnd(rng, 2 * n).astype(np.float64)) for _ in range(parts)]).astype(np.complex64)
    else:
        spectra, filt = crnd(rng, batch + parts - 1, bins), crnd(rng, parts, bins)
    fn = getattr(emu, f"emu_rows_parts_{n}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, vp, vp, vp, vp, vp, sz, sz, sz, C.c_float, C.c_uint, C.c_longlong, C.c_longlong]
    mode = 6 if mono else 5
    variant = host.mcu_rows_variant(n, mode, OUT_COMPLEX)
    tw = engine_table(host, n, 1, variant)
    rtw = resolve_twiddles(1, n)
    out = np.full((batch, n), np.nan, np.complex64)
    norm = 1.0 / (2 * n if mono else n)
    first = spectra[parts - 1:]  # S[0]; S[-1], S[-2] sit in front of it
    assert fn(mode, out.ctypes.data, first.ctypes.data, filt.ctypes.data, tw.ctypes.data, rtw.ctypes.data, n, bins, batch, norm,
              parts, -bins, bins) == 0
    for b in range(batch):
        acc = sum(spectra[parts - 1 + b - p_].astype(np.complex128) * filt[p_].astype(np.complex128) for p_ in range(parts))
        want = np.fft.irfft(acc, 2 * n) if mono else np.fft.ifft(acc)
        got = out[b].view(np.float32) if mono else out[b]
        assert cpu_libs.rel_l2(got, want) <= 2 * tol(n)


@pytest.mark.parametrize("n", ROWS_SIZES)
def test_rows_c2r(emu_rows, host, oracle, n):
    emu = emu_rows
    rng = np.random.default_rng(3 * n)
    spec = np.stack([oracle.r2c(rnd(rng, 2 * n)) for _ in range(2)])  # [2][n+1]
    got = rows(emu, host, n, 1, IN_C2R, OUT_COMPLEX, spec, n, rtw=resolve_twiddles(1, n))
    for b in range(2):
        assert cpu_libs.rel_l2(got[b].view(np.float32), oracle.c2r(spec[b], 2 * n)) <= tol(2 * n)


@pytest.mark.parametrize("n", [512, 16384])
def test_rows_spectral_multiply(emu_rows, host, n):
    """Load stage of mufft_execute_conv_output: norm * a .* b, then the inverse transform (ref kernel.c:22-34)."""
    emu = emu_rows
    rng = np.random.default_rng(5 * n)
    a, f = crnd(rng, 2, n), crnd(rng, n)
    got = rows(emu, host, n, 1, IN_MUL, OUT_COMPLEX, a, n, in2=f, norm=1.0 / n)
    want = np.fft.ifft(a.astype(np.complex128) * f.astype(np.complex128), axis=1)
    for b in range(2):
        assert cpu_libs.rel_l2(got[b], want[b]) <= tol(n)


class PeerDest(C.Structure):  # abi.h: mcu_peer_dest
    _fields_ = [("base", vp * 16), ("row0", sz), ("pitch", sz), ("shift", C.c_uint), ("last", C.c_uint), ("pitch_last", sz)]


class Args2D(C.Structure):  # abi.h: mcu_2d_args
    _fields_ = [("out", vp), ("in_", vp), ("tw_x", vp), ("tw_y", vp), ("rtw", vp), ("kind", C.c_int), ("counters", vp), ("mid", vp),
                ("ring", C.c_uint), ("out_image", sz), ("in_image", sz), ("batch", sz), ("nx", C.c_uint), ("ny", C.c_uint),
                ("direction", C.c_int), ("lag", C.c_uint)]


class ArgsLong(C.Structure):  # abi.h: mcu_long_args
    _fields_ = [("out", vp), ("in_", vp), ("tw", vp), ("roots", vp), ("scratch", vp), ("scratch_slots", sz), ("out_batch", sz),
                ("in_batch", sz), ("batch", sz), ("n", C.c_uint), ("n1", C.c_uint), ("n2", C.c_uint), ("direction", C.c_int),
                ("in_mode", C.c_int), ("to_peers", C.c_int), ("peers", PeerDest)]


KIND_C2C, KIND_R2C, KIND_R2C_FULL, KIND_C2R = 0, 1, 2, 3


def fused_2d(emu, nx, ny, ctas, **kw):
    fn = getattr(emu, f"emu_launch_2d_{nx}x{ny}")
    fn.restype, fn.argtypes = C.c_int, [C.POINTER(Args2D), C.c_uint]
    a = Args2D(nx=nx, ny=ny, **kw)
    return fn(C.byref(a), ctas)


@pytest.mark.parametrize("nx,ny,batch,ctas,lag", [
    (256, 256, 3, 1, 1),    # one CTA walks every ticket alone
    (256, 256, 4, 3, 1),    # CTAs side by side, column tiles right behind their rows (the deferred release gets exercised)
    (256, 256, 2, 5, 2),    # lag >= batch: all rows first
    (512, 256, 3, 2, 1),    # rectangular: different row / column tile shapes, two twiddle tables
    (2048, 256, 3, 2, 1),   # two warps per row: rows on named barriers
])
@pytest.mark.parametrize("direction", [-1, 1])
def test_fused_2d(emu, host, oracle, nx, ny, batch, ctas, lag, direction):
    """The persistent fused 2D kernel with CTAs really running side by side: producer (row) tickets, lagged
    consumer (column) tickets, the tiles-done counters and the deferred release (fft2d_kernel.cuh), in the
    flavour the product build uses for the shape."""
    rng = np.random.default_rng(11 + ctas)
    x = crnd(rng, batch, ny, nx)
    out = np.full((batch, ny, nx), np.nan, np.complex64)
    counters = np.zeros(1 + 2 * batch, np.uint32)
    twx, twy = engine_table(host, nx, direction), engine_table(host, ny, direction)
    assert fused_2d(emu, nx, ny, ctas, out=out.ctypes.data, in_=x.ctypes.data, tw_x=twx.ctypes.data, tw_y=twy.ctypes.data,
                    kind=KIND_C2C, counters=counters.ctypes.data, out_image=nx * ny, in_image=nx * ny, batch=batch,
                    direction=direction, lag=lag) == 0
    for b in range(batch):
        assert cpu_libs.rel_l2(out[b], oracle.c2c_2d(x[b], direction)) <= tol(nx * ny)
    assert list(counters[1:1 + batch]) == [getattr(emu, f"emu_2d_row_tiles_{nx}x{ny}")()] * batch  # producer (row) tiles per image


@pytest.mark.parametrize("nx,ny,ctas", [(256, 256, 1), (512, 256, 3), (2048, 256, 2)])
@pytest.mark.parametrize("full", [0, 1])
def test_fused_2d_real(emu, host, oracle, nx, ny, ctas, full):
    """Real images in the fused kernel: row tiles run the R2C row transform (engine of nx complex points + the
    resolve epilogue; 512 uses the mirrored-last-pass shape) into rows of pitch 2 nx bins, column
    tiles transform the nx + 1 (2 nx with FULL_R2C) columns that carry data, the last one partly masked."""
    batch = 3
    rng = np.random.default_rng(nx + ny + full)
    x = rnd(rng, batch, ny, 2 * nx)
    out = np.full((batch, ny, 2 * nx), np.nan + 1j * np.nan, np.complex64)
    counters = np.zeros(1 + 2 * batch, np.uint32)
    kind = KIND_R2C_FULL if full else KIND_R2C
    vfn = getattr(emu, f"emu_2d_rows_variant_{nx}x{ny}")
    vfn.restype, vfn.argtypes = C.c_int, [C.c_int]
    twx, twy, rtw = engine_table(host, nx, -1, vfn(kind)), engine_table(host, ny, -1), resolve_twiddles(-1, nx)
    assert fused_2d(emu, nx, ny, ctas, out=out.ctypes.data, in_=x.ctypes.data, tw_x=twx.ctypes.data, tw_y=twy.ctypes.data, rtw=rtw.ctypes.data,
                    kind=kind, counters=counters.ctypes.data, out_image=2 * nx * ny, in_image=nx * ny, batch=batch, direction=-1, lag=1) == 0
    width = 2 * nx if full else nx + 1
    for b in range(batch):
        want = np.fft.fft2(x[b].astype(np.float64))[:, :width]
        assert cpu_libs.rel_l2(out[b][:, :width], want) <= tol(2 * nx * ny)
    want = oracle.r2c_2d(x[0], cpu_libs.FLAG_FULL_R2C if full else 0)
    assert cpu_libs.rel_l2(out[0][:, :width], want[:, :width]) <= tol(2 * nx * ny)
    if not full:
        assert np.all(np.isnan(out[:, :, width:].real)), "columns beyond Nx/2 + 1 must stay untouched"


@pytest.mark.parametrize("nx,ny,batch,ctas,lag,ring", [
    (256, 256, 1, 2, 1, 1),   # one image: the fft.h call
    (256, 256, 5, 3, 1, 2),   # ring shorter than the batch: producers wait for the slot's previous consumers
    (512, 256, 4, 2, 2, 3),   # mirrored-first-pass rows; two images of lag
    (2048, 256, 3, 2, 1, 3),  # general row shape (the mirrored one would change the tile shape), ring = batch
])
def test_fused_2d_c2r(emu, host, oracle, nx, ny, batch, ctas, lag, ring):
    """Complex-to-real images in the fused kernel (ref fft.c:861-916: columns first): column tiles transform the
    nx + 1 columns of the input spectrum into a ring slot, row tiles run the C2R resolve + inverse row transform
    from the slot into the dense real output; the slot is only reused once its consumers are done."""
    rng = np.random.default_rng(nx + ny + batch)
    real = rnd(rng, batch, ny, 2 * nx)
    spec = np.zeros((batch, ny, 2 * nx), np.complex64)
    spec[:, :, :nx + 1] = np.fft.rfft2(real.astype(np.float64)).astype(np.complex64)
    spec[:, :, nx + 1:] = np.nan  # padding columns: never read
    out = np.full((batch, ny, 2 * nx), np.nan, np.float32)
    counters = np.zeros(1 + 2 * batch, np.uint32)
    mid = np.full((ring, ny, nx + 2), np.nan, np.complex64)
    vfn = getattr(emu, f"emu_2d_rows_variant_{nx}x{ny}")
    vfn.restype, vfn.argtypes = C.c_int, [C.c_int]
    twx, twy, rtw = engine_table(host, nx, 1, vfn(KIND_C2R)), engine_table(host, ny, 1), resolve_twiddles(1, nx)
    assert fused_2d(emu, nx, ny, ctas, out=out.ctypes.data, in_=spec.ctypes.data, tw_x=twx.ctypes.data, tw_y=twy.ctypes.data, rtw=rtw.ctypes.data,
                    kind=KIND_C2R, counters=counters.ctypes.data, mid=mid.ctypes.data, ring=ring, out_image=nx * ny, in_image=2 * nx * ny,
                    batch=batch, direction=1, lag=lag) == 0
    for b in range(batch):
        assert cpu_libs.rel_l2(out[b], real[b].astype(np.float64) * (2 * nx * ny)) <= tol(2 * nx * ny)
    clean = np.where(np.isnan(spec[0].real), 0, spec[0]).astype(np.complex64)
    assert cpu_libs.rel_l2(out[0], oracle.c2r_2d(clean)) <= tol(2 * nx * ny)


@pytest.mark.parametrize("n1,n2", [(16, 1024), (32, 1024), (32, 2048)])
@pytest.mark.parametrize("case", ["fwd", "inv", "zero_pad", "peers"])
def test_four_step_rows(emu, host, oracle, n1, n2, case):
    """long_kernel.cuh: n1 x n2 points per transform in one launch - register column transforms with derived
    inter-stage twiddles into the CTA's scratch slot, row transforms with the mapping switch of the last pass.
    Five transforms over two CTAs (strided in and out), so that slot reuse and the barriers between consecutive
    transforms are exercised."""
    fn = getattr(emu, f"emu_launch_long_{n1}x{n2}")
    fn.restype, fn.argtypes = C.c_int, [C.POINTER(ArgsLong), C.c_uint]
    n = n1 * n2
    direction = 1 if case == "inv" else -1
    batch, ctas = 5, 2
    rng = np.random.default_rng(n + len(case))
    tw = engine_table(host, n2, direction)
    j = np.arange(n, dtype=np.float64)
    roots = (np.cos(2 * np.pi * direction * j / n).astype(np.float32) + 1j * np.sin(2 * np.pi * direction * j / n).astype(np.float32)).astype(np.complex64)
    scratch = np.full((ctas, n), np.nan, np.complex64)
    in_len = n // 2 if case == "zero_pad" else n
    in_stride, out_stride = in_len + 8, n + 4
    xbuf = np.full((batch, in_stride), np.nan, np.complex64)
    xbuf[:, :in_len] = crnd(rng, batch, in_len)
    x = np.zeros((batch, n), np.complex128)
    x[:, :in_len] = xbuf[:, :in_len]
    want = np.fft.fft(x, axis=1) if direction < 0 else np.fft.ifft(x, axis=1, norm="forward")
    a = ArgsLong(in_=xbuf.ctypes.data, tw=tw.ctypes.data, roots=roots.ctypes.data, scratch=scratch.ctypes.data, scratch_slots=ctas,
                 in_batch=in_stride, batch=batch, n=n, n1=n1, n2=n2, direction=direction, in_mode=IN_HALF if case == "zero_pad" else IN_COMPLEX)
    if case == "peers":
        peers, row0, rows_total = 8, 1, 7
        width = n // peers
        slabs = [np.full((rows_total, width), np.nan, np.complex64) for _ in range(peers)]
        a.to_peers = 1
        for p_, s_ in enumerate(slabs):
            a.peers.base[p_] = s_.ctypes.data
        a.peers.row0, a.peers.pitch, a.peers.shift = row0, width, int(math.log2(width))
        assert fn(C.byref(a), ctas) == 0
        out = np.concatenate([s_[row0:row0 + batch] for s_ in slabs], axis=1)
        assert all(np.all(np.isnan(s_[:row0].real)) and np.all(np.isnan(s_[row0 + batch:].real)) for s_ in slabs)
    else:
        obuf = np.full((batch, out_stride), np.nan, np.complex64)
        a.out, a.out_batch = obuf.ctypes.data, out_stride
        assert fn(C.byref(a), ctas) == 0
        out = obuf[:, :n]
        assert np.all(np.isnan(obuf[:, n:].real)), "nothing is written between the transforms"
    for b in range(batch):
        assert cpu_libs.rel_l2(out[b], want[b]) <= tol(n)
    if case in ("fwd", "inv"):
        assert cpu_libs.rel_l2(out[0], oracle.c2c(xbuf[0, :n].copy(), direction)) <= tol(n)


@pytest.mark.parametrize("n1,n2", [(64, 256), (32, 1024), (128, 512)])
@pytest.mark.parametrize("case", ["fwd", "inv", "zero_pad"])
def test_two_kernel_long_transform(emu, host, oracle, n1, n2, case):
    """N = N1 * N2 beyond the single-kernel limit: column kernel on the [N1][N2] view, then rowst_kernel.cuh
    (inter-stage twiddle, N2-point rows, transposed store), chained through scratch as mufft_host.c does."""
    n = n1 * n2
    direction = 1 if case == "inv" else -1
    fn = getattr(emu, f"emu_two_kernel_{n1}x{n2}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, sz, C.c_int]
    rng = np.random.default_rng(n + len(case))
    batch, half = 2, case == "zero_pad"
    x = crnd(rng, batch, n // 2 if half else n)
    out = np.full((batch, n), np.nan, np.complex64)
    scratch = np.zeros((batch, n), np.complex64)
    j = np.arange(n, dtype=np.float64)
    roots = (np.cos(2 * np.pi * direction * j / n).astype(np.float32) + 1j * np.sin(2 * np.pi * direction * j / n).astype(np.float32)).astype(np.complex64)
    assert fn(direction, out.ctypes.data, scratch.ctypes.data, x.ctypes.data, engine_table(host, n1, direction).ctypes.data,
              engine_table(host, n2, direction).ctypes.data, roots.ctypes.data, batch, int(half)) == 0
    for b in range(batch):
        want = oracle.c2c(x[b], direction, cpu_libs.FLAG_ZERO_PAD if half else 0, n=n)
        assert cpu_libs.rel_l2(out[b], want) <= tol(n)


@pytest.mark.parametrize("n", [64, 1024, 4096])
@pytest.mark.parametrize("mono", [0, 1])
@pytest.mark.parametrize("half", [0, 1])
def test_fused_convolution_kernel(emu, host, oracle, n, mono, half):
    """conv_kernel.cuh: load, forward transform, multiply by the filter spectrum (stereo) or resolve /
    multiply / fold back (mono), inverse transform, store - one kernel per batch of blocks."""
    fn = getattr(emu, f"emu_conv_{n}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, sz, sz, sz, C.c_float]
    big = 2 * n if mono else n                         # the convolution's FFT length in fft.h terms
    method = (0 if mono else 1) | (4 if half else 0) | 8
    rng = np.random.default_rng(n + 2 * mono + half)
    batch = 3
    in_len = big // 2 if half else big                 # samples per block (complex for stereo, real for mono)
    taps = rnd(rng, big // 2)
    if mono:
        blocks = rnd(rng, batch, in_len)
        x_arg = blocks.view(np.complex64)
    else:
        blocks = crnd(rng, batch, in_len)
        x_arg = blocks
    _, _, fspec = oracle.conv(blocks[0], taps, big, method)  # the filter spectrum exactly as conv_input(SECOND) leaves it
    out = np.full((batch, n), np.nan + 1j * np.nan, np.complex64)
    assert fn(mono, half, out.ctypes.data, x_arg.ctypes.data, fspec.ctypes.data, engine_table(host, n, -1).ctypes.data,
              engine_table(host, n, 1).ctypes.data, resolve_twiddles(-1, n).ctypes.data, resolve_twiddles(1, n).ctypes.data,
              n, x_arg.shape[1], batch, 1.0 / big) == 0
    for b in range(batch):
        want, _, _ = oracle.conv(blocks[b], taps, big, method)
        got = out[b].view(np.float32) if mono else out[b]
        assert cpu_libs.rel_l2(got, want) <= 2 * tol(big)


@pytest.mark.parametrize("n", ALL_ROWS_SIZES)
def test_every_engine_shape(emu, host, oracle, n):
    """Forward and inverse C2C plus the real transforms of 2n points on every length of engine_shape(), in a
    batch that leaves the last CTA partly filled."""
    rng = np.random.default_rng(1000 + n)
    per_cta = max(1, 256 // max(1, n // 32))
    batch = min(5, per_cta + 1) if n <= 2048 else 2
    x = crnd(rng, batch, n)
    for direction in (-1, 1):
        got = rows(emu, host, n, direction, IN_COMPLEX, OUT_COMPLEX, x, n)
        for b in range(batch):
            assert cpu_libs.rel_l2(got[b], oracle.c2c(x[b], direction)) <= max(tol(n), 2e-7)
    r = rnd(rng, batch, 2 * n)
    spec = rows(emu, host, n, -1, IN_COMPLEX, OUT_R2C, r.view(np.complex64), n + 1, rtw=resolve_twiddles(-1, n))
    back = rows(emu, host, n, 1, IN_C2R, OUT_COMPLEX, spec, n, rtw=resolve_twiddles(1, n))
    for b in range(batch):
        assert cpu_libs.rel_l2(spec[b], oracle.r2c(r[b])) <= max(tol(2 * n), 2e-7)
        assert cpu_libs.rel_l2(back[b].view(np.float32), oracle.c2r(spec[b], 2 * n)) <= max(tol(2 * n), 2e-7)


@pytest.mark.parametrize("n", [2, 8, 32, 128, 512, 2048])
@pytest.mark.parametrize("direction", [-1, 1])
def test_column_kernel(emu, host, oracle, n, direction):
    """cols_kernel.cuh: transforms down the columns of [n][width] images without a transpose; a width that is
    not a multiple of the tile, different input / output pitches, two images, and the zero-padded lower half
    (rows_valid) the two-kernel path uses."""
    fn = getattr(emu, f"emu_cols_{n}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, vp, vp, vp, sz, sz, sz, sz, sz, C.c_uint, C.c_uint]
    rng = np.random.default_rng(n + direction)
    width, in_pitch, out_pitch, images = 21, 24, 29, 2
    x = crnd(rng, images, n, in_pitch)
    tw = engine_table(host, n, direction)
    for rows_valid in (0, n // 2):
        out = np.full((images, n, out_pitch), np.nan + 1j * np.nan, np.complex64)
        assert fn(direction, out.ctypes.data, x.ctypes.data, tw.ctypes.data, out_pitch, in_pitch, n * out_pitch, n * in_pitch, images, width, rows_valid) == 0
        xin = x[:, :, :width].astype(np.complex128)
        if rows_valid:
            xin[:, rows_valid:, :] = 0
        want = np.fft.fft(xin, axis=1) if direction < 0 else np.fft.ifft(xin, axis=1, norm="forward")
        assert cpu_libs.rel_l2(out[:, :, :width], want) <= max(tol(n), 2e-7)
        assert np.all(np.isnan(out[:, :, width:].real)), "wrote beyond the requested columns"


class ArgsLongCols(C.Structure):  # abi.h: mcu_longcols_args
    _fields_ = [("out", vp), ("in_", vp), ("tw", vp), ("roots", vp), ("scratch", vp), ("counters", vp), ("ring", C.c_uint), ("lag", C.c_uint),
                ("out_pitch", sz), ("in_pitch", sz), ("width", C.c_uint), ("n", C.c_uint), ("n1", C.c_uint), ("direction", C.c_int)]


@pytest.mark.parametrize("n1,width,ctas,lag,ring,inplace", [
    (4, 32, 1, 1, 1, False),     # one block, one CTA
    (4, 100, 3, 1, 2, True),     # four blocks (the last one partly masked), ring shorter than the blocks, in place
    (8, 64, 2, 2, 3, False),
    (32, 40, 3, 1, 2, True),     # the BASELINE configs[4] column length
])
@pytest.mark.parametrize("direction", [-1, 1])
def test_four_step_columns(emu, host, oracle, n1, width, ctas, lag, ring, inplace, direction):
    """longcols_kernel.cuh: n1 x 1024-point transforms down the columns of a [n][width] matrix in one persistent launch -
    A tiles (register transforms + derived twiddles) into a ring of scratch slots, lagged B tiles (engine column
    transforms) out of it; slot reuse waits for the consumers."""
    if n1 == 32 and direction == 1:
        pytest.skip("one direction is enough at the large shape")
    fn = emu.emu_launch_longcols
    fn.restype, fn.argtypes = C.c_int, [C.POINTER(ArgsLongCols), C.c_uint]
    n = n1 * 1024
    pitch = width + 3
    rng = np.random.default_rng(n1 + width)
    buf = np.full((n, pitch), np.nan, np.complex64)
    buf[:, :width] = crnd(rng, n, width)
    x = buf[:, :width].astype(np.complex128)
    want = np.fft.fft(x, axis=0) if direction < 0 else np.fft.ifft(x, axis=0, norm="forward")
    out = buf if inplace else np.full((n, pitch), np.nan, np.complex64)
    first_col = buf[:, 0].copy()
    tw = engine_table(host, 1024, direction)
    j = np.arange(n, dtype=np.float64)
    roots = (np.cos(2 * np.pi * direction * j / n).astype(np.float32) + 1j * np.sin(2 * np.pi * direction * j / n).astype(np.float32)).astype(np.complex64)
    blocks = (width + 31) // 32
    scratch = np.full((ring, n * 32), np.nan, np.complex64)
    counters = np.zeros(1 + 2 * blocks, np.uint32)
    a = ArgsLongCols(out=out.ctypes.data, in_=buf.ctypes.data, tw=tw.ctypes.data, roots=roots.ctypes.data, scratch=scratch.ctypes.data,
                     counters=counters.ctypes.data, ring=ring, lag=lag, out_pitch=pitch, in_pitch=pitch, width=width, n=n, n1=n1, direction=direction)
    assert fn(C.byref(a), ctas) == 0
    assert cpu_libs.rel_l2(out[:, :width], want) <= tol(n)
    assert np.all(np.isnan(out[:, width:].real)), "nothing is written beyond the last column"
    assert cpu_libs.rel_l2(out[:, 0], oracle.c2c(first_col, direction)) <= tol(n)


# ---- the kernels that move data with the TMA unit (tests/emu/emu_tma_inst.cpp; tma_util.cuh, MUFFT_EMU branch) ----------
class ArgsRows(C.Structure):  # abi.h: mcu_rows_args
    _fields_ = [("out", vp), ("in_", vp), ("in2", vp), ("tw", vp), ("rtw", vp), ("out_stride", sz), ("in_stride", sz), ("in2_stride", sz),
                ("batch", sz), ("n", C.c_uint), ("direction", C.c_int), ("in_mode", C.c_int), ("out_mode", C.c_int), ("norm", C.c_float),
                ("parts", C.c_uint), ("in_part_stride", C.c_longlong), ("in2_part_stride", C.c_longlong), ("peers", PeerDest)]


def prefetch_rows(emu, host, n, direction, in_mode, out_mode, xbuf, in_stride, out_elems, batch, ctas, in2=None, norm=1.0, rtw=None):
    """rows_prefetch_kernel.cuh on `ctas` persistent CTAs; xbuf holds the rows `in_stride` complex elements apart."""
    fn = getattr(emu, f"emu_launch_rows_prefetch_{n}")
    fn.restype, fn.argtypes = C.c_int, [C.POINTER(ArgsRows), C.c_uint]
    tw = engine_table(host, n, direction, host.mcu_rows_variant(n, in_mode, out_mode))
    out = np.full((batch, out_elems + 3), np.nan + 1j * np.nan, np.complex64)
    a = ArgsRows(out=out.ctypes.data, in_=xbuf.ctypes.data, in2=None if in2 is None else in2.ctypes.data, tw=tw.ctypes.data,
                 rtw=None if rtw is None else rtw.ctypes.data, out_stride=out_elems + 3, in_stride=in_stride, batch=batch, n=n,
                 direction=direction, in_mode=in_mode, out_mode=out_mode, norm=norm)
    rc = fn(C.byref(a), ctas)
    assert rc == 0, f"emu_launch_rows_prefetch_{n} returned {rc}"
    assert np.all(np.isnan(out[:, out_elems:].real)), "nothing is written between the transforms"
    return out[:, :out_elems]


@pytest.mark.parametrize("n", [2048, 8192])
def test_tma_prefetch_rows(emu, host, oracle, n):
    """The persistent row kernel that lets the "TMA unit" fetch the next transform into the exchange area: five transforms
    on two CTAs (so every CTA prefetches, waits on its mbarrier and reuses its buffer), every load mode it takes."""
    rng = np.random.default_rng(n)
    batch, ctas = 5, 2
    x = crnd(rng, batch, n)
    for direction in (-1, 1):
        got = prefetch_rows(emu, host, n, direction, IN_COMPLEX, OUT_COMPLEX, x, n, n, batch, ctas)
        for b in range(batch):
            assert cpu_libs.rel_l2(got[b], oracle.c2c(x[b], direction)) <= tol(n)
    half = crnd(rng, batch, n // 2)
    got = prefetch_rows(emu, host, n, -1, IN_HALF, OUT_COMPLEX, half, n // 2, n, batch, ctas)
    for b in range(batch):
        assert cpu_libs.rel_l2(got[b], oracle.c2c(half[b], -1, cpu_libs.FLAG_ZERO_PAD, n=n)) <= tol(n)
    f = crnd(rng, n)
    got = prefetch_rows(emu, host, n, 1, IN_MUL, OUT_COMPLEX, x, n, n, batch, ctas, in2=f, norm=1.0 / n)
    want = np.fft.ifft(x.astype(np.complex128) * f.astype(np.complex128), axis=1)
    for b in range(batch):
        assert cpu_libs.rel_l2(got[b], want[b]) <= tol(n)
    # C2R: spectrum rows of n + 1 bins, tightly packed (every other row 8 bytes off a 16-byte boundary) and at an even pitch
    spec = np.stack([oracle.r2c(rnd(rng, 2 * n)) for _ in range(batch)])  # [batch][n + 1]
    for pitch in (n + 1, n + 2):
        sbuf = np.full((batch, pitch), np.nan, np.complex64)
        sbuf[:, :n + 1] = spec
        assert sbuf.ctypes.data % 16 == 0
        got = prefetch_rows(emu, host, n, 1, IN_C2R, OUT_COMPLEX, sbuf, pitch, n, batch, ctas, rtw=resolve_twiddles(1, n))
        for b in range(batch):
            assert cpu_libs.rel_l2(got[b].view(np.float32), oracle.c2r(spec[b], 2 * n)) <= tol(2 * n)
    if n == 8192:  # R2C with the resolve in the registers of the mirrored last pass
        r = rnd(rng, batch, 2 * n)
        for full in (0, 1):
            elems = 2 * n if full else n + 1
            got = prefetch_rows(emu, host, n, -1, IN_COMPLEX, OUT_R2C_FULL if full else OUT_R2C, r.view(np.complex64), n, elems, batch, ctas,
                                rtw=resolve_twiddles(-1, n))
            for b in range(batch):
                assert cpu_libs.rel_l2(got[b], oracle.r2c(r[b], cpu_libs.FLAG_FULL_R2C if full else 0)) <= tol(2 * n)


@pytest.mark.parametrize("nx,ny,batch,ctas,lag", [
    (512, 512, 4, 2, 1),    # six buffers, column tiles right behind their rows
    (512, 512, 3, 3, 2),
    (512, 1024, 3, 2, 1),   # three buffers (the 1024-point column tile sets the size), two twiddle tables
])
@pytest.mark.parametrize("direction", [-1, 1])
def test_tma_fused_2d(emu, host, oracle, nx, ny, batch, ctas, lag, direction):
    """fft2d_tma_kernel.cuh with its loader, two compute groups and storers as OS threads: mbarrier hand-overs (one `full`
    barrier per buffer and group), tickets, the rows-done counter, EXIT entries; bulk copies and tensor boxes emulated."""
    fn = getattr(emu, f"emu_launch_2d_tma_{nx}x{ny}")
    fn.restype, fn.argtypes = C.c_int, [C.POINTER(Args2D), C.c_uint]
    rng = np.random.default_rng(nx + ny + ctas)
    x = crnd(rng, batch, ny, nx)
    out = np.full((batch, ny, nx), np.nan, np.complex64)
    counters = np.zeros(1 + 2 * batch, np.uint32)
    twx, twy = engine_table(host, nx, direction), engine_table(host, ny, direction)
    a = Args2D(nx=nx, ny=ny, out=out.ctypes.data, in_=x.ctypes.data, tw_x=twx.ctypes.data, tw_y=twy.ctypes.data, kind=KIND_C2C,
               counters=counters.ctypes.data, out_image=nx * ny, in_image=nx * ny, batch=batch, direction=direction, lag=lag)
    assert fn(C.byref(a), ctas) == 0
    for b in range(batch):
        assert cpu_libs.rel_l2(out[b], oracle.c2c_2d(x[b], direction)) <= tol(nx * ny)
    assert list(counters[1:1 + batch]) == [getattr(emu, f"emu_2d_tma_row_tiles_{nx}x{ny}")()] * batch  # row tiles announced per image


@pytest.mark.parametrize("ny,ctas", [(512, 2), (1024, 3)])
def test_tma_fused_2d_r2c(emu, host, oracle, ny, ctas):
    """Real images 1024 wide on the TMA kernel: rows resolved in registers on the way back into the buffer, bin 512 stored
    apart, 65 column tiles of which the last carries one column - stored from registers, the padding columns untouched."""
    nx, batch = 512, 3  # complex points per row
    fn = getattr(emu, f"emu_launch_2d_tma_{nx}x{ny}")
    fn.restype, fn.argtypes = C.c_int, [C.POINTER(Args2D), C.c_uint]
    vfn = getattr(emu, f"emu_2d_rows_variant_512x256")
    vfn.restype, vfn.argtypes = C.c_int, [C.c_int]
    rng = np.random.default_rng(ny)
    x = rnd(rng, batch, ny, 2 * nx)
    out = np.full((batch, ny, 2 * nx), 7.0 + 7.0j, np.complex64)
    counters = np.zeros(1 + 2 * batch, np.uint32)
    twx, twy, rtw = engine_table(host, nx, -1, vfn(KIND_R2C)), engine_table(host, ny, -1), resolve_twiddles(-1, nx)
    a = Args2D(nx=nx, ny=ny, out=out.ctypes.data, in_=x.ctypes.data, tw_x=twx.ctypes.data, tw_y=twy.ctypes.data, rtw=rtw.ctypes.data,
               kind=KIND_R2C, counters=counters.ctypes.data, out_image=2 * nx * ny, in_image=nx * ny, batch=batch, direction=-1, lag=1)
    assert fn(C.byref(a), ctas) == 0
    for b in range(batch):
        assert cpu_libs.rel_l2(out[b, :, :nx + 1], oracle.r2c_2d(x[b], 0)[:, :nx + 1]) <= tol(2 * nx * ny)
    assert np.all(out[:, :, nx + 1:] == 7.0 + 7.0j), "columns beyond nx + 1 are padding: never written"


@pytest.mark.parametrize("n1,n2", [(16, 1024), (32, 1024)])
@pytest.mark.parametrize("case", ["fwd", "inv", "zero_pad"])
def test_tma_four_step_rows(emu, host, n1, n2, case):
    """long_kernel.cuh with the inputs of step A arriving through the exchange area (TMA_A): sweeps of 256 columns asked for
    one ahead, the first sweep of the next transform when step B's last group has left the area."""
    fn = getattr(emu, f"emu_launch_long_tma_{n1}x{n2}")
    fn.restype, fn.argtypes = C.c_int, [C.POINTER(ArgsLong), C.c_uint]
    n = n1 * n2
    direction = 1 if case == "inv" else -1
    batch, ctas = 5, 2
    rng = np.random.default_rng(n + len(case))
    tw = engine_table(host, n2, direction)
    j = np.arange(n, dtype=np.float64)
    roots = (np.cos(2 * np.pi * direction * j / n).astype(np.float32) + 1j * np.sin(2 * np.pi * direction * j / n).astype(np.float32)).astype(np.complex64)
    scratch = np.full((ctas, n), np.nan, np.complex64)
    in_len = n // 2 if case == "zero_pad" else n
    in_stride, out_stride = in_len + 8, n + 4  # even strides: 16-byte aligned transforms
    xbuf = np.full((batch, in_stride), np.nan, np.complex64)
    xbuf[:, :in_len] = crnd(rng, batch, in_len)
    x = np.zeros((batch, n), np.complex128)
    x[:, :in_len] = xbuf[:, :in_len]
    want = np.fft.fft(x, axis=1) if direction < 0 else np.fft.ifft(x, axis=1, norm="forward")
    obuf = np.full((batch, out_stride), np.nan, np.complex64)
    a = ArgsLong(out=obuf.ctypes.data, in_=xbuf.ctypes.data, tw=tw.ctypes.data, roots=roots.ctypes.data, scratch=scratch.ctypes.data, scratch_slots=ctas,
                 out_batch=out_stride, in_batch=in_stride, batch=batch, n=n, n1=n1, n2=n2, direction=direction,
                 in_mode=IN_HALF if case == "zero_pad" else IN_COMPLEX)
    assert fn(C.byref(a), ctas) == 0
    assert np.all(np.isnan(obuf[:, n:].real)), "nothing is written between the transforms"
    for b in range(batch):
        assert cpu_libs.rel_l2(obuf[b, :n], want[b]) <= tol(n)
