"""The C host layer (mufft_b200/csrc/mufft_host.c, unchanged) on top of an emulated CUDA side
(tests/emu/emu_abi.cpp + the kernel templates on the CPU emulation): the whole fft.h /
mufft_cuda.h surface runs on the CPU and is checked against the oracle.

What this covers that the kernel-level emulation tests do not: plan construction and validation,
which path and which twiddle table a plan takes, chunking, the staged path for ordinary host
memory, the 2D and convolution orchestration, the path switches.  It is test infrastructure:
libmufft_hostemu.so lives under tests/emu/_build and nothing in the package or the bench loads it.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

import cpu_libs

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
vp, sz, u, i32 = C.c_void_p, C.c_size_t, C.c_uint, C.c_int


def tol(n):
    return max(1e-6 * math.log2(n), 2e-7)


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-s", "-C", EMU_DIR, "-j", str(min(8, os.cpu_count() or 1))], check=True)
    path = os.path.join(EMU_DIR, "_build", "libmufft_hostemu.so")
    wrapped = cpu_libs._CpuLib(path, "mufft_")  # the numpy-level wrapper the GPU parity tests use
    raw = wrapped.lib
    for name, res, args in [
        ("mufft_cuda_launch_count", C.c_ulonglong, [i32]), ("mufft_cuda_plan_2d_paths", i32, [vp]),
        ("mufft_cuda_last_error", C.c_char_p, []), ("mufft_alloc", vp, [sz]), ("mufft_free", None, [vp]),
        ("mufft_execute_plan_1d_batch", i32, [vp, vp, vp, sz, sz, sz]), ("mufft_execute_plan_2d_batch", i32, [vp, vp, vp, sz]),
        ("mufft_execute_conv_batch", i32, [vp, vp, vp, vp, sz, sz]), ("mufft_execute_conv_input_batch", i32, [vp, u, vp, vp, sz]),
        ("mufft_execute_conv_output_partitioned_batch", i32, [vp, vp, vp, vp, u, sz]),
        ("mufft_execute_plan_1d_batch_scatter_async", i32, [vp, C.POINTER(vp), u, sz, vp, sz, sz]), ("mufft_cuda_synchronize", i32, []),
        ("mufft_cuda_device_alloc", vp, [sz]), ("mufft_cuda_device_free", None, [vp]), ("mufft_cuda_plan_1d_uses_engine", i32, [vp]),
    ]:
        fn = getattr(raw, name)
        fn.restype, fn.argtypes = res, args
    return wrapped


def rnd(rng, *shape):
    return rng.random(shape, np.float32) - np.float32(0.5)


def crnd(rng, *shape):
    return (rnd(rng, *shape) + 1j * rnd(rng, *shape)).astype(np.complex64)


def managed(lib, arr):
    """A copy of arr in 'managed' memory of the emulated device (used in place by the host layer)."""
    arr = np.ascontiguousarray(arr)
    ptr = lib.lib.mufft_alloc(max(arr.nbytes, 8))
    view = np.ctypeslib.as_array((C.c_uint8 * arr.nbytes).from_address(ptr)).view(arr.dtype).reshape(arr.shape)
    view[...] = arr
    return ptr, view


@pytest.mark.parametrize("n,want", [(65536, (2, 0)), (1024, (0, 0))])
def test_managed_memory_prefetch_policy(lib, oracle, n, want):
    """Synchronous calls on mufft_alloc memory prefetch buffers of 256 KiB or more to the device in front of the
    kernel (input and output; measured 1.3x faster than fault-driven migration); small buffers are left to the page
    faults, and nothing is prefetched back (the next call may consume the output on the device).  The emulated
    CUDA side counts the calls."""
    raw = lib.lib
    raw.emu_prefetch_calls.restype, raw.emu_prefetch_calls.argtypes = u, [i32, i32]
    x = crnd(np.random.default_rng(n), n)
    px, vx = managed(lib, x)
    py, vy = managed(lib, np.zeros(n, np.complex64))
    plan = lib._create_plan_1d_c2c(n, -1, 0)
    try:
        raw.emu_prefetch_calls(0, 1), raw.emu_prefetch_calls(1, 1)
        lib._execute_plan_1d(plan, py, px)
        assert (raw.emu_prefetch_calls(1, 1), raw.emu_prefetch_calls(0, 1)) == want
        assert cpu_libs.rel_l2(vy, oracle.c2c(x, -1)) <= tol(n)
        # host memory is staged, never prefetched
        assert cpu_libs.rel_l2(lib.c2c(x, -1), oracle.c2c(x, -1)) <= tol(n)
        assert (raw.emu_prefetch_calls(1, 1), raw.emu_prefetch_calls(0, 1)) == (0, 0)
    finally:
        lib._free_plan_1d(plan)
        raw.mufft_free(px)
        raw.mufft_free(py)


def test_a_plan_belongs_to_the_device_it_was_created_on(lib, oracle):
    """One host thread per GPU: a plan's tables and scratch live on the device that was current at
    creation.  Executing it with another device current is refused with a message instead of
    faulting inside a kernel (the emulation has two devices for this)."""
    raw = lib.lib
    raw.mufft_cuda_set_device.restype, raw.mufft_cuda_set_device.argtypes = i32, [i32]
    n = 64
    x = cpu_libs.aligned_copy(crnd(np.random.default_rng(1), n), np.complex64)
    y = cpu_libs.aligned_zeros(n, np.complex64)
    plan = lib._create_plan_1d_c2c(n, -1, 0)
    plan2 = lib._create_plan_2d_c2c(16, 16, -1, 0)
    try:
        assert raw.mufft_cuda_set_device(1) == 0
        raw.mufft_cuda_launch_count(1)
        assert raw.mufft_execute_plan_1d_batch(plan, y.ctypes.data, x.ctypes.data, 1, 0, 0) != 0
        assert b"device 0 but device 1" in raw.mufft_cuda_last_error()
        img = cpu_libs.aligned_zeros(256, np.complex64)
        assert raw.mufft_execute_plan_2d_batch(plan2, img.ctypes.data, img.ctypes.data, 1) != 0
        assert raw.mufft_cuda_launch_count(1) == 0
    finally:
        assert raw.mufft_cuda_set_device(0) == 0
    assert raw.mufft_execute_plan_1d_batch(plan, y.ctypes.data, x.ctypes.data, 1, 0, 0) == 0
    assert cpu_libs.rel_l2(y, oracle.c2c(x, -1)) <= tol(n)
    lib._free_plan_1d(plan)
    lib._free_plan_2d(plan2)


@pytest.mark.parametrize("where", ["host", "managed"])
@pytest.mark.parametrize("n", [16, 4096])
def test_buffers_with_4_byte_alignment_are_staged(lib, oracle, where, n):
    """A float array need not be 8-byte aligned in C; the reference's scalar path takes it.  The
    kernels move 8-byte values, so the host layer must stage such a buffer (whatever memory it is in)
    instead of using it in place - and must not skip the transform."""
    rng = np.random.default_rng(n)
    r = rnd(rng, n)
    if where == "host":
        raw_in = np.zeros(n * 4 + 64, np.uint8)
        raw_out = np.zeros((n // 2 + 1) * 8 + 64, np.uint8)
        base_in, base_out = raw_in.ctypes.data, raw_out.ctypes.data
    else:
        base_in, raw_in = managed(lib, np.zeros(n * 4 + 64, np.uint8))
        base_out, raw_out = managed(lib, np.zeros((n // 2 + 1) * 8 + 64, np.uint8))
    off_in = (-base_in) % 8 + 4    # 4 mod 8
    off_out = (-base_out) % 8 + 4
    raw_in[off_in:off_in + n * 4] = r.view(np.uint8)
    plan = lib._create_plan_1d_r2c(n, 0)
    lib.lib.mufft_cuda_launch_count(1)
    lib._execute_plan_1d(plan, base_out + off_out, base_in + off_in)
    assert lib.lib.mufft_cuda_launch_count(1) == 1, lib.lib.mufft_cuda_last_error()
    got = raw_out[off_out:off_out + (n // 2 + 1) * 8].copy().view(np.complex64)
    assert cpu_libs.rel_l2(got, oracle.r2c(r)) <= tol(n)
    # and back: complex input aligned, real output misaligned
    plan_inv = lib._create_plan_1d_c2r(n, 0)
    spec = cpu_libs.aligned_copy(got, np.complex64)
    raw_in[:] = 0
    lib._execute_plan_1d(plan_inv, base_in + off_in, spec.ctypes.data)
    back = raw_in[off_in:off_in + n * 4].copy().view(np.float32)
    assert cpu_libs.rel_l2(back, oracle.c2r(spec, n)) <= tol(n)
    # the asynchronous calls cannot stage: they refuse
    if where == "managed":
        fn = lib.lib.mufft_execute_plan_1d_batch_async
        fn.restype, fn.argtypes = i32, [vp, vp, vp, sz, sz, sz]
        assert fn(plan, base_out + off_out, base_in + off_in, 1, 0, 0) != 0
        assert b"aligned" in lib.lib.mufft_cuda_last_error()
        lib.lib.mufft_free(base_in)
        lib.lib.mufft_free(base_out)
    lib._free_plan_1d(plan)
    lib._free_plan_1d(plan_inv)


@pytest.mark.parametrize("n", [2, 4, 8, 64, 512, 4096, 16384])
def test_fft_h_1d(lib, oracle, n):
    """mufft_execute_plan_1d on ordinary host memory (the staged path): C2C both directions, zero pad,
    R2C half / full, C2R - the reference's own call sequence (ref test.c:33-62)."""
    rng = np.random.default_rng(n)
    x = crnd(rng, n)
    for d in (-1, 1):
        assert cpu_libs.rel_l2(lib.c2c(x, d), oracle.c2c(x, d)) <= tol(n)
    if n >= 4:
        assert cpu_libs.rel_l2(lib.c2c(x[: n // 2], -1, cpu_libs.FLAG_ZERO_PAD, n=n), oracle.c2c(x[: n // 2], -1, cpu_libs.FLAG_ZERO_PAD, n=n)) <= tol(n)
        r = rnd(rng, n)
        for flags in (0, cpu_libs.FLAG_FULL_R2C):
            assert cpu_libs.rel_l2(lib.r2c(r, flags), oracle.r2c(r, flags)) <= tol(n)
        spec = oracle.r2c(r)
        assert cpu_libs.rel_l2(lib.c2r(spec, n), oracle.c2r(spec, n)) <= tol(n)


def test_constructor_validation_matches_the_reference(lib, oracle):
    """NULL for non powers of two, N == 1, real transforms of 2 points (ref fft.c:478-481, 517-520, 628-631)."""
    for kind in ("c2c", "r2c", "c2r"):
        for n in (0, 1, 2, 3, 6, 12, 100, 4, 64):
            want = oracle.plan_1d_is_null(kind, n) if n else True  # the reference returns a degenerate plan for 0; we refuse it
            assert lib.plan_1d_is_null(kind, n) == want, (kind, n)


SEPARATE_2D, TWO_KERNEL, GENERIC_PASSES, FOUR_STEP_16K = 1 << 20, 1 << 21, 1 << 22, 1 << 23  # mufft_cuda.h plan flags


@pytest.mark.parametrize("n,flags,launches,one_launch", [
    (16384, 0, 1, 1),                 # one CTA per transform on the engine
    (16384, FOUR_STEP_16K, 1, 1),     # 16 x 1024 on the four-step kernel
    (16384, TWO_KERNEL, 2, 0),        # column kernel + transposing row kernel through scratch
    (32768, 0, 1, 1),                 # 32 x 1024 four-step kernel, intermediate in the CTA's L2 slot
    (65536, 0, 1, 1),                 # 32 x 2048
    (65536, TWO_KERNEL, 2, 0),
    (131072, 0, 2, 0),                # beyond the four-step shapes: two kernels
    (32768, GENERIC_PASSES, 5, 0),    # one radix-8/4/2 pass per launch, as the reference (fft.c:811-850)
])
@pytest.mark.parametrize("zero_pad", [0, 1])
def test_long_transforms_pick_their_path(lib, n, flags, launches, one_launch, zero_pad):
    """Which path a plan takes is fixed by its creation flags (no process-global switch); launch counts tell which
    one ran; results against numpy float64, plain and with the zero-padded upper half."""
    raw = lib.lib
    rng = np.random.default_rng(n)
    cols = n // 2 if zero_pad else n
    x = crnd(rng, 2, cols)
    pin, _ = managed(lib, x)
    pout, yv = managed(lib, np.zeros((2, n), np.complex64))
    plan = lib._create_plan_1d_c2c(n, -1, flags | (cpu_libs.FLAG_ZERO_PAD if zero_pad else 0))
    assert plan
    raw.mufft_cuda_launch_count(1)
    assert raw.mufft_execute_plan_1d_batch(plan, pout, pin, 2, 0, 0) == 0, raw.mufft_cuda_last_error()
    assert raw.mufft_cuda_launch_count(1) == launches
    assert raw.mufft_cuda_plan_1d_uses_engine(plan) == one_launch
    lib._free_plan_1d(plan)
    xp = np.zeros((2, n), np.complex128)
    xp[:, :cols] = x
    assert cpu_libs.rel_l2(yv, np.fft.fft(xp, axis=1)) <= tol(n)
    raw.mufft_free(pin)
    raw.mufft_free(pout)


@pytest.mark.parametrize("n", [65536, 131072])
def test_long_real_transforms(lib, oracle, n):
    """R2C / C2R of n reals with n/2 complex points beyond the on-chip engine: resolve sweep + four-step kernel
    (32768 points) or two kernels (65536 points) through plan scratch."""
    rng = np.random.default_rng(n)
    r = rnd(rng, n)
    spec = lib.r2c(r)
    assert cpu_libs.rel_l2(spec, np.fft.rfft(r.astype(np.float64))) <= tol(n)
    back = lib.c2r(spec, n)
    assert cpu_libs.rel_l2(back, r.astype(np.float64) * n) <= tol(n)


@pytest.mark.parametrize("nx,ny", [(16, 8), (64, 64), (256, 256), (512, 256), (32, 2048), (4096, 4)])
def test_fft_h_2d(lib, oracle, nx, ny):
    """mufft_execute_plan_2d: C2C both directions (fused kernel for 256 x 256 and 512 x 256 here, row + column
    kernels otherwise), R2C half / full with the reference's pitch, C2R with its scratch-sized output."""
    rng = np.random.default_rng(nx + ny)
    x = crnd(rng, ny, nx)
    for d in (-1, 1):
        assert cpu_libs.rel_l2(lib.c2c_2d(x, d), oracle.c2c_2d(x, d)) <= tol(nx * ny)
    if nx >= 4:
        r = rnd(rng, ny, nx)
        for flags in (0, cpu_libs.FLAG_FULL_R2C):
            w = nx if flags else nx // 2 + 1
            assert cpu_libs.rel_l2(lib.r2c_2d(r, flags)[:, :w], oracle.r2c_2d(r, flags)[:, :w]) <= tol(nx * ny)
        spec = oracle.r2c_2d(r)
        assert cpu_libs.rel_l2(lib.c2r_2d(spec), oracle.c2r_2d(spec)) <= tol(nx * ny)


@pytest.mark.parametrize("kind", ["c2c", "r2c", "r2c_full", "c2r"])
@pytest.mark.parametrize("separate", [0, 1])
def test_fused_2d_is_the_default_and_a_plan_flag_turns_it_off(lib, oracle, kind, separate):
    """2D plans of 512 x 256 points (256-point row transforms for the real kinds): the whole batch is ONE launch of
    the fused kernel (C2R: only from 512 complex points per row on, through the plan's ring of intermediates) unless
    the plan was created with MUFFT_CUDA_FLAG_SEPARATE_2D: then a row launch and a column launch through plan scratch."""
    raw = lib.lib
    nx, ny, batch = 512, 256, 5
    flags = SEPARATE_2D if separate else 0
    rng = np.random.default_rng(5)
    if kind == "c2c":
        x = crnd(rng, batch, ny, nx)
        plan = lib._create_plan_2d_c2c(nx, ny, -1, flags)
        pout, yv = managed(lib, np.zeros((batch, ny, nx), np.complex64))
    elif kind == "c2r":
        real = rnd(rng, batch, ny, nx)
        x = np.zeros((batch, ny, nx), np.complex64)
        x[:, :, : nx // 2 + 1] = np.fft.rfft2(real.astype(np.float64)).astype(np.complex64)
        plan = lib._create_plan_2d_c2r(nx, ny, flags)
        pout, yv = managed(lib, np.zeros((batch, ny, nx), np.float32))
    else:
        x = rnd(rng, batch, ny, nx)
        plan = lib._create_plan_2d_r2c(nx, ny, flags | (cpu_libs.FLAG_FULL_R2C if kind == "r2c_full" else 0))
        pout, yv = managed(lib, np.zeros((batch, ny, nx), np.complex64))
    assert plan
    fused = not separate and kind != "c2r"  # C2R rows of 256 complex points are too short for the fused kernel to pay
    assert bool(raw.mufft_cuda_plan_2d_paths(plan) & 4) == fused
    pin, _ = managed(lib, x)
    raw.mufft_cuda_launch_count(1)
    assert raw.mufft_execute_plan_2d_batch(plan, pout, pin, batch) == 0, raw.mufft_cuda_last_error()
    assert raw.mufft_cuda_launch_count(1) == (1 if fused else 2)
    lib._free_plan_2d(plan)
    for b in range(batch):
        if kind == "c2c":
            assert cpu_libs.rel_l2(yv[b], oracle.c2c_2d(x[b], -1)) <= tol(nx * ny)
        elif kind == "c2r":
            assert cpu_libs.rel_l2(yv[b], oracle.c2r_2d(x[b])) <= tol(nx * ny)
        else:
            w = nx if kind == "r2c_full" else nx // 2 + 1
            want = oracle.r2c_2d(x[b], cpu_libs.FLAG_FULL_R2C if kind == "r2c_full" else 0)[:, :w]
            assert cpu_libs.rel_l2(yv[b][:, :w], want) <= tol(nx * ny)
    raw.mufft_free(pin)
    raw.mufft_free(pout)


def test_error_messages_survive_the_layers_above(lib):
    """An inner layer's message is what mufft_cuda_last_error() returns: a stereo partitioned convolution beyond the
    on-chip engine is refused by run_rows, two TRY() layers below the entry point (GPUTEST_r01 was red on this)."""
    raw = lib.lib
    plan = lib._create_plan_conv(32768, 0, 1 | 4 | 8)
    assert plan
    blk = 32768 * 8
    bufs = [raw.mufft_alloc(4 * blk) for _ in range(3)]
    assert raw.mufft_execute_conv_output_partitioned_batch(plan, bufs[0], bufs[1] + blk, bufs[2], 2, 1) != 0
    assert b"fits the on-chip engine" in raw.mufft_cuda_last_error()
    for b_ in bufs:
        raw.mufft_free(b_)
    lib._free_plan_conv(plan)


@pytest.mark.parametrize("n", [8, 256, 2048, 16384, 32768])
@pytest.mark.parametrize("method", [0, 1, 1 | 4 | 8, 0 | 4 | 8])
def test_fft_h_convolution(lib, oracle, n, method):
    """conv_input x 2 + conv_output as the reference's callers do (ref test.c:375-431), mono and stereo, with
    and without zero padding; block sizes as the reference reports them."""
    rng = np.random.default_rng(n + method)
    stereo = method & 1
    la = n // 2 if method & 4 else n
    lb = n // 2 if method & 8 else n
    a = crnd(rng, la) if stereo else rnd(rng, la)
    b = rnd(rng, lb)
    got, s0, s1 = lib.conv(a, b, n, method)
    want, w0, w1 = oracle.conv(a, b, n, method)
    assert s0.size == w0.size
    assert cpu_libs.rel_l2(got, want) <= 2 * tol(n)


@pytest.mark.parametrize("n,parts,stereo", [(64, 3, 1), (1024, 2, 0), (8192, 3, 0), (16384, 2, 1)])
def test_partitioned_convolution(lib, oracle, n, parts, stereo):
    """mufft_execute_conv_output_partitioned_batch through the host layer (the mono case needs the table of the
    mirrored-first-pass shape where one exists: 8192 reals = 4096 complex points)."""
    raw = lib.lib
    method = stereo | 4 | 8
    batch, half = 3, n // 2
    total = batch + parts - 1
    rng = np.random.default_rng(n + parts)
    x = crnd(rng, total, half) if stereo else rnd(rng, total, half)
    taps = rnd(rng, parts, half)
    plan = lib._create_plan_conv(n, 0, method)
    blk = lib._conv_get_transformed_block_size(plan)
    px, _ = managed(lib, x)
    pt, _ = managed(lib, taps)
    ps, _ = managed(lib, np.zeros(total * blk, np.uint8))
    pf, _ = managed(lib, np.zeros(parts * blk, np.uint8))
    po, out = managed(lib, np.zeros((batch, n), np.complex64 if stereo else np.float32))
    assert raw.mufft_execute_conv_input_batch(plan, 0, ps, px, total) == 0
    assert raw.mufft_execute_conv_input_batch(plan, 1, pf, pt, parts) == 0
    assert raw.mufft_execute_conv_output_partitioned_batch(plan, po, ps + (parts - 1) * blk, pf, parts, batch) == 0, raw.mufft_cuda_last_error()
    lib._free_plan_conv(plan)
    for b in range(batch):
        want = sum(oracle.conv(x[b + parts - 1 - p], taps[p], n, method)[0].astype(np.complex128 if stereo else np.float64) for p in range(parts))
        assert cpu_libs.rel_l2(out[b], want) <= 2 * tol(n)
    for p in (px, pt, ps, pf, po):
        raw.mufft_free(p)


@pytest.mark.parametrize("kind,n", [("c2c", 1024), ("c2c", 16384), ("r2c", 2048), ("r2c", 32768)])
def test_scatter_to_column_slabs(lib, oracle, kind, n):
    """mufft_execute_plan_1d_batch_scatter_async (exchange step of the slab transforms) with every 'peer' slab in
    local memory: C2C rows, and R2C rows whose Nyquist bin rides with the last slab."""
    raw = lib.lib
    peers, batch, row0 = 4, 3, 2
    rng = np.random.default_rng(n)
    if kind == "c2c":
        x = crnd(rng, batch, n)
        plan, bins, want = lib._create_plan_1d_c2c(n, -1, 0), n, [oracle.c2c(x[b], -1) for b in range(batch)]
    else:
        x = rnd(rng, batch, n)
        plan, bins, want = lib._create_plan_1d_r2c(n, 0), n // 2, [oracle.r2c(x[b]) for b in range(batch)]
    width = bins // peers
    widths = [width] * (peers - 1) + [width + (1 if kind == "r2c" else 0)]
    rows_total = batch + row0 + 1
    slabs = [raw.mufft_cuda_device_alloc(rows_total * w * 8) for w in widths]
    views = [np.ctypeslib.as_array((C.c_float * (2 * rows_total * w)).from_address(s)).view(np.complex64).reshape(rows_total, w) for s, w in zip(slabs, widths)]
    for v in views:
        v[...] = np.nan
    pin, _ = managed(lib, x)
    arr = (vp * peers)(*slabs)
    assert raw.mufft_execute_plan_1d_batch_scatter_async(plan, arr, peers, row0, pin, batch, 0) == 0, raw.mufft_cuda_last_error()
    assert raw.mufft_cuda_synchronize() == 0
    for b in range(batch):
        got = np.concatenate([v[row0 + b] for v in views])
        assert cpu_libs.rel_l2(got, want[b]) <= tol(n)
    assert all(np.all(np.isnan(v[:row0].real)) and np.all(np.isnan(v[row0 + batch:].real)) for v in views)
    lib._free_plan_1d(plan)
    raw.mufft_free(pin)
    for s in slabs:
        raw.mufft_cuda_device_free(s)


@pytest.mark.skipif(os.environ.get("MUFFT_TEST_SLOW") != "1", reason="about 8 minutes of CPU: set MUFFT_TEST_SLOW=1")
def test_reference_test_suite_passes_on_the_emulation(lib):
    """The reference's OWN test.c (compiled unchanged, FFTW replaced by oracle/fftw3_shim) against the host layer
    on the emulated CUDA side: all 3144 cases of its sweep (ref test.c:493-579) with the reference's thresholds.
    The same binary built against libmufft_b200.so runs on the GPU in tests/test_consumers.py."""
    if not os.path.isdir("/root/reference") or not os.path.exists(os.path.join(EMU_DIR, "..", "..", "oracle", "_ref", "fftw3_shim.o")):
        pytest.skip("needs /root/reference and `make -C oracle consumers`")
    subprocess.run(["make", "-s", "-C", EMU_DIR, "consumer"], check=True)
    res = subprocess.run([os.path.join(EMU_DIR, "_build", "muFFT-test-hostemu")], capture_output=True, text=True, timeout=3000)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    assert res.stdout.rstrip().endswith("All tests passed!")
    assert res.stdout.count("... Passed") > 2500


@pytest.mark.parametrize("kind,nx,ny,ndev", [("c2c", 256, 128, 2), ("c2c", 64, 64, 1), ("c2c", 32768, 8, 2), ("r2c", 512, 64, 2), ("r2c", 64, 32, 1)])
def test_slab_plan_over_the_devices_of_one_process(lib, oracle, kind, nx, ny, ndev):
    """mufft_cuda_create_plan_2d_slab_*: one 2D transform over `ndev` devices of this process (the emulated CUDA side
    has two): row pass with the exchange fused in (stores into every device's column slab), events, column pass in
    place; the _host call gathers the natural [Ny][Nx] order, i.e. what mufft_execute_plan_2d computes.  32768-point
    rows take the four-step kernel with the scatter epilogue."""
    raw = lib.lib
    raw.mufft_cuda_create_plan_2d_slab_c2c.restype, raw.mufft_cuda_create_plan_2d_slab_c2c.argtypes = vp, [u, u, i32, u, C.POINTER(i32), u]
    raw.mufft_cuda_create_plan_2d_slab_r2c.restype, raw.mufft_cuda_create_plan_2d_slab_r2c.argtypes = vp, [u, u, u, C.POINTER(i32), u]
    raw.mufft_cuda_free_plan_2d_slab.restype, raw.mufft_cuda_free_plan_2d_slab.argtypes = None, [vp]
    raw.mufft_cuda_execute_plan_2d_slab_host.restype, raw.mufft_cuda_execute_plan_2d_slab_host.argtypes = i32, [vp, vp, vp]
    raw.mufft_cuda_slab_output_columns.restype, raw.mufft_cuda_slab_output_columns.argtypes = u, [vp, u, C.POINTER(u)]
    rng = np.random.default_rng(nx + ny)
    if kind == "c2c":
        x = crnd(rng, ny, nx)
        plan = raw.mufft_cuda_create_plan_2d_slab_c2c(nx, ny, -1, ndev, None, 0)
        out = np.full((ny, nx), np.nan, np.complex64)
        want = oracle.c2c_2d(x, -1) if nx * ny <= 1 << 16 else np.fft.fft2(x.astype(np.complex128))
    else:
        x = rnd(rng, ny, nx)
        plan = raw.mufft_cuda_create_plan_2d_slab_r2c(nx, ny, ndev, None, 0)
        out = np.full((ny, nx // 2 + 1), np.nan, np.complex64)
        want = oracle.r2c_2d(x)[:, : nx // 2 + 1]
    assert plan, raw.mufft_cuda_last_error()
    first = u(0)
    widths = [raw.mufft_cuda_slab_output_columns(plan, g, C.byref(first)) for g in range(ndev)]
    assert sum(widths) == out.shape[1]
    for _ in range(2):
        assert raw.mufft_cuda_execute_plan_2d_slab_host(plan, out.ctypes.data, x.ctypes.data) == 0, raw.mufft_cuda_last_error()
    raw.mufft_cuda_free_plan_2d_slab(plan)
    assert cpu_libs.rel_l2(out, want) <= tol(nx * ny)
    # argument checks: the devices must divide the array, and exist
    assert not raw.mufft_cuda_create_plan_2d_slab_c2c(64, 64, -1, 3, None, 0)
    assert not raw.mufft_cuda_create_plan_2d_slab_c2c(64, 64, -1, 4, None, 0)  # the emulated side has two devices
    assert b"visible" in raw.mufft_cuda_last_error()


def test_small_repeated_calls_try_the_graph_path_and_fall_back(lib, oracle):
    """The latency path: the second identical small synchronous call on device-visible memory is captured into a CUDA
    graph.  The emulated CUDA side refuses captures, so here the host layer must try exactly once, switch the path off
    for the plan and keep giving right answers the ordinary way."""
    raw = lib.lib
    raw.emu_capture_attempts.restype, raw.emu_capture_attempts.argtypes = u, [i32]
    x = crnd(np.random.default_rng(3), 64)
    px, vx = managed(lib, x)
    py, vy = managed(lib, np.zeros(64, np.complex64))
    plan = lib._create_plan_1d_c2c(64, -1, 0)
    raw.emu_capture_attempts(1)
    for k in range(5):
        vx[:] = x * (k + 1)
        raw.mufft_cuda_launch_count(1)
        lib._execute_plan_1d(plan, py, px)
        assert raw.mufft_cuda_launch_count(1) == 1
        assert cpu_libs.rel_l2(vy, oracle.c2c((x * (k + 1)).astype(np.complex64), -1)) <= tol(64)
    assert raw.emu_capture_attempts(1) == 1
    lib._free_plan_1d(plan)
    raw.mufft_free(px)
    raw.mufft_free(py)


@pytest.mark.parametrize("ny,flags,launches", [(16384, 0, 2), (16384, TWO_KERNEL, 3), (4096, 0, 3), (2048, 0, 2)])
def test_long_columns_of_a_2d_plan(lib, oracle, ny, flags, launches):
    """2D plans with more than 2048 rows: the column pass is ONE launch of the four-step column kernel (16384 / 32768
    rows), else or with MUFFT_CUDA_FLAG_TWO_KERNEL two sweeps of the column kernel; up to 2048 rows one column kernel.
    (+ 1 launch for the rows.)"""
    raw = lib.lib
    nx = 40 if False else 32
    x = crnd(np.random.default_rng(ny), ny, nx)
    pin, _ = managed(lib, x)
    pout, yv = managed(lib, np.zeros((ny, nx), np.complex64))
    plan = lib._create_plan_2d_c2c(nx, ny, 1, flags)
    assert plan
    raw.mufft_cuda_launch_count(1)
    assert raw.mufft_execute_plan_2d_batch(plan, pout, pin, 1) == 0, raw.mufft_cuda_last_error()
    assert raw.mufft_cuda_launch_count(1) == launches
    lib._free_plan_2d(plan)
    assert cpu_libs.rel_l2(yv, oracle.c2c_2d(x, 1)) <= tol(nx * ny)
    raw.mufft_free(pin)
    raw.mufft_free(pout)


def test_fused_2d_c2r_through_the_host_layer(lib, oracle):
    """C2R images with rows of 2048 complex points (4096 reals) x 256 rows: ONE launch of the fused kernel - columns into
    the plan's ring of intermediates, C2R rows out of it - for a batch longer than the ring."""
    raw = lib.lib
    nx, ny, batch = 4096, 256, 6
    rng = np.random.default_rng(9)
    real = rnd(rng, batch, ny, nx)
    x = np.zeros((batch, ny, nx), np.complex64)
    x[:, :, : nx // 2 + 1] = np.fft.rfft2(real.astype(np.float64)).astype(np.complex64)
    plan = lib._create_plan_2d_c2r(nx, ny, 0)
    assert plan and raw.mufft_cuda_plan_2d_paths(plan) & 4
    pin, _ = managed(lib, x)
    pout, yv = managed(lib, np.zeros((batch, ny, nx), np.float32))
    raw.mufft_cuda_launch_count(1)
    assert raw.mufft_execute_plan_2d_batch(plan, pout, pin, batch) == 0, raw.mufft_cuda_last_error()
    assert raw.mufft_cuda_launch_count(1) == 1
    lib._free_plan_2d(plan)
    for b in (0, batch - 1):
        assert cpu_libs.rel_l2(yv[b], oracle.c2r_2d(x[b])) <= tol(nx * ny)
    raw.mufft_free(pin)
    raw.mufft_free(pout)
