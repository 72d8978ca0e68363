"""Batched entry points of include/mufft_cuda.h against the oracle: device-resident, strided,
pinned-host and pageable-host buffers.  Batches are large enough that the persistent kernels run
several tiles per CTA, and ragged enough that the last group of a row kernel has idle transform slots."""
import ctypes
import math

import numpy as np
import pytest

from cpu_libs import FLAG_FULL_R2C, FLAG_ZERO_PAD, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import mufft_b200
    if mufft_b200.mufft_cuda_device_count() < 1:
        pytest.fail("no CUDA device visible")
    return mufft_b200


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def tol(n):
    return 1e-6 * max(1.0, math.log2(n))


def sample_indices(batch, count=24, seed=0):
    rng = np.random.default_rng(seed)
    idx = set(range(min(batch, 4))) | set(range(max(0, batch - 4), batch))
    idx |= set(int(i) for i in rng.integers(0, batch, size=min(count, batch)))
    return sorted(idx)


def dev_random(torch, shape, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return torch.rand(shape, dtype=torch.float32, device="cuda", generator=g) - 0.5


def batch_for(n):
    """Enough transforms for > 3 groups per persistent CTA on a 148-SM part, not a multiple of anything."""
    return max(7, min(300000, (6 << 20) // n + 13))


@pytest.mark.parametrize("n", [2, 4, 16, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
@pytest.mark.parametrize("direction", [-1, 1])
def test_c2c_batch_device(m, torch, oracle, n, direction):
    batch = batch_for(n)
    x = dev_random(torch, (batch, 2 * n), n + direction)
    y = torch.zeros_like(x)
    plan = m.mufft_create_plan_1d_c2c(n, direction, 0)
    assert plan
    assert m.mufft_execute_plan_1d_batch(plan, y.data_ptr(), x.data_ptr(), batch, 0, 0) == 0, m.mufft_cuda_last_error()
    # secondary check on every transform (cuFFT through torch), primary on a sample (oracle)
    xc = torch.view_as_complex(x.view(batch, n, 2))
    ref = torch.fft.fft(xc, dim=1) if direction < 0 else torch.fft.ifft(xc, dim=1, norm="forward")
    yc = torch.view_as_complex(y.view(batch, n, 2))
    err = (torch.linalg.vector_norm(yc - ref, dim=1) / torch.linalg.vector_norm(ref, dim=1)).max().item()
    assert err <= tol(n), f"worst transform rel L2 {err:.3e}"
    idx = sample_indices(batch)
    xs, ys = x[idx].cpu().numpy().view(np.complex64), y[idx].cpu().numpy().view(np.complex64)
    for i in range(len(idx)):
        assert rel_l2(ys[i], oracle.c2c(xs[i], direction)) <= tol(n)
    # async flavour on the current stream gives the same bits
    y2 = torch.zeros_like(x)
    assert m.mufft_execute_plan_1d_batch_async(plan, y2.data_ptr(), x.data_ptr(), batch, 0, 0) == 0
    assert m.mufft_cuda_synchronize() == 0
    assert torch.equal(y, y2)
    m.mufft_free_plan_1d(plan)


@pytest.mark.parametrize("n", [64, 1024, 4096])
def test_c2c_batch_strided_and_zero_pad(m, torch, oracle, n):
    batch = 1500
    in_pitch, out_pitch = n // 2 + 6, n + 10  # complex elements; input rows are only N/2 long (zero pad)
    x = dev_random(torch, (batch, 2 * in_pitch), 5)
    y = torch.full((batch, 2 * out_pitch), 7.0, dtype=torch.float32, device="cuda")
    plan = m.mufft_create_plan_1d_c2c(n, -1, FLAG_ZERO_PAD)
    assert m.mufft_execute_plan_1d_batch(plan, y.data_ptr(), x.data_ptr(), batch, out_pitch * 8, in_pitch * 8) == 0
    m.mufft_free_plan_1d(plan)
    assert torch.all(y[:, 2 * n:] == 7.0), "wrote beyond N outputs of a row"
    idx = sample_indices(batch)
    xs = x[idx].cpu().numpy().view(np.complex64)[:, : n // 2]
    ys = y[idx].cpu().numpy().view(np.complex64)[:, :n]
    for i in range(len(idx)):
        assert rel_l2(ys[i], oracle.c2c(xs[i], -1, FLAG_ZERO_PAD, n=n)) <= tol(n)


@pytest.mark.parametrize("n", [4, 16, 128, 1024, 4096, 8192, 16384, 32768])
@pytest.mark.parametrize("flags", [0, FLAG_FULL_R2C, FLAG_ZERO_PAD])
@pytest.mark.parametrize("pitch", ["tight", "aligned"])
def test_r2c_c2r_batch_device(m, torch, oracle, n, flags, pitch):
    batch = batch_for(n // 2)
    n_in = n // 2 if flags & FLAG_ZERO_PAD else n
    bins = n if flags & FLAG_FULL_R2C else n // 2 + 1
    spec_pitch = bins if pitch == "tight" else (bins + 3) // 2 * 2  # complex elements, even = 16-byte rows
    x = dev_random(torch, (batch, n_in), n)
    spec = torch.zeros((batch, 2 * spec_pitch), dtype=torch.float32, device="cuda")
    fwd = m.mufft_create_plan_1d_r2c(n, flags)
    assert fwd
    assert m.mufft_execute_plan_1d_batch(fwd, spec.data_ptr(), x.data_ptr(), batch, spec_pitch * 8, 0) == 0
    m.mufft_free_plan_1d(fwd)
    idx = sample_indices(batch)
    xs = x[idx].cpu().numpy()
    ss = spec[idx].cpu().numpy().view(np.complex64)[:, :bins]
    for i in range(len(idx)):
        assert rel_l2(ss[i], oracle.r2c(xs[i], flags, n=n)) <= tol(n)
    if flags == 0:
        back = torch.zeros((batch, n), dtype=torch.float32, device="cuda")
        inv = m.mufft_create_plan_1d_c2r(n, 0)
        assert m.mufft_execute_plan_1d_batch(inv, back.data_ptr(), spec.data_ptr(), batch, 0, spec_pitch * 8) == 0
        m.mufft_free_plan_1d(inv)
        bs = back[idx].cpu().numpy()
        for i in range(len(idx)):
            assert rel_l2(bs[i], oracle.c2r(ss[i], n)) <= tol(n)
        # round trip, every transform: C2R(R2C(x)) = N x  (ref fft.md:108-113, unnormalised)
        err = (torch.linalg.vector_norm(back / n - x, dim=1) / torch.linalg.vector_norm(x, dim=1)).max().item()
        assert err <= 2 * tol(n)


@pytest.mark.parametrize("kind", ["pinned", "pageable"])
@pytest.mark.parametrize("n", [256, 4096])
def test_c2c_batch_host_memory(m, torch, oracle, n, kind):
    """Host buffers are staged through the plan's device buffers in pipelined chunks."""
    batch = (80 << 20) // (8 * n) + 5  # > 2 staging chunks
    x = (np.random.default_rng(3).random((batch, 2 * n), np.float32) - np.float32(0.5))
    if kind == "pinned":
        xt = torch.from_numpy(x).pin_memory()
        yt = torch.zeros((batch, 2 * n), dtype=torch.float32).pin_memory()
        xp, yp = xt.data_ptr(), yt.data_ptr()
    else:
        y = np.zeros_like(x)
        xp, yp = x.ctypes.data, y.ctypes.data
    plan = m.mufft_create_plan_1d_c2c(n, -1, 0)
    assert m.mufft_execute_plan_1d_batch(plan, yp, xp, batch, 0, 0) == 0, m.mufft_cuda_last_error()
    # the async flavour must refuse pointers the device cannot reach
    if kind == "pageable":
        assert m.mufft_execute_plan_1d_batch_async(plan, yp, xp, batch, 0, 0) != 0
    m.mufft_free_plan_1d(plan)
    got = (yt.numpy() if kind == "pinned" else y).view(np.complex64)
    xd = torch.from_numpy(x).cuda()
    ref = torch.fft.fft(torch.view_as_complex(xd.view(batch, n, 2)), dim=1).cpu().numpy()
    assert rel_l2(got, ref) <= tol(n)
    for i in sample_indices(batch, 8):
        assert rel_l2(got[i], oracle.c2c(x[i].view(np.complex64), -1)) <= tol(n)


@pytest.mark.parametrize("n,batch", [(8192, 1500), (8192, 1024), (16384, 1100)])
def test_conv_stereo_large_batch_prefetch_rows(m, torch, oracle, n, batch):
    """Stereo convolution in batches of 1024 blocks or more: both stages (zero-padded forward transform, spectral multiply
    + inverse transform) run on the persistent prefetching row kernel (rows_prefetch_kernel.cuh).  Every block against
    torch's FFT convolution, a sample against the oracle."""
    method = 1 | 4 | 8
    plan = m.mufft_create_plan_conv(n, 0, method)
    assert plan
    blk = m.mufft_conv_get_transformed_block_size(plan) // 8
    a = dev_random(torch, (batch, n), 21)            # n/2 complex samples per block
    filt = dev_random(torch, (1, n // 2), 22)
    sa = torch.zeros((batch, 2 * blk), dtype=torch.float32, device="cuda")
    sf = torch.zeros((1, 2 * blk), dtype=torch.float32, device="cuda")
    out = torch.zeros((batch, 2 * n), dtype=torch.float32, device="cuda")
    assert m.mufft_execute_conv_input_batch(plan, 1, sf.data_ptr(), filt.data_ptr(), 1) == 0
    m.mufft_cuda_launch_count(1)
    assert m.mufft_execute_conv_input_batch(plan, 0, sa.data_ptr(), a.data_ptr(), batch) == 0, m.mufft_cuda_last_error()
    assert m.mufft_execute_conv_output_batch(plan, out.data_ptr(), sa.data_ptr(), sf.data_ptr(), batch, 0) == 0, m.mufft_cuda_last_error()
    assert m.mufft_cuda_launch_count(1) == 2
    m.mufft_free_plan_conv(plan)
    ac = torch.view_as_complex(a.view(batch, n // 2, 2))
    fr = filt[0]
    spec = torch.fft.fft(ac, n=n, dim=1) * torch.fft.fft(fr.to(torch.complex64), n=n)
    ref = torch.fft.ifft(spec, dim=1)  # the plan normalises by 1/n as ifft does
    got = torch.view_as_complex(out.view(batch, n, 2))
    err = (torch.linalg.vector_norm(got - ref, dim=1) / torch.linalg.vector_norm(ref, dim=1)).max().item()
    assert err <= 2 * tol(n), f"worst block rel L2 {err:.3e}"
    f = filt[0].cpu().numpy()
    for i in (0, batch // 2, batch - 1):
        want, _, _ = oracle.conv(a[i].cpu().numpy().view(np.complex64), f, n, method)
        assert rel_l2(out[i].cpu().numpy().view(np.complex64), want) <= 2 * tol(n)


@pytest.mark.parametrize("n", [64, 2048, 8192, 32768])
@pytest.mark.parametrize("stereo", [0, 1])
def test_conv_batch_device(m, torch, oracle, n, stereo):
    method = stereo | 4 | 8  # both inputs zero padded, as bench.c:243-310 uses the API
    batch = batch_for(n) // 2 + 3
    plan = m.mufft_create_plan_conv(n, 0, method)
    assert plan
    blk = m.mufft_conv_get_transformed_block_size(plan) // 8
    a = dev_random(torch, (batch, n if stereo else n // 2), 11)       # n/2 stereo (complex) or mono samples
    filt = dev_random(torch, (1, n // 2), 12)
    sa = torch.zeros((batch, 2 * blk), dtype=torch.float32, device="cuda")
    sf = torch.zeros((1, 2 * blk), dtype=torch.float32, device="cuda")
    out = torch.zeros((batch, 2 * n if stereo else n), dtype=torch.float32, device="cuda")
    assert m.mufft_execute_conv_input_batch(plan, 1, sf.data_ptr(), filt.data_ptr(), 1) == 0
    assert m.mufft_execute_conv_input_batch(plan, 0, sa.data_ptr(), a.data_ptr(), batch) == 0
    assert m.mufft_execute_conv_output_batch(plan, out.data_ptr(), sa.data_ptr(), sf.data_ptr(), batch, 0) == 0
    m.mufft_free_plan_conv(plan)
    f = filt[0].cpu().numpy()
    idx = sample_indices(batch, 10)
    as_, os_ = a[idx].cpu().numpy(), out[idx].cpu().numpy()
    for i in range(len(idx)):
        ai = as_[i].view(np.complex64) if stereo else as_[i]
        want, _, _ = oracle.conv(ai, f, n, method)
        got = os_[i].view(np.complex64) if stereo else os_[i]
        assert rel_l2(got, want) <= 2 * tol(n)
        assert np.max(np.abs(got - want)) < 2e-6 * math.sqrt(n)  # test.c:431,474


@pytest.mark.parametrize("n", [4, 8, 64, 1024, 2048, 8192, 16384, 32768])
@pytest.mark.parametrize("stereo", [0, 1])
@pytest.mark.parametrize("padded", [1, 0])
def test_conv_fused_batch_device(m, torch, oracle, n, stereo, padded):
    """mufft_execute_conv_batch (one kernel per batch up to the engine limit, the two library steps
    through scratch beyond it) against the oracle's conv_input + conv_output of every sampled block,
    with the reference's own criterion (test.c:431,474) and against the two-call GPU path."""
    method = stereo | (12 if padded else 0)
    batch = batch_for(n) // 2 + 3
    plan = m.mufft_create_plan_conv(n, 0, method)
    assert plan
    blk = m.mufft_conv_get_transformed_block_size(plan) // 8
    samples = n // 2 if padded else n
    a = dev_random(torch, (batch, 2 * samples if stereo else samples), 21)
    filt = dev_random(torch, (1, samples), 22)
    sf = torch.zeros((1, 2 * blk), dtype=torch.float32, device="cuda")
    out = torch.zeros((batch, 2 * n if stereo else n), dtype=torch.float32, device="cuda")
    assert m.mufft_execute_conv_input_batch(plan, 1, sf.data_ptr(), filt.data_ptr(), 1) == 0
    assert m.mufft_execute_conv_batch(plan, out.data_ptr(), a.data_ptr(), sf.data_ptr(), batch, 0) == 0, m.mufft_cuda_last_error()
    # the two-call path on the same inputs
    sa = torch.zeros((batch, 2 * blk), dtype=torch.float32, device="cuda")
    out2 = torch.zeros_like(out)
    assert m.mufft_execute_conv_input_batch(plan, 0, sa.data_ptr(), a.data_ptr(), batch) == 0
    assert m.mufft_execute_conv_output_batch(plan, out2.data_ptr(), sa.data_ptr(), sf.data_ptr(), batch, 0) == 0
    # one filter per block (stride = block size) must give the same as the shared filter
    if batch <= 4096:
        sfs = sf.repeat(batch, 1).contiguous()
        out3 = torch.zeros_like(out)
        assert m.mufft_execute_conv_batch(plan, out3.data_ptr(), a.data_ptr(), sfs.data_ptr(), batch, blk * 8) == 0
        assert torch.equal(out3, out)
    m.mufft_free_plan_conv(plan)
    scale = float(out2.abs().max())
    assert float((out - out2).abs().max()) <= 4e-6 * math.sqrt(n) * max(scale, 1e-3) + 1e-7
    f = filt[0].cpu().numpy()
    idx = sample_indices(batch, 10)
    as_, os_ = a[idx].cpu().numpy(), out[idx].cpu().numpy()
    for i in range(len(idx)):
        ai = as_[i].view(np.complex64) if stereo else as_[i]
        want, _, _ = oracle.conv(ai, f, n, method)
        got = os_[i].view(np.complex64) if stereo else os_[i]
        assert rel_l2(got, want) <= 2 * tol(n)
        assert np.max(np.abs(got - want)) < 2e-6 * math.sqrt(n) * (1 if padded else 2)


@pytest.mark.parametrize("n,parts", [(64, 1), (64, 5), (1024, 2), (2048, 3), (8192, 8), (16384, 4), (32768, 2)])
@pytest.mark.parametrize("stereo", [0, 1])
def test_conv_partitioned_device(m, torch, oracle, n, parts, stereo):
    """Uniformly partitioned convolution (filter of parts * N/2 taps): one inverse transform per block
    on the sum of `parts` spectral products, against the oracle's per-partition
    conv_input + conv_output results added up in float64 (what a caller of the reference does,
    fft.h:170-174)."""
    method = stereo | 4 | 8
    batch = min(batch_for(n) // 4 + 3, 2000)
    plan = m.mufft_create_plan_conv(n, 0, method)
    assert plan
    blk = m.mufft_conv_get_transformed_block_size(plan) // 8
    half = n // 2
    total = batch + parts - 1                                  # parts-1 blocks of history in front
    x = dev_random(torch, (total, 2 * half if stereo else half), 31)
    taps = dev_random(torch, (parts, half), 32)
    spectra = torch.zeros((total, 2 * blk), dtype=torch.float32, device="cuda")
    fspec = torch.zeros((parts, 2 * blk), dtype=torch.float32, device="cuda")
    out = torch.zeros((batch, 2 * n if stereo else n), dtype=torch.float32, device="cuda")
    assert m.mufft_execute_conv_input_batch(plan, 0, spectra.data_ptr(), x.data_ptr(), total) == 0
    assert m.mufft_execute_conv_input_batch(plan, 1, fspec.data_ptr(), taps.data_ptr(), parts) == 0
    first = spectra.data_ptr() + (parts - 1) * blk * 8
    err = m.mufft_execute_conv_output_partitioned_batch(plan, out.data_ptr(), first, fspec.data_ptr(), parts, batch)
    if stereo and n > 16384:  # MCU_ENGINE_BIG_N: the longest transform of the on-chip engine
        assert err != 0 and b"on-chip engine" in m.mufft_cuda_last_error()
        m.mufft_free_plan_conv(plan)
        return
    assert err == 0, m.mufft_cuda_last_error()
    if parts == 1:  # one partition = plain conv_output
        out1 = torch.zeros_like(out)
        assert m.mufft_execute_conv_output_batch(plan, out1.data_ptr(), first, fspec.data_ptr(), batch, 0) == 0
        assert float((out - out1).abs().max()) <= 1e-5
    m.mufft_free_plan_conv(plan)
    xs, hs = x.cpu().numpy(), taps.cpu().numpy()
    for b in sample_indices(batch, 3, seed=n):
        want = np.zeros(n, np.complex128 if stereo else np.float64)
        for p in range(parts):
            blk_in = xs[b + parts - 1 - p]
            piece, _, _ = oracle.conv(blk_in.view(np.complex64) if stereo else blk_in, hs[p], n, method)
            want += piece
        got = out[b].cpu().numpy()
        got = got.view(np.complex64) if stereo else got
        assert rel_l2(got, want) <= 2 * tol(n)


def test_conv_fused_batch_host_memory(m, torch, oracle):
    """Pageable host buffers go through the staged path (chunks over several streams)."""
    n, batch, method = 1024, 3000, 1 | 4 | 8
    plan = m.mufft_create_plan_conv(n, 0, method)
    blk = m.mufft_conv_get_transformed_block_size(plan) // 8
    rng = np.random.default_rng(5)
    a = (rng.random((batch, n), np.float32) - np.float32(0.5))          # n/2 complex per block
    taps = (rng.random(n // 2, np.float32) - np.float32(0.5))
    sf = np.zeros(2 * blk, np.float32)
    out = np.zeros((batch, 2 * n), np.float32)
    m.mufft_execute_conv_input(plan, 1, sf.ctypes.data, taps.ctypes.data)
    assert m.mufft_execute_conv_batch(plan, out.ctypes.data, a.ctypes.data, sf.ctypes.data, batch, 0) == 0, m.mufft_cuda_last_error()
    m.mufft_free_plan_conv(plan)
    for i in (0, 1, batch // 2, batch - 1):
        want, _, _ = oracle.conv(a[i].view(np.complex64), taps, n, method)
        assert rel_l2(out[i].view(np.complex64), want) <= 2 * tol(n)


SEPARATE_2D, TWO_KERNEL, GENERIC_PASSES, FOUR_STEP_16K = 1 << 20, 1 << 21, 1 << 22, 1 << 23  # mufft_cuda.h plan flags


@pytest.mark.parametrize("nx,ny", [(64, 64), (256, 128), (256, 256), (512, 512), (512, 2048), (2048, 256), (1024, 1024), (128, 2048),
                                   (8192, 8), (16, 4096), (32, 16384), (16384, 32), (64, 8192)])
@pytest.mark.parametrize("separate", [0, 1])
def test_c2c_2d_batch_device(m, torch, oracle, nx, ny, separate):
    """Default: one persistent kernel for the whole batch where the shape has one (256..2048 per
    side); MUFFT_CUDA_FLAG_SEPARATE_2D: row kernel + column kernel per group of images."""
    batch = max(3, min(40, (64 << 20) // (8 * nx * ny)))
    x = dev_random(torch, (batch, ny, nx, 2), nx + ny)
    y = torch.zeros_like(x)
    fusable = all(256 <= v <= 2048 for v in (nx, ny))
    for direction in (-1, 1):
        plan = m.mufft_create_plan_2d_c2c(nx, ny, direction, SEPARATE_2D if separate else 0)
        assert plan
        assert bool(m.mufft_cuda_plan_2d_paths(plan) & 4) == (fusable and not separate)
        m.mufft_cuda_launch_count(1)
        assert m.mufft_execute_plan_2d_batch(plan, y.data_ptr(), x.data_ptr(), batch) == 0, m.mufft_cuda_last_error()
        if fusable and not separate:
            assert m.mufft_cuda_launch_count(1) == 1, "the fused path is one launch for the whole batch"
        m.mufft_free_plan_2d(plan)
        xc = torch.view_as_complex(x)
        ref = torch.fft.fft2(xc) if direction < 0 else torch.fft.ifft2(xc, norm="forward")
        yc = torch.view_as_complex(y)
        err = (torch.linalg.vector_norm((yc - ref).reshape(batch, -1), dim=1) /
               torch.linalg.vector_norm(ref.reshape(batch, -1), dim=1)).max().item()
        assert err <= tol(nx * ny)
        for i in (0, batch - 1):
            want = oracle.c2c_2d(xc[i].cpu().numpy(), direction)
            assert rel_l2(yc[i].cpu().numpy(), want) <= tol(nx * ny)


@pytest.mark.parametrize("nx,ny", [(512, 256), (1024, 1024), (2048, 512), (4096, 256), (64, 128), (8192, 64)])
@pytest.mark.parametrize("flags", [0, FLAG_FULL_R2C, SEPARATE_2D])
def test_r2c_2d_batch_device(m, torch, oracle, nx, ny, flags):
    """2D R2C images of nx reals per row: the fused persistent kernel where the shape has one (rows with the resolve
    epilogue straight into the output, columns in place: one launch), else / with SEPARATE_2D row kernel + column kernel."""
    batch = max(3, min(24, (64 << 20) // (8 * nx * ny)))
    g = torch.Generator(device="cuda")
    g.manual_seed(nx + ny)
    x = torch.rand((batch, ny, nx), dtype=torch.float32, device="cuda", generator=g) - 0.5
    out = torch.zeros((batch, ny, nx, 2), dtype=torch.float32, device="cuda")
    fused = all(256 <= v <= 2048 for v in (nx // 2, ny)) and not flags & SEPARATE_2D
    plan = m.mufft_create_plan_2d_r2c(nx, ny, flags)
    assert plan
    m.mufft_cuda_launch_count(1)
    assert m.mufft_execute_plan_2d_batch(plan, out.data_ptr(), x.data_ptr(), batch) == 0, m.mufft_cuda_last_error()
    assert m.mufft_cuda_launch_count(1) == (1 if fused else 2)
    m.mufft_free_plan_2d(plan)
    width = nx if flags & FLAG_FULL_R2C else nx // 2 + 1
    got = torch.view_as_complex(out)[:, :, :width]
    ref = torch.fft.fft2(x)[:, :, :width]
    err = (torch.linalg.vector_norm((got - ref).reshape(batch, -1), dim=1) / torch.linalg.vector_norm(ref.reshape(batch, -1), dim=1)).max().item()
    assert err <= tol(nx * ny)
    want = oracle.r2c_2d(x[batch - 1].cpu().numpy(), flags & FLAG_FULL_R2C)[:, :width]
    assert rel_l2(got[batch - 1].cpu().numpy(), want) <= tol(nx * ny)


@pytest.mark.parametrize("nx,ny,batch", [(512, 256, 41), (1024, 1024, 19), (1024, 256, 33), (2048, 512, 7), (4096, 256, 9), (1024, 1024, 1), (64, 128, 5), (8192, 64, 3)])
@pytest.mark.parametrize("flags", [0, SEPARATE_2D])
def test_c2r_2d_batch_device(m, torch, oracle, nx, ny, batch, flags):
    """2D C2R images (spectrum rows of pitch nx complex, nx/2 + 1 carry data; dense real output): the fused kernel does
    the columns first into the plan's ring of intermediates and the C2R rows out of it (ref fft.c:861-916) - batches
    longer than the ring exercise the slot hand-over - else / with SEPARATE_2D column kernel + row kernel."""
    g = torch.Generator(device="cuda")
    g.manual_seed(nx + ny + batch)
    real = torch.rand((batch, ny, nx), dtype=torch.float32, device="cuda", generator=g) - 0.5
    spec = torch.full((batch, ny, nx), float("nan"), dtype=torch.complex64, device="cuda")  # the padding columns are never read
    spec[:, :, : nx // 2 + 1] = torch.fft.rfft2(real.double()).to(torch.complex64)
    out = torch.zeros((batch, ny, nx), dtype=torch.float32, device="cuda")
    fused = 512 <= nx // 2 <= 2048 and 256 <= ny <= 2048 and not flags & SEPARATE_2D  # shorter rows: separate kernels are faster
    plan = m.mufft_create_plan_2d_c2r(nx, ny, flags)
    assert plan
    for _ in range(2):  # twice: the ring and the counters are reused
        m.mufft_cuda_launch_count(1)
        assert m.mufft_execute_plan_2d_batch(plan, out.data_ptr(), torch.view_as_real(spec).data_ptr(), batch) == 0, m.mufft_cuda_last_error()
        assert m.mufft_cuda_launch_count(1) == (1 if fused else 2)
    m.mufft_free_plan_2d(plan)
    want = real.double() * (nx * ny)
    err = (torch.linalg.vector_norm((out.double() - want).reshape(batch, -1), dim=1) / torch.linalg.vector_norm(want.reshape(batch, -1), dim=1)).max().item()
    assert err <= tol(nx * ny)
    clean = torch.nan_to_num(torch.view_as_real(spec[batch - 1]), nan=0.0).cpu().numpy().view(np.complex64).reshape(ny, nx)
    assert rel_l2(out[batch - 1].cpu().numpy(), oracle.c2r_2d(clean)) <= tol(nx * ny)


@pytest.mark.parametrize("case", ["c2c2d-fused", "c2c2d-separate", "r2c2d", "c2r2d", "c2c-long"])
def test_staged_host_batches_that_share_plan_scratch(m, torch, oracle, case):
    """Pageable host buffers, batches long enough for several staging chunks on different streams:
    the kernels of consecutive chunks share plan-owned scratch (2D intermediate, ticket counters,
    two-kernel scratch), so they must be chained.  Every image / transform is checked against the
    device-resident result of the same call."""
    rng = np.random.default_rng(17)
    if case == "c2c-long":
        n, batch = 32768, 600
        plan = m.mufft_create_plan_1d_c2c(n, -1, 0)
        x = rng.random((batch, 2 * n), np.float32) - np.float32(0.5)
        out = np.zeros_like(x)
        assert m.mufft_execute_plan_1d_batch(plan, out.ctypes.data, x.ctypes.data, batch, 0, 0) == 0
        xd = torch.from_numpy(x).cuda()
        od = torch.zeros_like(xd)
        assert m.mufft_execute_plan_1d_batch(plan, od.data_ptr(), xd.data_ptr(), batch, 0, 0) == 0
        m.mufft_free_plan_1d(plan)
        assert np.array_equal(out, od.cpu().numpy())
        want = oracle.c2c(x[batch - 1].view(np.complex64), -1)
        assert rel_l2(out[batch - 1].view(np.complex64), want) <= tol(n)
        return
    nx = ny = 256
    batch = 500
    if case.startswith("c2c2d"):
        plan = m.mufft_create_plan_2d_c2c(nx, ny, -1, 0 if case == "c2c2d-fused" else SEPARATE_2D)
        x = rng.random((batch, ny * nx * 2), np.float32) - np.float32(0.5)
        out = np.zeros_like(x)
    elif case == "r2c2d":
        plan = m.mufft_create_plan_2d_r2c(nx, ny, 0)
        x = rng.random((batch, ny * nx), np.float32) - np.float32(0.5)
        out = np.zeros((batch, ny * nx * 2), np.float32)
    else:
        plan = m.mufft_create_plan_2d_c2r(nx, ny, 0)
        spec = np.fft.rfft2(rng.random((batch, ny, nx)) - 0.5).astype(np.complex64)
        x = np.zeros((batch, ny, nx), np.complex64)
        x[:, :, : nx // 2 + 1] = spec
        x = x.view(np.float32).reshape(batch, -1)
        out = np.zeros((batch, ny * nx), np.float32)
    try:
        assert m.mufft_execute_plan_2d_batch(plan, out.ctypes.data, x.ctypes.data, batch) == 0, m.mufft_cuda_last_error()
        xd = torch.from_numpy(x).cuda()
        od = torch.zeros(out.shape, dtype=torch.float32, device="cuda")
        assert m.mufft_execute_plan_2d_batch(plan, od.data_ptr(), xd.data_ptr(), batch) == 0
    finally:
        m.mufft_free_plan_2d(plan)
    got, dev = out.reshape(batch, ny, -1), od.cpu().numpy().reshape(batch, ny, -1)
    if case == "r2c2d":  # only columns 0..Nx/2 of the pitch-Nx complex rows are defined
        got, dev = got[:, :, : 2 * (nx // 2 + 1)], dev[:, :, : 2 * (nx // 2 + 1)]
    assert np.array_equal(got, dev)


@pytest.mark.parametrize("n", [16384, 32768, 65536, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22])
@pytest.mark.parametrize("path", ["default", "two_kernel", "generic", "four_step_16k"])
def test_large_n(m, torch, oracle, n, path):
    """N above one CTA's reach.  Default: the single-launch four-step kernel (32768 / 65536 points, intermediate in L2),
    the two-kernel split (N = N1*N2) up to 2^21, generic radix passes beyond; the plan flags force the fallbacks.
    C2C both directions, zero pad, R2C and C2R of twice the length."""
    if path == "generic" and n > (1 << 19):
        pytest.skip("generic passes are covered at the smaller lengths")
    if path == "two_kernel" and n > (1 << 21):
        pytest.skip("no two-kernel split beyond 2^21")
    if path == "four_step_16k" and n != 16384:
        pytest.skip("the flag only concerns 16384 points")
    flags = {"default": 0, "two_kernel": TWO_KERNEL, "generic": GENERIC_PASSES, "four_step_16k": FOUR_STEP_16K}[path]
    one_launch = (path == "default" and n <= 65536) or path == "four_step_16k"
    batch = 3 if n > 65536 else 301  # more transforms than resident CTAs: the four-step kernel reuses its scratch slots
    x = dev_random(torch, (batch, 2 * n), n)
    y = torch.zeros_like(x)
    xc = torch.view_as_complex(x.view(batch, n, 2))
    for direction in (-1, 1):
        plan = m.mufft_create_plan_1d_c2c(n, direction, flags)
        assert plan
        assert m.mufft_cuda_plan_1d_uses_engine(plan) == (1 if one_launch else 0)
        m.mufft_cuda_launch_count(1)
        assert m.mufft_execute_plan_1d_batch(plan, y.data_ptr(), x.data_ptr(), batch, 0, 0) == 0, m.mufft_cuda_last_error()
        if one_launch:
            assert m.mufft_cuda_launch_count(1) == 1
        m.mufft_free_plan_1d(plan)
        ref = torch.fft.fft(xc, dim=1) if direction < 0 else torch.fft.ifft(xc, dim=1, norm="forward")
        yc = torch.view_as_complex(y.view(batch, n, 2))
        err = (torch.linalg.vector_norm(yc - ref, dim=1) / torch.linalg.vector_norm(ref, dim=1)).max().item()
        assert err <= tol(n)
        if n <= (1 << 20):
            got = yc[batch - 1].cpu().numpy()
            assert rel_l2(got, oracle.c2c(xc[batch - 1].cpu().numpy(), direction)) <= tol(n)
    # zero-padded forward: the input rows really are N/2 long
    xh = dev_random(torch, (batch, n), 7)
    plan = m.mufft_create_plan_1d_c2c(n, -1, FLAG_ZERO_PAD | flags)
    assert m.mufft_execute_plan_1d_batch(plan, y.data_ptr(), xh.data_ptr(), batch, 0, 0) == 0
    m.mufft_free_plan_1d(plan)
    xp = torch.zeros((batch, n), dtype=torch.complex64, device="cuda")
    xp[:, : n // 2] = torch.view_as_complex(xh.view(batch, n // 2, 2))
    ref = torch.fft.fft(xp, dim=1)
    yc = torch.view_as_complex(y.view(batch, n, 2))
    assert (torch.linalg.vector_norm(yc - ref) / torch.linalg.vector_norm(ref)).item() <= tol(n)
    # R2C + C2R of real length 2N (complex engine length N)
    batch = min(batch, 37)
    r = dev_random(torch, (batch, 2 * n), 9)
    spec = torch.zeros((batch, 2 * (n + 1)), dtype=torch.float32, device="cuda")
    fwd = m.mufft_create_plan_1d_r2c(2 * n, flags)
    assert m.mufft_execute_plan_1d_batch(fwd, spec.data_ptr(), r.data_ptr(), batch, 0, 0) == 0
    m.mufft_free_plan_1d(fwd)
    ref = torch.fft.rfft(r, dim=1)
    sc = torch.view_as_complex(spec.view(batch, n + 1, 2))
    assert (torch.linalg.vector_norm(sc - ref) / torch.linalg.vector_norm(ref)).item() <= tol(2 * n)
    back = torch.zeros_like(r)
    inv = m.mufft_create_plan_1d_c2r(2 * n, flags)
    assert m.mufft_execute_plan_1d_batch(inv, back.data_ptr(), spec.data_ptr(), batch, 0, 0) == 0
    m.mufft_free_plan_1d(inv)
    assert (torch.linalg.vector_norm(back / (2 * n) - r) / torch.linalg.vector_norm(r)).item() <= 2 * tol(2 * n)


def test_plans_with_different_path_flags_run_side_by_side(m, torch):
    """There is no process-global switch: a 16384-point plan on the engine, one on the four-step kernel and one on the
    two-kernel split coexist and give the same transform."""
    n, batch = 16384, 37
    x = dev_random(torch, (batch, 2 * n), 3)
    plans = [(m.mufft_create_plan_1d_c2c(n, -1, f), launches) for f, launches in ((0, 1), (FOUR_STEP_16K, 1), (TWO_KERNEL, 2))]
    outs = []
    for plan, launches in plans + plans[::-1]:
        y = torch.zeros_like(x)
        m.mufft_cuda_launch_count(1)
        assert m.mufft_execute_plan_1d_batch(plan, y.data_ptr(), x.data_ptr(), batch, 0, 0) == 0
        assert m.mufft_cuda_launch_count(1) == launches
        outs.append(torch.view_as_complex(y.view(batch, n, 2)))
    for plan, _ in plans:
        m.mufft_free_plan_1d(plan)
    ref = torch.fft.fft(torch.view_as_complex(x.view(batch, n, 2)), dim=1)
    for y in outs:
        assert (torch.linalg.vector_norm(y - ref) / torch.linalg.vector_norm(ref)).item() <= tol(n)


@pytest.mark.parametrize("where", ["pageable", "managed", "device"])
def test_buffers_with_4_byte_alignment_are_staged(m, torch, oracle, where):
    """The real side of R2C / C2R is a float array, which C aligns to 4 bytes only (the reference's
    scalar path takes it).  The kernels move 8-byte values, so such a buffer is never used in place:
    it takes the staged path whatever memory it lives in (same check on the emulated host layer:
    tests/test_hostemu.py)."""
    n, batch = 4096, 37
    r = np.random.default_rng(11).random((batch, n), np.float32) - np.float32(0.5)
    out_floats = batch * (n // 2 + 1) * 2
    if where == "pageable":
        raw_in = np.zeros(batch * n + 16, np.float32)
        raw_out = np.zeros(out_floats + 16, np.float32)
        base_in, base_out = raw_in.ctypes.data, raw_out.ctypes.data
    elif where == "managed":
        base_in, base_out = m.mufft_alloc((batch * n + 16) * 4), m.mufft_alloc((out_floats + 16) * 4)
        raw_in = np.ctypeslib.as_array((ctypes.c_float * (batch * n + 16)).from_address(base_in))
        raw_out = np.ctypeslib.as_array((ctypes.c_float * (out_floats + 16)).from_address(base_out))
    else:
        t_in = torch.zeros(batch * n + 16, dtype=torch.float32, device="cuda")
        t_out = torch.zeros(out_floats + 16, dtype=torch.float32, device="cuda")
        base_in, base_out = t_in.data_ptr(), t_out.data_ptr()
    o_in = ((-base_in) % 8) // 4 + 1   # in floats: the address is 4 mod 8
    o_out = ((-base_out) % 8) // 4 + 1
    if where == "device":
        t_in[o_in:o_in + batch * n] = torch.from_numpy(r.ravel()).cuda()
        torch.cuda.synchronize()
    else:
        raw_in[o_in:o_in + batch * n] = r.ravel()
    plan = m.mufft_create_plan_1d_r2c(n, 0)
    assert m.mufft_execute_plan_1d_batch(plan, base_out + 4 * o_out, base_in + 4 * o_in, batch, 0, 0) == 0, m.mufft_cuda_last_error()
    assert m.mufft_execute_plan_1d_batch_async(plan, base_out + 4 * o_out, base_in + 4 * o_in, batch, 0, 0) != 0  # cannot stage
    m.mufft_free_plan_1d(plan)
    if where == "device":
        got = t_out[o_out:o_out + out_floats].cpu().numpy()
    else:
        got = raw_out[o_out:o_out + out_floats].copy()
    got = got.view(np.complex64).reshape(batch, n // 2 + 1)
    for i in (0, 1, batch // 2, batch - 1):
        assert rel_l2(got[i], oracle.r2c(r[i])) <= tol(n)
    if where == "managed":
        m.mufft_free(base_in)
        m.mufft_free(base_out)


def test_second_device_after_the_first(m, torch):
    """Function attributes (the opt-in to more than 48 KB of shared memory) are per device: kernels that were first
    launched on device 0 must also run on device 1 of the same process (skipped with one GPU)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    results = []
    try:
        for dev in (0, 1):
            assert m.mufft_cuda_set_device(dev) == 0
            with torch.cuda.device(dev):
                x = torch.rand((5, 2 * 8192), dtype=torch.float32, device=f"cuda:{dev}") - 0.5
                y = torch.zeros_like(x)
                plan = m.mufft_create_plan_1d_c2c(8192, -1, 0)
                assert m.mufft_execute_plan_1d_batch(plan, y.data_ptr(), x.data_ptr(), 5, 0, 0) == 0, m.mufft_cuda_last_error()
                m.mufft_free_plan_1d(plan)
                ref = torch.fft.fft(torch.view_as_complex(x.view(5, 8192, 2)), dim=1)
                results.append((torch.linalg.vector_norm(torch.view_as_complex(y.view(5, 8192, 2)) - ref) / torch.linalg.vector_norm(ref)).item())
                img = torch.rand((3, 1024, 1024, 2), dtype=torch.float32, device=f"cuda:{dev}") - 0.5
                o2 = torch.zeros_like(img)
                p2 = m.mufft_create_plan_2d_c2c(1024, 1024, -1, 0)
                assert m.mufft_execute_plan_2d_batch(p2, o2.data_ptr(), img.data_ptr(), 3) == 0, m.mufft_cuda_last_error()
                m.mufft_free_plan_2d(p2)
                ref2 = torch.fft.fft2(torch.view_as_complex(img))
                results.append((torch.linalg.vector_norm(torch.view_as_complex(o2) - ref2) / torch.linalg.vector_norm(ref2)).item())
                # the batch sizes that select the warp-specialised TMA kernel (tensor map, 211 KB of shared memory), the
                # prefetching row kernel and the four-step kernel with its inputs through the TMA unit
                img = torch.rand((9, 1024, 1024, 2), dtype=torch.float32, device=f"cuda:{dev}") - 0.5
                o2 = torch.zeros_like(img)
                p2 = m.mufft_create_plan_2d_c2c(1024, 1024, -1, 0)
                assert m.mufft_execute_plan_2d_batch(p2, o2.data_ptr(), img.data_ptr(), 9) == 0, m.mufft_cuda_last_error()
                m.mufft_free_plan_2d(p2)
                ref2 = torch.fft.fft2(torch.view_as_complex(img))
                results.append((torch.linalg.vector_norm(torch.view_as_complex(o2) - ref2) / torch.linalg.vector_norm(ref2)).item())
                for n, batch in ((16384, 1100), (32768, 300)):
                    x = torch.rand((batch, 2 * n), dtype=torch.float32, device=f"cuda:{dev}") - 0.5
                    y = torch.zeros_like(x)
                    plan = m.mufft_create_plan_1d_c2c(n, -1, 0)
                    assert m.mufft_execute_plan_1d_batch(plan, y.data_ptr(), x.data_ptr(), batch, 0, 0) == 0, m.mufft_cuda_last_error()
                    m.mufft_free_plan_1d(plan)
                    ref = torch.fft.fft(torch.view_as_complex(x.view(batch, n, 2)), dim=1)
                    results.append((torch.linalg.vector_norm(torch.view_as_complex(y.view(batch, n, 2)) - ref) / torch.linalg.vector_norm(ref)).item())
    finally:
        m.mufft_cuda_set_device(0)
    assert max(results) <= tol(1024 * 1024)


def test_linearity_and_parseval_at_full_size(m, torch):
    """Size-independent properties at the BASELINE configs[1] shape (N=4096) on a 64 Ki batch:
    FFT(a x + y) = a FFT(x) + FFT(y), and sum |X|^2 = N sum |x|^2."""
    n, batch = 4096, 65536
    x, y = dev_random(torch, (batch, 2 * n), 1), dev_random(torch, (batch, 2 * n), 2)
    plan = m.mufft_create_plan_1d_c2c(n, -1, 0)
    fx, fy, fz = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    for dst, src in ((fx, x), (fy, y), (fz, 0.5 * x + y)):
        assert m.mufft_execute_plan_1d_batch(plan, dst.data_ptr(), src.data_ptr(), batch, 0, 0) == 0
    m.mufft_free_plan_1d(plan)
    lin = torch.linalg.vector_norm(fz - (0.5 * fx + fy)) / torch.linalg.vector_norm(fz)
    assert lin.item() <= tol(n)
    e_in = (x.double() ** 2).sum(dim=1)
    e_out = (fx.double() ** 2).sum(dim=1)
    assert torch.max(torch.abs(e_out / (n * e_in) - 1)).item() < 1e-5


@pytest.mark.parametrize("kind", ["c2c64", "r2c4096", "conv1024", "c2c65536"])
def test_repeated_small_calls_are_replayed_from_a_graph(m, torch, oracle, kind):
    """The latency path (BASELINE configs[0] as `muFFT-bench 1000000 64` runs it): a small synchronous fft.h call that
    comes again with the same pointers is captured into a CUDA graph and replayed.  The data changes between the
    calls, the pointers change in the middle of the run: results must follow, launch counts stay those of the path."""
    rng = np.random.default_rng(5)
    if kind == "conv1024":
        n = 1024
        method = 1 | 4 | 8
        plan = m.mufft_create_plan_conv(n, 0, method)
        blk = m.mufft_conv_get_transformed_block_size(plan)
        taps = (rng.random(n // 2, np.float32) - np.float32(0.5))
        bufs = [m.mufft_alloc(blk) for _ in range(3)] + [m.mufft_alloc(n * 8), m.mufft_alloc(n * 8)]
        fspec, spec, inp = bufs[0], bufs[1], bufs[3]
        view_in = np.ctypeslib.as_array((ctypes.c_float * n).from_address(inp))
        t_dev = m.mufft_alloc(n * 2)
        np.ctypeslib.as_array((ctypes.c_float * (n // 2)).from_address(t_dev))[:] = taps
        m.mufft_execute_conv_input(plan, 1, fspec, t_dev)
        for k in range(6):
            out = bufs[4] if k < 3 else bufs[2]  # blk = n * 8 bytes for the stereo method: both can hold an output
            a = (rng.random(n, np.float32) - np.float32(0.5))
            view_in[:] = a
            m.mufft_execute_conv_input(plan, 0, spec, inp)
            m.mufft_execute_conv_output(plan, out, spec, fspec)
            got = np.ctypeslib.as_array((ctypes.c_float * (2 * n)).from_address(out)).view(np.complex64).copy()
            want = oracle.conv(a.view(np.complex64), taps, n, method)[0]
            assert rel_l2(got, want) <= 2 * tol(n), k
        m.mufft_free_plan_conv(plan)
        for b in bufs + [t_dev]:
            m.mufft_free(b)
        return
    n = int(kind[3:])
    real = kind.startswith("r2c")
    plan = m.mufft_create_plan_1d_r2c(n, 0) if real else m.mufft_create_plan_1d_c2c(n, -1, 0)
    in_floats, out_floats = (n, 2 * (n // 2 + 1)) if real else (2 * n, 2 * n)
    ins = [m.mufft_alloc(in_floats * 4) for _ in range(2)]
    outs = [m.mufft_alloc(out_floats * 4) for _ in range(2)]
    dev_in = torch.zeros(in_floats, dtype=torch.float32, device="cuda")
    dev_out = torch.zeros(out_floats, dtype=torch.float32, device="cuda")
    for k in range(9):
        x = rng.random(in_floats, np.float32) - np.float32(0.5)
        want = oracle.r2c(x) if real else oracle.c2c(x.view(np.complex64), -1)
        if k < 6:
            pi, po = ins[k // 3], outs[k // 3]  # three calls per pointer pair: ordinary, captured, replayed
            np.ctypeslib.as_array((ctypes.c_float * in_floats).from_address(pi))[:] = x
        else:
            dev_in.copy_(torch.from_numpy(x))
            torch.cuda.synchronize()
            pi, po = dev_in.data_ptr(), dev_out.data_ptr()
        m.mufft_cuda_launch_count(1)
        m.mufft_execute_plan_1d(plan, po, pi)
        assert m.mufft_cuda_launch_count(1) == (1 if n <= 65536 and not (real and n > 32768) else 2), k
        got = dev_out.cpu().numpy() if k >= 6 else np.ctypeslib.as_array((ctypes.c_float * out_floats).from_address(po)).copy()
        assert rel_l2(got.view(np.complex64), want) <= tol(n), k
    m.mufft_free_plan_1d(plan)
    for b in ins + outs:
        m.mufft_free(b)


@pytest.mark.parametrize("n,width", [(4096, 37), (16384, 256), (32768, 1000), (32768, 4096), (65536, 64)])
@pytest.mark.parametrize("path", ["default", "two_kernel"])
def test_long_columns(m, torch, oracle, n, width, path):
    """mufft_execute_plan_1d_columns beyond 2048 rows (the column pass of a slab transform): one persistent launch of
    the four-step column kernel for 16384 / 32768 rows - blocks of 32 columns, the last one partly masked, intermediate in
    L2 - or (plan flag, and beyond 32768 rows) two sweeps of the column kernel through scratch; in place and out of place."""
    if path == "two_kernel" and width > 1000:
        pytest.skip("the fallback is covered at the smaller widths")
    flags = TWO_KERNEL if path == "two_kernel" else 0
    pitch = width + 5
    x = dev_random(torch, (n, pitch, 2), n + width)
    xc = torch.view_as_complex(x)[:, :width]
    ref = torch.fft.fft(xc.to(torch.complex128), dim=0)
    plan = m.mufft_create_plan_1d_c2c(n, -1, flags)
    assert plan
    y = torch.full_like(x, float("nan"))
    m.mufft_cuda_launch_count(1)
    assert m.mufft_execute_plan_1d_columns(plan, y.data_ptr(), x.data_ptr(), width, pitch * 8, pitch * 8) == 0, m.mufft_cuda_last_error()
    torch.cuda.synchronize()
    one_launch = path == "default" and n in (16384, 32768)
    assert (m.mufft_cuda_launch_count(1) == 1) == one_launch
    yc = torch.view_as_complex(y)
    assert (torch.linalg.vector_norm(yc[:, :width] - ref) / torch.linalg.vector_norm(ref)).item() <= tol(n)
    assert torch.isnan(yc[:, width:].real).all(), "nothing is written beyond the last column"
    col = xc[:, width - 1].contiguous().cpu().numpy()
    assert rel_l2(yc[:, width - 1].cpu().numpy(), oracle.c2c(col, -1)) <= tol(n)
    # in place, twice on the same plan (ring and counters are reused): FFT of the FFT = the reversed input times n
    assert m.mufft_execute_plan_1d_columns(plan, y.data_ptr(), y.data_ptr(), width, pitch * 8, pitch * 8) == 0
    torch.cuda.synchronize()
    m.mufft_free_plan_1d(plan)
    back = torch.roll(torch.flip(yc[:, :width], dims=[0]), 1, dims=0) / n
    assert (torch.linalg.vector_norm(back - xc) / torch.linalg.vector_norm(xc)).item() <= 2 * tol(n)


def test_library_allocators_and_copy_helpers(m, torch, oracle):
    """mufft_cuda_host_alloc (pinned), mufft_cuda_device_alloc, mufft_cuda_copy / _memset: a C caller's way to keep data
    on the device or in fast staging memory without any other CUDA binding."""
    n, batch = 1024, 300
    nbytes = batch * n * 8
    hin, hout = m.mufft_cuda_host_alloc(nbytes), m.mufft_cuda_host_alloc(nbytes)
    din, dout = m.mufft_cuda_device_alloc(nbytes), m.mufft_cuda_device_alloc(nbytes)
    assert hin and hout and din and dout, m.mufft_cuda_last_error()
    vin = np.ctypeslib.as_array((ctypes.c_float * (2 * n * batch)).from_address(hin))
    vout = np.ctypeslib.as_array((ctypes.c_float * (2 * n * batch)).from_address(hout))
    vin[:] = np.random.default_rng(1).random(2 * n * batch, np.float32) - np.float32(0.5)
    plan = m.mufft_create_plan_1d_c2c(n, -1, 0)
    # pinned host buffers through the staged synchronous call
    assert m.mufft_execute_plan_1d_batch(plan, hout, hin, batch, 0, 0) == 0, m.mufft_cuda_last_error()
    staged = vout.copy()
    # the same through device memory the caller manages itself
    assert m.mufft_cuda_memset(dout, 0xff, nbytes) == 0
    assert m.mufft_cuda_copy(din, hin, nbytes) == 0
    assert m.mufft_execute_plan_1d_batch_async(plan, dout, din, batch, 0, 0) == 0, m.mufft_cuda_last_error()
    vout[:] = 0
    assert m.mufft_cuda_copy(hout, dout, nbytes) == 0  # ordered after the transform on the same stream, complete on return
    assert np.array_equal(vout, staged)
    for b in (0, batch - 1):
        want = oracle.c2c(vin[b * 2 * n:(b + 1) * 2 * n].view(np.complex64).copy(), -1)
        assert rel_l2(vout[b * 2 * n:(b + 1) * 2 * n].view(np.complex64), want) <= tol(n)
    m.mufft_free_plan_1d(plan)
    m.mufft_cuda_host_free(hin); m.mufft_cuda_host_free(hout)
    m.mufft_cuda_device_free(din); m.mufft_cuda_device_free(dout)


@pytest.mark.parametrize("nx,ny", [(1024, 1024), (512, 1024), (1024, 512)])
def test_c2c_2d_tma_kernel_and_its_fallbacks(m, torch, nx, ny):
    """Batches of 8 images or more of these shapes run on the warp-specialised TMA kernel (fft2d_tma_kernel.cuh);
    smaller batches and output buffers that are not 128-byte aligned stay on the ticket kernel.  Every image of an odd
    batch, both directions, against cuFFT; the two kernels against each other."""
    batch = 13
    x = dev_random(torch, (batch, ny, nx, 2), 3 * nx + ny)
    spare = torch.zeros((batch * ny * nx + 1, 2), dtype=torch.float32, device="cuda")
    y = torch.zeros_like(x)
    y_off = spare[1:].view(batch, ny, nx, 2)  # 8 bytes off the allocation: not a TMA destination
    assert y.data_ptr() % 128 == 0 and y_off.data_ptr() % 128 == 8
    for direction in (-1, 1):
        plan = m.mufft_create_plan_2d_c2c(nx, ny, direction, 0)
        assert plan
        xc = torch.view_as_complex(x)
        ref = torch.fft.fft2(xc) if direction < 0 else torch.fft.ifft2(xc, norm="forward")
        for out, count in ((y, batch), (y_off, batch), (y, 7)):
            out.zero_()
            m.mufft_cuda_launch_count(1)
            assert m.mufft_execute_plan_2d_batch(plan, out.data_ptr(), x.data_ptr(), count) == 0, m.mufft_cuda_last_error()
            assert m.mufft_cuda_launch_count(1) == 1
            oc = torch.view_as_complex(out)[:count]
            err = (torch.linalg.vector_norm((oc - ref[:count]).reshape(count, -1), dim=1) /
                   torch.linalg.vector_norm(ref[:count].reshape(count, -1), dim=1)).max().item()
            assert err <= tol(nx * ny)
        # in place, on the TMA kernel
        z = x.clone()
        assert m.mufft_execute_plan_2d_batch(plan, z.data_ptr(), z.data_ptr(), batch) == 0, m.mufft_cuda_last_error()
        zc = torch.view_as_complex(z)
        err = (torch.linalg.vector_norm((zc - ref).reshape(batch, -1), dim=1) / torch.linalg.vector_norm(ref.reshape(batch, -1), dim=1)).max().item()
        assert err <= tol(nx * ny)
        m.mufft_free_plan_2d(plan)


@pytest.mark.parametrize("n,batch", [(2048, 1023), (2048, 5000), (8192, 1023), (8192, 1024), (8192, 2500), (16384, 1023), (16384, 1400)])
def test_c2c_prefetch_rows(m, torch, oracle, n, batch):
    """2048 / 8192 / 16384-point C2C: batches of 1024 transforms or more run on the persistent row kernel that lets the TMA unit
    fetch the next transform into the exchange area (rows_prefetch_kernel.cuh), smaller ones on the row kernel; same
    results, checked on every transform, out of place and in place."""
    x = dev_random(torch, (batch, 2 * n), batch)
    for direction in (-1, 1):
        y = torch.zeros_like(x)
        plan = m.mufft_create_plan_1d_c2c(n, direction, 0)
        assert plan
        m.mufft_cuda_launch_count(1)
        assert m.mufft_execute_plan_1d_batch(plan, y.data_ptr(), x.data_ptr(), batch, 0, 0) == 0, m.mufft_cuda_last_error()
        assert m.mufft_cuda_launch_count(1) == 1
        xc = torch.view_as_complex(x.view(batch, n, 2))
        ref = torch.fft.fft(xc, dim=1) if direction < 0 else torch.fft.ifft(xc, dim=1, norm="forward")
        yc = torch.view_as_complex(y.view(batch, n, 2))
        err = (torch.linalg.vector_norm(yc - ref, dim=1) / torch.linalg.vector_norm(ref, dim=1)).max().item()
        assert err <= tol(n), f"worst transform rel L2 {err:.3e}"
        for i in (0, batch - 1):
            assert rel_l2(y[i].cpu().numpy().view(np.complex64), oracle.c2c(x[i].cpu().numpy().view(np.complex64), direction)) <= tol(n)
        z = x.clone()
        assert m.mufft_execute_plan_1d_batch(plan, z.data_ptr(), z.data_ptr(), batch, 0, 0) == 0
        assert torch.equal(z, y)
        m.mufft_free_plan_1d(plan)


def test_r2c_2d_tma_kernel(m, torch, oracle):
    """Real images 1024 wide and 1024 high, 8 or more of them: rows (with the resolve) and columns on the warp-specialised
    TMA kernel; the last column tile holds ONE column that carries data (bin 512) and seven the tensor map clips.
    Every image against cuFFT, one against the oracle, the padding columns of the output untouched; a batch of 7 (ticket
    kernel) gives the same images."""
    nx, ny, batch = 1024, 1024, 11
    g = torch.Generator(device="cuda")
    g.manual_seed(77)
    x = torch.rand((batch, ny, nx), dtype=torch.float32, device="cuda", generator=g) - 0.5
    out = torch.full((batch, ny, nx, 2), 7.0, dtype=torch.float32, device="cuda")
    plan = m.mufft_create_plan_2d_r2c(nx, ny, 0)
    assert plan
    m.mufft_cuda_launch_count(1)
    assert m.mufft_execute_plan_2d_batch(plan, out.data_ptr(), x.data_ptr(), batch) == 0, m.mufft_cuda_last_error()
    assert m.mufft_cuda_launch_count(1) == 1
    width = nx // 2 + 1
    got = torch.view_as_complex(out)
    ref = torch.fft.fft2(x)[:, :, :width]
    err = (torch.linalg.vector_norm((got[:, :, :width] - ref).reshape(batch, -1), dim=1) / torch.linalg.vector_norm(ref.reshape(batch, -1), dim=1)).max().item()
    assert err <= tol(nx * ny)
    assert bool((out[:, :, width:, :] == 7.0).all()), "columns beyond nx/2 + 1 are padding: never written"
    want = oracle.r2c_2d(x[3].cpu().numpy(), 0)[:, :width]
    assert rel_l2(got[3, :, :width].cpu().numpy(), want) <= tol(nx * ny)
    small = torch.zeros((7, ny, nx, 2), dtype=torch.float32, device="cuda")
    assert m.mufft_execute_plan_2d_batch(plan, small.data_ptr(), x.data_ptr(), 7) == 0, m.mufft_cuda_last_error()
    ds = torch.view_as_complex(small)[:, :, :width] - got[:7, :, :width]
    assert (torch.linalg.vector_norm(ds) / torch.linalg.vector_norm(ref[:7])).item() <= tol(nx * ny)
    m.mufft_free_plan_2d(plan)


@pytest.mark.parametrize("n", [16384, 32768])
@pytest.mark.parametrize("flags", [0, FLAG_FULL_R2C, FLAG_ZERO_PAD])
def test_r2c_large_batch_prefetch_rows(m, torch, oracle, n, flags):
    """R2C of 16384 / 32768 reals in batches of 1024 or more: the persistent prefetching row kernel with the resolve in the
    registers of the mirrored last pass (rows_prefetch_kernel.cuh).  Every transform against cuFFT, a sample against the
    oracle, a batch of 1023 (row kernel) gives the same bits."""
    batch = 1100
    n_in = n // 2 if flags & FLAG_ZERO_PAD else n
    bins = n if flags & FLAG_FULL_R2C else n // 2 + 1
    pitch = (bins + 3) // 2 * 2
    x = dev_random(torch, (batch, n_in), n + flags)
    spec = torch.zeros((batch, 2 * pitch), dtype=torch.float32, device="cuda")
    fwd = m.mufft_create_plan_1d_r2c(n, flags)
    assert fwd
    m.mufft_cuda_launch_count(1)
    assert m.mufft_execute_plan_1d_batch(fwd, spec.data_ptr(), x.data_ptr(), batch, pitch * 8, 0) == 0, m.mufft_cuda_last_error()
    assert m.mufft_cuda_launch_count(1) == 1
    xp = torch.nn.functional.pad(x, (0, n - n_in))
    ref = torch.fft.fft(xp, dim=1)[:, :bins]
    got = torch.view_as_complex(spec.view(batch, pitch, 2))[:, :bins]
    err = (torch.linalg.vector_norm(got - ref, dim=1) / torch.linalg.vector_norm(ref, dim=1)).max().item()
    assert err <= tol(n), f"worst transform rel L2 {err:.3e}"
    for i in (0, batch - 1):
        assert rel_l2(got[i].cpu().numpy(), oracle.r2c(x[i].cpu().numpy(), flags, n=n)) <= tol(n)
    small = torch.zeros((1023, 2 * pitch), dtype=torch.float32, device="cuda")
    assert m.mufft_execute_plan_1d_batch(fwd, small.data_ptr(), x.data_ptr(), 1023, pitch * 8, 0) == 0
    assert torch.equal(small, spec[:1023])
    m.mufft_free_plan_1d(fwd)


@pytest.mark.parametrize("n", [4096, 16384, 32768])
@pytest.mark.parametrize("pitch", ["tight", "aligned"])
def test_c2r_large_batch_prefetch_rows(m, torch, oracle, n, pitch):
    """C2R of 4096 / 16384 / 32768 reals in batches of 1024 or more on the persistent prefetching row kernel: spectrum rows of
    n/2 + 1 bins, tightly packed (every other row starts 8 bytes off a 16-byte boundary: the bulk copy starts one bin
    early) or at an even pitch.  Every transform against the real input it came from, a sample against the oracle, a batch
    of 1023 (row kernel) gives the same bits."""
    batch = 1100
    bins = n // 2 + 1
    sp = bins if pitch == "tight" else bins + 1
    x = dev_random(torch, (batch, n), n + 5)
    spec = torch.zeros((batch, sp), dtype=torch.complex64, device="cuda")
    spec[:, :bins] = torch.fft.rfft(x.double(), dim=1).to(torch.complex64)
    back = torch.zeros((batch, n), dtype=torch.float32, device="cuda")
    inv = m.mufft_create_plan_1d_c2r(n, 0)
    assert inv
    m.mufft_cuda_launch_count(1)
    assert m.mufft_execute_plan_1d_batch(inv, back.data_ptr(), torch.view_as_real(spec).data_ptr(), batch, 0, sp * 8) == 0, m.mufft_cuda_last_error()
    assert m.mufft_cuda_launch_count(1) == 1
    err = (torch.linalg.vector_norm(back / n - x, dim=1) / torch.linalg.vector_norm(x, dim=1)).max().item()
    assert err <= 2 * tol(n), f"worst transform rel L2 {err:.3e}"
    ss = torch.view_as_real(spec).cpu().numpy().view(np.complex64).reshape(batch, sp)
    for i in (0, 1, batch - 1):
        assert rel_l2(back[i].cpu().numpy(), oracle.c2r(ss[i, :bins].copy(), n)) <= tol(n)
    small = torch.zeros((1023, n), dtype=torch.float32, device="cuda")
    assert m.mufft_execute_plan_1d_batch(inv, small.data_ptr(), torch.view_as_real(spec).data_ptr(), 1023, 0, sp * 8) == 0
    assert torch.equal(small, back[:1023])
    m.mufft_free_plan_1d(inv)
