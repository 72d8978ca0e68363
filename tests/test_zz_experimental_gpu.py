"""GPU tests of paths that are validated on the CPU emulation (tests/test_emu.py) but have not run
on hardware yet.  They are skipped unless MUFFT_TEST_EXPERIMENTAL=1, so that the suite the driver
runs only contains paths with hardware evidence; tools/probe_optin.py runs the same checks without
torch in a few seconds under gpurun.  Move a test into the regular files once its path has passed here.
"""
import math
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MUFFT_TEST_EXPERIMENTAL") != "1", reason="set MUFFT_TEST_EXPERIMENTAL=1 to run")]


def tol(n):
    return 1e-6 * math.log2(n)


@pytest.fixture(scope="module")
def m():
    import mufft_b200
    assert mufft_b200.mufft_cuda_device_count() >= 1
    return mufft_b200


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


@pytest.mark.parametrize("n", [32768, 65536])
def test_cluster_rows(m, torch, oracle, n):
    """mufft_cuda_set_cluster_rows(1): 32768 / 65536-point rows in one kernel on a cluster of 2 / 4 CTAs."""
    batch = 301
    g = torch.Generator(device="cuda")
    g.manual_seed(n)
    x = torch.rand((batch, 2 * n), dtype=torch.float32, device="cuda", generator=g) - 0.5
    xc = torch.view_as_complex(x.view(batch, n, 2))
    m.mufft_cuda_set_cluster_rows(1)
    try:
        for direction in (-1, 1):
            plan = m.mufft_create_plan_1d_c2c(n, direction, 0)
            y = torch.zeros_like(x)
            m.mufft_cuda_launch_count(1)
            assert m.mufft_execute_plan_1d_batch(plan, y.data_ptr(), x.data_ptr(), batch, 0, 0) == 0, m.mufft_cuda_last_error()
            assert m.mufft_cuda_launch_count(1) == 1
            m.mufft_free_plan_1d(plan)
            ref = torch.fft.fft(xc, dim=1) if direction < 0 else torch.fft.ifft(xc, dim=1, norm="forward")
            yc = torch.view_as_complex(y.view(batch, n, 2))
            err = (torch.linalg.vector_norm(yc - ref, dim=1) / torch.linalg.vector_norm(ref, dim=1)).max().item()
            assert err <= tol(n)
            got = yc[batch - 1].cpu().numpy()
            want = oracle.c2c(xc[batch - 1].cpu().numpy(), direction)
            assert float(np.linalg.norm(got - want) / np.linalg.norm(want)) <= tol(n)
        # zero-padded forward transform
        plan = m.mufft_create_plan_1d_c2c(n, -1, m.MUFFT_FLAG_ZERO_PAD_UPPER_HALF)
        xh = x[:, :n].contiguous()
        y = torch.zeros_like(x)
        assert m.mufft_execute_plan_1d_batch(plan, y.data_ptr(), xh.data_ptr(), batch, 0, 0) == 0
        m.mufft_free_plan_1d(plan)
        xp = torch.zeros((batch, n), dtype=torch.complex64, device="cuda")
        xp[:, : n // 2] = torch.view_as_complex(xh.view(batch, n // 2, 2))
        ref = torch.fft.fft(xp, dim=1)
        yc = torch.view_as_complex(y.view(batch, n, 2))
        assert (torch.linalg.vector_norm(yc - ref) / torch.linalg.vector_norm(ref)).item() <= tol(n)
    finally:
        m.mufft_cuda_set_cluster_rows(0)


@pytest.mark.parametrize("nx,ny", [(512, 256), (1024, 1024), (2048, 512), (4096, 256)])
@pytest.mark.parametrize("flags", [0, 1 << 16])
def test_fused_2d_real(m, torch, oracle, nx, ny, flags):
    """mufft_cuda_set_fused_2d_real(1): 2D R2C images (nx reals per row) in the fused persistent kernel."""
    batch = max(3, min(24, (64 << 20) // (8 * nx * ny)))
    g = torch.Generator(device="cuda")
    g.manual_seed(nx + ny)
    x = torch.rand((batch, ny, nx), dtype=torch.float32, device="cuda", generator=g) - 0.5
    out = torch.zeros((batch, ny, nx, 2), dtype=torch.float32, device="cuda")
    m.mufft_cuda_set_fused_2d_real(1)
    try:
        plan = m.mufft_create_plan_2d_r2c(nx, ny, flags)
        m.mufft_cuda_launch_count(1)
        assert m.mufft_execute_plan_2d_batch(plan, out.data_ptr(), x.data_ptr(), batch) == 0, m.mufft_cuda_last_error()
        assert m.mufft_cuda_launch_count(1) == 1
        m.mufft_free_plan_2d(plan)
    finally:
        m.mufft_cuda_set_fused_2d_real(0)
    width = nx if flags else nx // 2 + 1
    got = torch.view_as_complex(out)[:, :, :width]
    ref = torch.fft.fft2(x)[:, :, :width]
    err = (torch.linalg.vector_norm((got - ref).reshape(batch, -1), dim=1) / torch.linalg.vector_norm(ref.reshape(batch, -1), dim=1)).max().item()
    assert err <= tol(nx * ny)
    want = oracle.r2c_2d(x[batch - 1].cpu().numpy(), flags)[:, :width]
    g1 = got[batch - 1].cpu().numpy()
    assert float(np.linalg.norm(g1 - want) / np.linalg.norm(want)) <= tol(nx * ny)


def _rworker_p2p(rank, world, port, nx, ny, result):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import mufft_b200
        from mufft_b200.slab import SlabRFFT2D
        assert mufft_b200.mufft_cuda_set_device(rank) == 0
        g = torch.Generator(device="cpu").manual_seed(13)
        full = torch.rand(ny, nx, generator=g) - 0.5
        rl = ny // world
        plan = SlabRFFT2D(nx, ny, device=dev, exchange="p2p")
        mine = full[rank * rl:(rank + 1) * rl].contiguous().to(dev)
        spec = torch.empty((ny, plan.cols_local), dtype=torch.complex64, device=dev)
        back = torch.empty((rl, nx), dtype=torch.float32, device=dev)
        buf = plan.buffers()
        for _ in range(2):
            plan.forward(spec, mine, buf=buf)
            plan.inverse(back, spec, buf=buf)
        torch.cuda.synchronize()
        ref = torch.fft.rfft2(full.to(dev).to(torch.float64))
        lo = rank * plan.cols_even
        want = ref[:, lo:lo + plan.cols_local]
        e1 = (torch.linalg.vector_norm(spec.to(torch.complex128) - want) / torch.linalg.vector_norm(want)).item()
        e2 = (torch.linalg.vector_norm(back.double() / (nx * ny) - mine.double()) / torch.linalg.vector_norm(mine.double())).item()
        result[rank] = max(e1, e2)
        plan.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nx,ny", [(256, 256), (512, 2048), (1024, 1024), (2048, 256)])
@pytest.mark.parametrize("mode", [9, 10, 11, 12, 13, 14, 15, 16])
def test_c2c_2d_fused_flavours_not_measured_yet(m, torch, oracle, nx, ny, mode):
    """mufft_cuda_set_fused_2d 9 / 10 (tickets drawn two tiles ahead), 11 (three CTAs per SM), 12 / 13 (alternating tickets): same checks as
    tests/test_batch_gpu.py::test_c2c_2d_fused_flavours runs on the measured flavours."""
    from cpu_libs import rel_l2
    m.mufft_cuda_set_fused_2d(mode)
    try:
        batch = max(5, min(40, (96 << 20) // (8 * nx * ny)))
        g = torch.Generator(device="cuda")
        g.manual_seed(nx + ny + mode)
        x = torch.rand((batch, ny, nx, 2), dtype=torch.float32, device="cuda", generator=g) - 0.5
        y = torch.zeros_like(x)
        xc = torch.view_as_complex(x)
        for direction in (-1, 1):
            plan = m.mufft_create_plan_2d_c2c(nx, ny, direction, 0)
            assert plan
            m.mufft_cuda_launch_count(1)
            assert m.mufft_execute_plan_2d_batch(plan, y.data_ptr(), x.data_ptr(), batch) == 0, m.mufft_cuda_last_error()
            assert m.mufft_cuda_launch_count(1) == 1
            m.mufft_free_plan_2d(plan)
            ref = torch.fft.fft2(xc) if direction < 0 else torch.fft.ifft2(xc, norm="forward")
            yc = torch.view_as_complex(y)
            err = (torch.linalg.vector_norm((yc - ref).reshape(batch, -1), dim=1) /
                   torch.linalg.vector_norm(ref.reshape(batch, -1), dim=1)).max().item()
            assert err <= tol(nx * ny)
            assert rel_l2(yc[batch - 1].cpu().numpy(), oracle.c2c_2d(xc[batch - 1].cpu().numpy(), direction)) <= tol(nx * ny)
    finally:
        m.mufft_cuda_set_fused_2d(1)


@pytest.mark.parametrize("where", ["pageable", "managed", "device"])
def test_buffers_with_4_byte_alignment_are_staged(m, torch, oracle, where):
    """The real side of R2C / C2R is a float array, which C aligns to 4 bytes only (the reference's
    scalar path takes it).  The kernels move 8-byte values, so such a buffer is never used in place:
    it takes the staged path whatever memory it lives in (same check on the emulated host layer:
    tests/test_hostemu.py)."""
    import ctypes
    from cpu_libs import rel_l2
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


@pytest.mark.parametrize("nx,ny", [(256, 128), (4096, 2048), (32768, 512)])
def test_slab_2d_real_p2p(nx, ny):
    """SlabRFFT2D(exchange="p2p"): the R2C row kernel's epilogue scatters the half spectrum into the column slabs
    (one rank with one GPU - the scatter still runs, every bin to the own slab -, one process per GPU otherwise)."""
    import socket

    import torch
    import torch.multiprocessing as mp
    world = min(4, torch.cuda.device_count())
    while world & (world - 1):
        world -= 1
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    ctx = mp.get_context("spawn")
    result = ctx.Manager().dict()
    procs = [ctx.Process(target=_rworker_p2p, args=(r, world, port, nx, ny, result)) for r in range(world)]
    [p_.start() for p_ in procs]
    [p_.join(300) for p_ in procs]
    assert all(p_.exitcode == 0 for p_ in procs), [p_.exitcode for p_ in procs]
    assert len(result) == world and max(result.values()) <= tol(nx * ny), dict(result)
