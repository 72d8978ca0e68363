"""Slab-decomposed 2D transform on the GPU(s): local passes through the CUDA library, exchange
over NCCL or the row kernel's peer stores.  With one visible GPU the group has one rank (the pack /
peer-scatter / column-slab paths still run);
with two or more the test spawns one process per GPU (up to 4)."""
import math
import os

import numpy as np
import pytest

from cpu_libs import rel_l2

pytestmark = pytest.mark.gpu


def tol(n):
    return 1e-6 * math.log2(n)


def _free_port():
    """A rendezvous token for one process group: a fresh file for torch's FileStore.  (A TCP port picked here and bound
    later by rank 0 can be taken in between - seen once on a GPU box as EADDRINUSE, which with `pytest -x` ends the run.)"""
    import tempfile
    fd, path = tempfile.mkstemp(prefix="mufft_rdzv_")
    os.close(fd)
    os.unlink(path)  # FileStore creates it
    return path


def _worker(rank, world, port, nx, ny, natural, direction, exchange, result):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="file://" + port, rank=rank, world_size=world, device_id=dev)
    try:
        import mufft_b200
        from mufft_b200.slab import SlabFFT2D
        assert mufft_b200.mufft_cuda_set_device(rank) == 0
        g = torch.Generator(device="cpu").manual_seed(7)
        full = torch.complex(torch.rand(ny, nx, generator=g) - 0.5, torch.rand(ny, nx, generator=g) - 0.5)
        rl, cl = ny // world, nx // world
        # with one rank the p2p path still runs the row kernel's peer-scatter epilogue (one "peer": itself)
        plan = SlabFFT2D(nx, ny, direction, device=dev, chunks=4, exchange=exchange)
        mine = full[rank * rl:(rank + 1) * rl].contiguous().to(dev)
        out = torch.empty((rl, nx) if natural else (ny, cl), dtype=torch.complex64, device=dev)
        buf = plan.buffers()
        for _ in range(2):  # buffers are reusable
            plan.execute(out, mine, buf=buf, natural=natural)
        torch.cuda.synchronize()
        fulld = full.to(dev).to(torch.complex128)
        ref = torch.fft.fft2(fulld) if direction < 0 else torch.fft.ifft2(fulld, norm="forward")
        want = ref[rank * rl:(rank + 1) * rl] if natural else ref[:, rank * cl:(rank + 1) * cl]
        result[rank] = (torch.linalg.vector_norm(out.to(torch.complex128) - want) / torch.linalg.vector_norm(want)).item()
        plan.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nx,ny", [(256, 128), (2048, 4096), (16384, 512), (1024, 32768)])
@pytest.mark.parametrize("natural", [False, True])
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_slab_2d(nx, ny, natural, exchange):
    import torch
    import torch.multiprocessing as mp
    world = min(4, torch.cuda.device_count())
    while world & (world - 1):
        world -= 1
    assert world >= 1, "no CUDA device visible"
    port = _free_port()
    ctx = mp.get_context("spawn")
    result = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, ny, natural, -1, exchange, result)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    tol = 1e-6 * math.log2(nx * ny)
    assert len(result) == world and max(result.values()) <= tol, dict(result)


def _rworker(rank, world, port, nx, ny, result):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="file://" + port, rank=rank, world_size=world, device_id=dev)
    try:
        import mufft_b200
        from mufft_b200.slab import SlabRFFT2D
        assert mufft_b200.mufft_cuda_set_device(rank) == 0
        g = torch.Generator(device="cpu").manual_seed(11)
        full = torch.rand(ny, nx, generator=g) - 0.5
        rl = ny // world
        plan = SlabRFFT2D(nx, ny, device=dev, chunks=4)
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


@pytest.mark.parametrize("nx,ny", [(256, 128), (4096, 2048), (32768, 512), (1024, 16384)])
def test_slab_2d_real(nx, ny):
    """R2C forward + C2R inverse of one large real image over the visible GPUs (1, 2 or 4 ranks):
    half spectrum against float64 rfft2, and inverse(forward(x)) = Nx*Ny*x."""
    import torch
    import torch.multiprocessing as mp
    world = min(4, torch.cuda.device_count())
    while world & (world - 1):
        world -= 1
    assert world >= 1, "no CUDA device visible"
    port = _free_port()
    ctx = mp.get_context("spawn")
    result = ctx.Manager().dict()
    procs = [ctx.Process(target=_rworker, args=(r, world, port, nx, ny, result)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    tol = 1e-6 * math.log2(nx * ny)
    assert len(result) == world and max(result.values()) <= tol, dict(result)


def _pow2_devices(torch, cap=8):
    world = min(cap, torch.cuda.device_count())
    while world & (world - 1):
        world -= 1
    return world


@pytest.mark.parametrize("kind,nx,ny", [("c2c", 4096, 4096), ("c2c", 32768, 256), ("c2c", 256, 128), ("r2c", 4096, 2048), ("r2c", 32768, 64)])
def test_slab_plan_in_c_against_the_oracle(oracle, kind, nx, ny):
    """include/mufft_cuda.h: mufft_cuda_create_plan_2d_slab_* - the slab transform behind the C ABI, one process
    driving every visible GPU (a power of two of them; a single GPU still runs the scatter row kernel, the
    events and the in-place column pass).  Compared with the ORACLE's mufft_execute_plan_2d on the whole array
    (4096 x 4096: SURVEY.md section 8c), not with a GPU library."""
    import ctypes as C

    import torch

    import mufft_b200 as m
    ndev = _pow2_devices(torch)
    rng = np.random.default_rng(nx + ny)
    if kind == "c2c":
        x = (rng.random((ny, nx, 2), np.float32) - np.float32(0.5)).view(np.complex64).reshape(ny, nx)
        plan = m.mufft_cuda_create_plan_2d_slab_c2c(nx, ny, -1, ndev, None, 0)
        out = np.full((ny, nx), np.nan, np.complex64)
        want = oracle.c2c_2d(x, -1)
    else:
        x = rng.random((ny, nx), np.float32) - np.float32(0.5)
        plan = m.mufft_cuda_create_plan_2d_slab_r2c(nx, ny, ndev, None, 0)
        out = np.full((ny, nx // 2 + 1), np.nan, np.complex64)
        want = oracle.r2c_2d(x)[:, : nx // 2 + 1]
    assert plan, m.mufft_cuda_last_error()
    assert m.mufft_cuda_slab_devices(plan) == ndev
    first = C.c_uint(0)
    assert sum(m.mufft_cuda_slab_output_columns(plan, g, C.byref(first)) for g in range(ndev)) == out.shape[1]
    for _ in range(2):  # the plan is reusable
        out[...] = np.nan
        assert m.mufft_cuda_execute_plan_2d_slab_host(plan, out.ctypes.data, x.ctypes.data) == 0, m.mufft_cuda_last_error()
    assert m.mufft_cuda_slab_last_ms(plan) > 0.0
    m.mufft_cuda_free_plan_2d_slab(plan)
    assert rel_l2(out, want) <= tol(nx * ny)
