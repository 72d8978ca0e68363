"""CPU test (gloo, world_size 2 and 4) of the slab decomposition's exchange logic
(mufft_b200/slab.py): partitioning, packing, the all-to-all placement and both output layouts.  The
local row/column transforms are injected (torch CPU FFT here); on the GPU box the same class runs
them through the CUDA library (tests/test_slab_gpu.py)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    """A rendezvous token for one process group: a fresh file for torch's FileStore.  (A TCP port picked here and bound
    later by rank 0 can be taken in between - seen once on a GPU box as EADDRINUSE, which with `pytest -x` ends the run.)"""
    import tempfile
    fd, path = tempfile.mkstemp(prefix="mufft_rdzv_")
    os.close(fd)
    os.unlink(path)  # FileStore creates it
    return path


def _worker(rank, world, port, nx, ny, chunks, natural, direction, result):
    dist.init_process_group("gloo", init_method="file://" + port, rank=rank, world_size=world)
    try:
        from mufft_b200.slab import SlabFFT2D
        g = torch.Generator().manual_seed(1234)
        full = torch.complex(torch.rand(ny, nx, generator=g) - 0.5, torch.rand(ny, nx, generator=g) - 0.5)
        fft = (lambda x, dim: torch.fft.fft(x, dim=dim)) if direction < 0 else (lambda x, dim: torch.fft.ifft(x, dim=dim, norm="forward"))

        def rows(out, inp):
            out.copy_(fft(inp, 1))

        def cols(out, inp):
            out.copy_(fft(inp, 0))
        plan = SlabFFT2D(nx, ny, direction, chunks=chunks, local_rows=rows, local_cols=cols)
        rl, cl = ny // world, nx // world
        mine = full[rank * rl:(rank + 1) * rl].contiguous()
        out = torch.empty((rl, nx) if natural else (ny, cl), dtype=torch.complex64)
        plan.execute(out, mine, natural=natural)
        ref = fft(fft(full, 1), 0)
        want = ref[rank * rl:(rank + 1) * rl] if natural else ref[:, rank * cl:(rank + 1) * cl]
        err = (torch.linalg.vector_norm(out - want) / torch.linalg.vector_norm(want)).item()
        result[rank] = err
    finally:
        dist.destroy_process_group()


# every shape and both output layouts on 2 ranks; 4 ranks on the shapes with several chunks (each case spawns its ranks:
# ~7 s of process start-up, which is what bounds the list)
@pytest.mark.parametrize("world,natural,nx,ny,chunks", [(2, False, 16, 8, 1), (2, True, 16, 8, 1), (2, False, 32, 64, 4), (2, True, 32, 64, 4),
                                                        (2, False, 64, 16, 3), (2, True, 64, 16, 3), (4, False, 32, 64, 4), (4, True, 64, 16, 3)])
def test_slab_exchange_gloo(world, natural, nx, ny, chunks):
    port = _free_port()
    ctx = mp.get_context("spawn")
    result = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, ny, chunks, natural, -1, result)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert len(result) == world and max(result.values()) < 1e-5, dict(result)


def test_slab_rejects_indivisible_shapes():
    port = _free_port()
    dist.init_process_group("gloo", init_method="file://" + port, rank=0, world_size=1)
    try:
        from mufft_b200.slab import SlabFFT2D
        plan = SlabFFT2D(8, 8, -1, local_rows=lambda o, i: o.copy_(torch.fft.fft(i, dim=1)), local_cols=lambda o, i: o.copy_(torch.fft.fft(i, dim=0)))
        x = torch.complex(torch.rand(8, 8), torch.rand(8, 8))
        out = torch.empty(8, 8, dtype=torch.complex64)
        plan.execute(out, x)
        assert torch.allclose(out, torch.fft.fft2(x), atol=1e-4)
    finally:
        dist.destroy_process_group()


def _rworker(rank, world, port, nx, ny, chunks, result):
    dist.init_process_group("gloo", init_method="file://" + port, rank=rank, world_size=world)
    try:
        from mufft_b200.slab import SlabRFFT2D
        g = torch.Generator().manual_seed(99)
        full = torch.rand(ny, nx, generator=g) - 0.5

        def r2c_rows(out, inp):
            out.copy_(torch.fft.rfft(inp, dim=1))

        def c2r_rows(out, inp):
            out.copy_(torch.fft.irfft(inp, n=nx, dim=1, norm="forward"))

        def cols(out, inp, direction):
            out.copy_(torch.fft.fft(inp, dim=0) if direction < 0 else torch.fft.ifft(inp, dim=0, norm="forward"))
        plan = SlabRFFT2D(nx, ny, chunks=chunks, local_r2c_rows=r2c_rows, local_c2r_rows=c2r_rows, local_cols=cols)
        rl = ny // world
        lo = rank * plan.cols_even
        mine = full[rank * rl:(rank + 1) * rl].contiguous()
        spec = torch.empty((ny, plan.cols_local), dtype=torch.complex64)
        plan.forward(spec, mine)
        want = torch.fft.rfft2(full)[:, lo:lo + plan.cols_local]
        e1 = (torch.linalg.vector_norm(spec - want) / torch.linalg.vector_norm(want)).item()
        back = torch.empty((rl, nx), dtype=torch.float32)
        plan.inverse(back, spec)
        e2 = (torch.linalg.vector_norm(back / (nx * ny) - mine) / torch.linalg.vector_norm(mine)).item()
        assert plan.cols_local == nx // 2 // world + (1 if rank == world - 1 else 0)
        result[rank] = max(e1, e2)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,ny,chunks", [(1, 32, 64, 4), (2, 16, 8, 1), (2, 32, 64, 4), (2, 64, 16, 3), (4, 64, 16, 3)])
def test_slab_real_exchange_gloo(world, nx, ny, chunks):
    """R2C forward (row slabs -> column slabs of the half spectrum, Nyquist column on the last rank)
    and the C2R inverse from that layout, against torch.fft.rfft2 / a round trip."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    result = ctx.Manager().dict()
    procs = [ctx.Process(target=_rworker, args=(r, world, port, nx, ny, chunks, result)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert len(result) == world and max(result.values()) < 1e-5, dict(result)
