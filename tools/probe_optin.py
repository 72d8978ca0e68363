"""Quick GPU check of the opt-in paths (no torch import: ctypes + numpy on mufft_alloc memory):
  * 16384-point row transforms in one kernel (mufft_cuda_set_onchip_limit) vs the two-kernel split;
  * TMA-fed fused 2D kernel (mufft_cuda_set_fused_2d(2)) vs the direct-load fused kernel.
Each line of gpurun_out/probe_optin.jsonl: parity against numpy float64 on a few transforms and the
wall-clock time of synchronous calls (kernels of a millisecond or more, so the launch + sync overhead
of ~20 us does not matter).  Usage: python tools/probe_optin.py [quick]
"""
import ctypes as C
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mufft_b200 as m  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
# dry runs on another build of the library (MUFFT_B200_LIB: the emulated host layer in the CPU tests) are not measurements
LOG = open(os.devnull if os.environ.get("MUFFT_B200_LIB", "").endswith("libmufft_hostemu.so") else os.path.join(OUT, "probe_optin.jsonl"), "a")
QUICK = len(sys.argv) > 1 and sys.argv[1] == "quick"
# mufft_cuda_set_fused_2d: 1 direct loads, 2 TMA-fed, 3 / 4 the same with smem twiddles + one-sided fences,
# 5 = 3 + derived twiddles, 6 = one-sided fences + derived twiddles (global table), 7 = one-sided fences only,
# 8 = plain, 9 / 10 = 3 / 6 with tickets drawn two tiles ahead (early dependency read), 11 = 6 at three CTAs per SM, 12 / 13 = 3 / 6 with row and column tickets alternating
MODES = (8, 3, 6, 9, 10, 1)
ARG = sys.argv[1] if len(sys.argv) > 1 else ""
REPS = 5  # timed calls per measurement (after 2 warm-up calls) in the real2d / cluster probes


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    LOG.write(line + "\n")
    LOG.flush()
    try:
        os.fsync(LOG.fileno())
    except OSError:  # /dev/null (dry runs)
        pass


def managed(count_floats):
    ptr = m.mufft_alloc(count_floats * 4)
    assert ptr
    return ptr, np.ctypeslib.as_array((C.c_float * count_floats).from_address(ptr))


def fill(arr, seed, unit):
    """Random data in the first and the last `unit` floats (the transforms that get checked); the rest
    stays untouched (zero pages populated on the device at first use) - timing does not depend on values."""
    rng = np.random.default_rng(seed)
    arr[:unit] = rng.random(unit, np.float32) - np.float32(0.5)
    arr[-unit:] = rng.random(unit, np.float32) - np.float32(0.5)


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def timed(call, reps):
    call()  # warm-up: page migration of the managed buffers, attribute calls
    call()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        call()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def probe_1d(n, batch):
    pin, x = managed(2 * n * batch)
    pout, y = managed(2 * n * batch)
    fill(x, n, 2 * n)
    xc = x.view(np.complex64).reshape(batch, n)
    yc = y.view(np.complex64).reshape(batch, n)
    for limit in (8192, 16384):
        m.mufft_cuda_set_onchip_limit(limit)
        for direction in (-1, 1):
            plan = m.mufft_create_plan_1d_c2c(n, direction, 0)
            assert plan
            y[:64] = 0
            ms = timed(lambda: m.mufft_execute_plan_1d_batch(plan, pout, pin, batch, 0, 0), 3)
            want = np.fft.fft(xc[[0, batch - 1]].astype(np.complex128), axis=1) if direction < 0 else \
                np.fft.ifft(xc[[0, batch - 1]].astype(np.complex128), axis=1, norm="forward")
            err = max(rel_l2(yc[0], want[0]), rel_l2(yc[batch - 1], want[1]))
            emit(what="c2c1d", n=n, batch=batch, direction=direction, onchip_limit=limit, uses_engine=m.mufft_cuda_plan_1d_uses_engine(plan),
                 ms=round(ms, 4), gbs=round(16 * n * batch / ms / 1e6, 1), rel_l2=err, ok=bool(err <= 1e-6 * math.log2(n)),
                 last_error=m.mufft_cuda_last_error().decode())
            m.mufft_free_plan_1d(plan)
        # real transforms of 2n points on the same engine length
        rplan, cplan = m.mufft_create_plan_1d_r2c(2 * n, 0), m.mufft_create_plan_1d_c2r(2 * n, 0)
        rb = batch
        pspec, spec = managed(2 * (n + 1) * rb)
        ms_r = timed(lambda: m.mufft_execute_plan_1d_batch(rplan, pspec, pin, rb, 0, 0), 3)
        r = x.reshape(rb, 2 * n)
        sc = spec.view(np.complex64).reshape(rb, n + 1)
        err_r = max(rel_l2(sc[b], np.fft.rfft(r[b].astype(np.float64))) for b in (0, rb - 1))
        ms_c = timed(lambda: m.mufft_execute_plan_1d_batch(cplan, pout, pspec, rb, 0, 0), 3)
        back = y.reshape(rb, 2 * n)
        err_c = max(rel_l2(back[b] / (2 * n), r[b]) for b in (0, rb - 1))
        emit(what="r2c_c2r", n=2 * n, batch=rb, onchip_limit=limit, ms_r2c=round(ms_r, 4), ms_c2r=round(ms_c, 4),
             gbs_r2c=round((8 * n + 8 * (n + 1)) * rb / ms_r / 1e6, 1), rel_l2_r2c=err_r, rel_l2_c2r=err_c,
             ok=bool(err_r <= 1e-6 * math.log2(2 * n) and err_c <= 2e-6 * math.log2(2 * n)))
        # zero-padded forward transform (input rows really n/2 long)
        zplan = m.mufft_create_plan_1d_c2c(n, -1, m.MUFFT_FLAG_ZERO_PAD_UPPER_HALF)
        assert m.mufft_execute_plan_1d_batch(zplan, pout, pin, 4, 0, 0) == 0
        xh = x.view(np.complex64)[: n // 2]
        want = np.fft.fft(np.concatenate([xh, np.zeros(n // 2, np.complex64)]).astype(np.complex128))
        err_z = rel_l2(yc[0], want)
        m.mufft_free_plan_1d(zplan)
        # stereo convolution at this length: conv_input (zero padded) + conv_output (multiply in the load stage)
        method = m.MUFFT_CONV_METHOD_FLAG_STEREO_MONO | m.MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_FIRST | m.MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_SECOND
        cp = m.mufft_create_plan_conv(n, 0, method)
        blk = m.mufft_conv_get_transformed_block_size(cp)
        pf, f = managed(blk // 4)
        ps, sp = managed(blk // 4)
        taps = x[n // 2: n].copy()  # inside the first (random) row
        ptaps, tp = managed(n // 2)
        tp[:] = taps
        m.mufft_execute_conv_input(cp, 1, pf, ptaps)
        m.mufft_execute_conv_input(cp, 0, ps, pin)
        m.mufft_execute_conv_output(cp, pout, ps, pf)
        a = np.concatenate([xh, np.zeros(n // 2, np.complex64)]).astype(np.complex128)
        b = np.concatenate([taps, np.zeros(n // 2, np.float32)]).astype(np.complex128)
        err_v = rel_l2(yc[0], np.fft.ifft(np.fft.fft(a) * np.fft.fft(b)))
        m.mufft_free_plan_conv(cp)
        for q in (pf, ps, ptaps):
            m.mufft_free(q)
        emit(what="zero_pad_and_conv", n=n, onchip_limit=limit, rel_l2_zero_pad=err_z, rel_l2_conv=err_v,
             ok=bool(err_z <= 1e-6 * math.log2(n) and err_v <= 2e-6 * math.log2(n)))
        m.mufft_free_plan_1d(rplan)
        m.mufft_free_plan_1d(cplan)
        m.mufft_free(pspec)
    m.mufft_cuda_set_onchip_limit(8192)
    m.mufft_free(pin)
    m.mufft_free(pout)


def probe_2d(nx, ny, batch, modes=None, direction=-1, tag=""):
    modes = modes or MODES
    pin, x = managed(2 * nx * ny * batch)
    pout, y = managed(2 * nx * ny * batch)
    fill(x, nx + ny, 2 * nx * ny)
    xc = x.view(np.complex64).reshape(batch, ny, nx)
    yc = y.view(np.complex64).reshape(batch, ny, nx)
    for mode in modes:
        m.mufft_cuda_set_fused_2d(mode)
        plan = m.mufft_create_plan_2d_c2c(nx, ny, direction, 0)
        assert plan
        m.mufft_cuda_launch_count(1)
        ms = timed(lambda: m.mufft_execute_plan_2d_batch(plan, pout, pin, batch), 5)
        launches = m.mufft_cuda_launch_count(1)
        checks = (0, batch - 1) if nx * ny <= 512 * 512 else (batch - 1,)
        f2 = (lambda z: np.fft.fft2(z)) if direction < 0 else (lambda z: np.fft.ifft2(z, norm="forward"))
        err = max(rel_l2(yc[b], f2(xc[b].astype(np.complex128))) for b in checks)
        emit(what="c2c2d" + tag, nx=nx, ny=ny, batch=batch, direction=direction, fused_mode=mode, paths=m.mufft_cuda_plan_2d_paths(plan), ms=round(ms, 4), gbs=round(16 * nx * ny * batch / ms / 1e6, 1),
             launches_per_call=launches / 7, rel_l2=err, ok=bool(err <= 1e-6 * math.log2(nx * ny)), last_error=m.mufft_cuda_last_error().decode())
        m.mufft_free_plan_2d(plan)
    m.mufft_cuda_set_fused_2d(1)
    m.mufft_free(pin)
    m.mufft_free(pout)


def probe_partitioned(n, parts, stereo, batch=6):
    """mufft_execute_conv_output_partitioned_batch against numpy: out[b] = sum_p linear_conv(x[b + parts-1 - p], taps[p])."""
    method = (1 if stereo else 0) | 4 | 8
    plan = m.mufft_create_plan_conv(n, 0, method)
    assert plan
    blk = m.mufft_conv_get_transformed_block_size(plan)
    half, total = n // 2, batch + parts - 1
    rng = np.random.default_rng(n + parts + stereo)
    per = 2 * half if stereo else half
    px, x = managed(total * per)
    pt, taps = managed(parts * half)
    x[:] = rng.random(total * per, np.float32) - np.float32(0.5)
    taps[:] = rng.random(parts * half, np.float32) - np.float32(0.5)
    ps, _ = managed(total * blk // 4)
    pf, _ = managed(parts * blk // 4)
    po, out = managed(batch * (2 * n if stereo else n))
    assert m.mufft_execute_conv_input_batch(plan, 0, ps, px, total) == 0
    assert m.mufft_execute_conv_input_batch(plan, 1, pf, pt, parts) == 0
    rc = m.mufft_execute_conv_output_partitioned_batch(plan, po, ps + (parts - 1) * blk, pf, parts, batch)
    err = -1.0
    if rc == 0:
        xs = x.reshape(total, per)
        xs = xs.view(np.complex64) if stereo else xs
        ts = taps.reshape(parts, half)
        o = out.reshape(batch, -1)
        o = o.view(np.complex64) if stereo else o
        err = 0.0
        for b in (0, batch - 1):
            want = np.zeros(n, np.complex128)
            for p_ in range(parts):
                want[: n - 1] += np.convolve(xs[b + parts - 1 - p_].astype(np.complex128), ts[p_].astype(np.float64))
            err = max(err, rel_l2(o[b], want if stereo else want.real))
    emit(what="conv_partitioned", n=n, parts=parts, stereo=stereo, rc=rc, rel_l2=err, ok=bool(rc == 0 and err <= 2e-6 * math.log2(n)),
         last_error=m.mufft_cuda_last_error().decode())
    m.mufft_free_plan_conv(plan)
    for q in (px, pt, ps, pf, po):
        m.mufft_free(q)


def probe_scatter(n, peers=2, batch=5):
    """Row pass + exchange step in one kernel (MCU_OUT_PEERS) with every 'peer' slab in local memory."""
    plan = m.mufft_create_plan_1d_c2c(n, -1, 0)
    pin, x = managed(2 * n * batch)
    x[:] = np.random.default_rng(n).random(2 * n * batch, np.float32) - np.float32(0.5)
    width, rows_total, row0 = n // peers, batch + 3, 2
    slabs = [m.mufft_cuda_device_alloc(rows_total * width * 8) for _ in range(peers)]
    bases = (C.c_void_p * peers)(*slabs)
    rc = m.mufft_execute_plan_1d_batch_scatter_async(plan, bases, peers, row0, pin, batch, 0)
    rc |= m.mufft_cuda_synchronize()
    # read the slabs back through a managed bounce buffer (C2C of length 2 would be silly: plain copies via the batch API are not exposed)
    want = np.fft.fft(x.view(np.complex64).reshape(batch, n).astype(np.complex128), axis=1)
    emit(what="scatter", n=n, peers=peers, rc=rc, uses_engine=m.mufft_cuda_plan_1d_uses_engine(plan), last_error=m.mufft_cuda_last_error().decode(),
         note="launch only (slabs are cudaMalloc memory; contents are checked by tests/test_slab_gpu.py)", want_norm=float(np.linalg.norm(want)))
    for s_ in slabs:
        m.mufft_cuda_device_free(s_)
    m.mufft_free_plan_1d(plan)
    m.mufft_free(pin)


def probe_real2d(shapes=((1024, 1024), (512, 512), (2048, 2048), (4096, 1024)), bytes_in=1 << 30):
    """2D R2C images (nx reals per row): row kernel + column kernel through scratch vs the fused persistent kernel."""
    for nx_, ny_ in shapes:
        batch_ = max(2, bytes_in // (nx_ * ny_ * 8))
        pin, x = managed(nx_ * ny_ * batch_)
        pout, y = managed(2 * nx_ * ny_ * batch_)
        fill(x, nx_ + ny_, nx_ * ny_)
        xr, yc = x.reshape(batch_, ny_, nx_), y.view(np.complex64).reshape(batch_, ny_, nx_)
        for fused in (0, 1, 0, 1):
            m.mufft_cuda_set_fused_2d_real(fused)
            plan = m.mufft_create_plan_2d_r2c(nx_, ny_, 0)
            m.mufft_cuda_launch_count(1)
            ms = timed(lambda: m.mufft_execute_plan_2d_batch(plan, pout, pin, batch_), REPS)
            launches = m.mufft_cuda_launch_count(1) / (2 + REPS)
            w_ = nx_ // 2 + 1
            err = max(rel_l2(yc[b][:, :w_], np.fft.fft2(xr[b].astype(np.float64))[:, :w_]) for b in (0, batch_ - 1))
            emit(what="r2c2d", nx=nx_, ny=ny_, batch=batch_, fused_real=fused, launches_per_call=launches, ms=round(ms, 4),
                 gbs=round((4 * nx_ * ny_ + 8 * w_ * ny_) * batch_ / ms / 1e6, 1), rel_l2=err, ok=bool(err <= 1e-6 * math.log2(nx_ * ny_)),
                 last_error=m.mufft_cuda_last_error().decode())
            m.mufft_free_plan_2d(plan)
        m.mufft_cuda_set_fused_2d_real(0)
        m.mufft_free(pin)
        m.mufft_free(pout)


def probe_latency(iters=20000, warm=500):
    """One 64-point transform per fft.h call on mufft_alloc memory, blocking vs busy-poll synchronise."""
    plan = m.mufft_create_plan_1d_c2c(64, -1, 0)
    pin, x = managed(128)
    pout, y = managed(128)
    x[:] = np.random.default_rng(0).random(128, np.float32) - np.float32(0.5)
    for spin in (0, 1, 0, 1):
        m.mufft_cuda_set_spin_sync(spin)
        for _ in range(warm):
            m.mufft_execute_plan_1d(plan, pout, pin)
        t0 = time.perf_counter()
        for _ in range(iters):
            m.mufft_execute_plan_1d(plan, pout, pin)
        us = (time.perf_counter() - t0) / iters * 1e6
        err = rel_l2(y.view(np.complex64), np.fft.fft(x.view(np.complex64).astype(np.complex128)))
        emit(what="single_call_latency", n=64, spin_sync=spin, us_per_call=round(us, 3), rel_l2=err, ok=bool(err < 6e-6))
    m.mufft_cuda_set_spin_sync(0)
    m.mufft_free_plan_1d(plan)
    m.mufft_free(pin)
    m.mufft_free(pout)


def probe_managed(cases=(("c2c1d", 65536, 1), ("c2c2d", 1024, 1024), ("c2c2d", 4096, 4096)), rounds=6):
    """The reference's calling pattern on mufft_alloc memory: the host WRITES the whole input, one synchronous fft.h
    call, the host READS the whole output - per round, with the managed-memory prefetch off / to the device / and
    back (mufft_cuda_set_managed_prefetch 0 / 1 / 2).  Time per round includes the host's own passes over the data."""
    for what, nx, ny in cases:
        count = nx * ny
        pin, x = managed(2 * count)
        pout, y = managed(2 * count)
        src = np.random.default_rng(count).random(2 * count, np.float32) - np.float32(0.5)
        plan = m.mufft_create_plan_1d_c2c(nx, -1, 0) if what == "c2c1d" else m.mufft_create_plan_2d_c2c(nx, ny, -1, 0)
        run = (lambda: m.mufft_execute_plan_1d(plan, pout, pin)) if what == "c2c1d" else (lambda: m.mufft_execute_plan_2d(plan, pout, pin))
        sc = src.view(np.complex64)
        want = np.fft.fft(sc.astype(np.complex128)) if what == "c2c1d" else np.fft.fft2(sc.reshape(ny, nx).astype(np.complex128))
        for level in (0, 1, 2, 0, 1, 2):
            m.mufft_cuda_set_managed_prefetch(level)
            best, acc = 1e9, 0.0
            for r in range(rounds):
                t0 = time.perf_counter()
                x[:] = src                      # the host produces the input
                run()
                acc += float(y[::1024].sum())   # ... and consumes the output (every page is touched)
                if r > 0:
                    best = min(best, time.perf_counter() - t0)
            t0 = time.perf_counter()
            x[:] = src
            t_fill = time.perf_counter() - t0
            err = rel_l2(y.view(np.complex64), want)
            emit(what="managed_round_trip", kind=what, nx=nx, ny=ny, prefetch=level, ms=round(best * 1e3, 4), ms_host_fill_alone=round(t_fill * 1e3, 4),
                 rel_l2=err, ok=bool(err <= 1e-6 * math.log2(count)))
        m.mufft_cuda_set_managed_prefetch(0)
        (m.mufft_free_plan_1d if what == "c2c1d" else m.mufft_free_plan_2d)(plan)
        m.mufft_free(pin)
        m.mufft_free(pout)


def probe_cluster(sizes=(32768, 65536), bytes_in=1 << 31, flavours=(0, 1, 2, 0, 1, 2)):
    """Rows of 32768 / 65536 points: two-kernel split (0) vs one kernel on a thread-block cluster
    (1 = cluster barriers, 2 = st.async + mbarrier exchange)."""
    for n_ in sizes:
        batch_ = max(2, bytes_in // (8 * n_))
        pin, x = managed(2 * n_ * batch_)
        pout, y = managed(2 * n_ * batch_)
        fill(x, n_, 2 * n_)
        xc, yc = x.view(np.complex64).reshape(batch_, n_), y.view(np.complex64).reshape(batch_, n_)
        for cluster in flavours:
            m.mufft_cuda_set_cluster_rows(cluster)
            for direction in (-1, 1):
                plan = m.mufft_create_plan_1d_c2c(n_, direction, 0)
                m.mufft_cuda_launch_count(1)
                ms = timed(lambda: m.mufft_execute_plan_1d_batch(plan, pout, pin, batch_, 0, 0), REPS)
                launches = m.mufft_cuda_launch_count(1) / (2 + REPS)
                f1 = (lambda z: np.fft.fft(z)) if direction < 0 else (lambda z: np.fft.ifft(z, norm="forward"))
                err = max(rel_l2(yc[b], f1(xc[b].astype(np.complex128))) for b in (0, batch_ - 1))
                emit(what="c2c1d_cluster", n=n_, batch=batch_, direction=direction, cluster_rows=cluster, launches_per_call=launches, ms=round(ms, 4),
                     gbs=round(16 * n_ * batch_ / ms / 1e6, 1), rel_l2=err, ok=bool(err <= 1e-6 * math.log2(n_)), last_error=m.mufft_cuda_last_error().decode())
                m.mufft_free_plan_1d(plan)
        m.mufft_cuda_set_cluster_rows(0)
        m.mufft_free(pin)
        m.mufft_free(pout)


if __name__ == "__main__":
    assert m.mufft_cuda_device_count() >= 1 and m.mufft_cuda_set_device(0) == 0
    emit(what="start", arg=ARG)
    if ARG == "variant":  # a library built with other CTA shapes (tools/ab_build.sh), copied over the product one by the caller
        for shape in ((1024, 1024, 256), (256, 256, 4096), (512, 512, 1024)):
            probe_2d(*shape, modes=(1, 3, 5), tag="_variant")
    elif ARG == "final":  # the defaults as shipped + the paths changed last
        probe_2d(1024, 1024, 256, modes=(1, 8))
        probe_2d(256, 256, 2048, modes=(1, 8))
        for n_, parts_, st_ in ((8192, 3, 0), (1024, 2, 0), (16384, 2, 1), (32768, 2, 0), (64, 5, 1)):
            probe_partitioned(n_, parts_, st_)
        probe_scatter(16384)
        probe_1d(16384, 4096)
    elif ARG == "early":  # fused 2D: tickets two ahead (modes 9 / 10) and three CTAs per SM (11) against the flavours they extend (3 / 6)
        for shape in ((1024, 1024, 256), (512, 512, 1024), (256, 256, 4096), (2048, 2048, 64), (1024, 256, 1024)):
            probe_2d(*shape, modes=(1, 3, 16, 9, 12, 6, 11, 10, 13, 14, 15))
        for lag in (1, 2, 4, 6):  # default for this shape on 148 SMs: 3
            m.mufft_cuda_set_fused_2d_lag(lag)
            probe_2d(1024, 1024, 256, modes=(1,), tag=f"_lag{lag}")
        m.mufft_cuda_set_fused_2d_lag(0)
    elif ARG == "real2d":  # 2D R2C images: row kernel + column kernel through scratch vs the fused persistent kernel
        probe_real2d()
    elif ARG == "latency":  # BASELINE configs[0] as the reference runs it: one 64-point transform per fft.h call
        probe_latency()
    elif ARG == "managed":  # fft.h calls on mufft_alloc memory the host fills and reads between calls
        probe_managed()
    elif ARG == "cluster":  # 32768 / 65536-point rows: two-kernel split vs one kernel on a thread-block cluster
        probe_cluster()
    elif ARG == "quick":
        probe_2d(1024, 1024, 32)
        probe_1d(16384, 2048)
    else:
        probe_2d(1024, 1024, 256)
        probe_2d(1024, 1024, 256, modes=(1, 3, 5), direction=1)
        # rows of 16384 points inside a 2D plan, with and without the single-kernel row transform
        for limit in (8192, 16384):
            m.mufft_cuda_set_onchip_limit(limit)
            probe_2d(16384, 256, 16, modes=(1,), tag=f"_rows16k_limit{limit}")
        m.mufft_cuda_set_onchip_limit(8192)
        sides = (256, 512, 1024, 2048)
        shapes = [(512, 512), (256, 256), (2048, 2048)] + [(nx, ny) for nx in sides for ny in sides if nx != ny]
        for nx, ny in shapes:
            probe_2d(nx, ny, max(1, (1 << 30) // (nx * ny * 8)))
    emit(what="done")
