"""Seconds-long GPU measurements without importing torch (a fresh box spends about a minute on that
import): ctypes + numpy, data resident in device memory (mufft_cuda_device_alloc), the asynchronous batch
calls of mufft_cuda.h timed with CUDA events on the library's stream (best and median of --reps after 2
warm-up calls), parity of the first and last unit against numpy float64.  For A/B loops under gpurun;
bench.py stays the measurement of record.

    python tools/fastbench.py c2c4096 r2c4096 c2r4096 conv8192 conv8192_fused c2c2d1024 r2c2d1024 c2r2d1024 c2c32768
    python tools/fastbench.py --label two_kernel --flags 0x200000 c2c32768:8192      # name[:units], plan flags of mufft_cuda.h
Lines of JSON go to stdout and gpurun_out/fastbench.jsonl.
"""
import argparse
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

PEAK_GBS = 6546.9
try:
    PEAK_GBS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
DEFAULT_BYTES = {"c2c": 1 << 30, "r2c": 1 << 30, "c2r": 1 << 30, "conv": 1 << 29, "2d": 1 << 31}  # bytes of input per run
DRY = os.environ.get("MUFFT_B200_LIB", "").endswith("libmufft_hostemu.so")  # the emulated host layer: not a measurement


class Events:
    """CUDA events through the CUDA runtime (a second runtime instance shares the primary context with the
    library); wall clock when the runtime cannot be loaded (dry runs on the emulation)."""

    def __init__(self):
        self.rt = None
        if not DRY:
            for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
                try:
                    self.rt = C.CDLL(name)
                    break
                except OSError:
                    continue
        if self.rt is not None:
            self.a, self.b = C.c_void_p(), C.c_void_p()
            if self.rt.cudaEventCreate(C.byref(self.a)) != 0 or self.rt.cudaEventCreate(C.byref(self.b)) != 0:
                self.rt = None

    def time_ms(self, call):
        if self.rt is None:
            t0 = time.perf_counter()
            rc = call()
            m.mufft_cuda_synchronize()
            return rc, (time.perf_counter() - t0) * 1e3
        self.rt.cudaEventRecord(self.a, m.mufft_cuda_get_stream())
        rc = call()
        self.rt.cudaEventRecord(self.b, m.mufft_cuda_get_stream())
        self.rt.cudaEventSynchronize(self.b)
        ms = C.c_float()
        self.rt.cudaEventElapsedTime(C.byref(ms), self.a, self.b)
        return rc, float(ms.value)


EV = None


def rnd(rng, count):
    return rng.random(count, np.float32) - np.float32(0.5)


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


class Dev:
    """Device buffer of `floats` floats, zero-filled, with host<->device access to single units."""

    def __init__(self, floats):
        self.floats = floats
        self.ptr = m.mufft_cuda_device_alloc(floats * 4)
        assert self.ptr, "device allocation failed: " + m.mufft_cuda_last_error().decode()
        assert m.mufft_cuda_memset(self.ptr, 0, floats * 4) == 0

    def put(self, offset_floats, arr):
        arr = np.ascontiguousarray(arr, np.float32)
        assert m.mufft_cuda_copy(self.ptr + 4 * offset_floats, arr.ctypes.data, arr.nbytes) == 0

    def get(self, offset_floats, count):
        out = np.empty(count, np.float32)
        assert m.mufft_cuda_copy(out.ctypes.data, self.ptr + 4 * offset_floats, out.nbytes) == 0
        return out

    def free(self):
        m.mufft_cuda_device_free(self.ptr)


def timed(call, reps):
    for _ in range(2):
        rc, _ = EV.time_ms(call)
        assert rc == 0, m.mufft_cuda_last_error()
    times = sorted(EV.time_ms(call)[1] for _ in range(reps))
    return times[0], times[len(times) // 2]


def run(name, units, reps, flags):
    rng = np.random.default_rng(len(name))
    bufs = []

    def dev(floats):
        d = Dev(floats)
        bufs.append(d)
        return d

    def ends(units_):
        return sorted({0, units_ - 1})

    if name[:5] in ("c2c2d", "r2c2d", "c2r2d"):
        kind = name[:3]
        dims = name[5:].split("x")
        nx, ny = int(dims[0]), int(dims[-1])
        units = units or max(1, DEFAULT_BYTES["2d"] // (8 * nx * ny))
        if kind == "c2c":
            x, y = dev(2 * nx * ny * units), dev(2 * nx * ny * units)
            plan = m.mufft_create_plan_2d_c2c(nx, ny, -1, flags)
            host = {b: rnd(rng, 2 * nx * ny) for b in ends(units)}
            for b, h in host.items():
                x.put(b * 2 * nx * ny, h)
            best, med = timed(lambda: m.mufft_execute_plan_2d_batch_async(plan, y.ptr, x.ptr, units), reps)
            err = max(rel_l2(y.get(b * 2 * nx * ny, 2 * nx * ny).view(np.complex64).reshape(ny, nx),
                             np.fft.fft2(h.view(np.complex64).reshape(ny, nx).astype(np.complex128))) for b, h in host.items())
            nbytes, flops = 16.0 * nx * ny, 5.0 * nx * ny * math.log2(nx * ny)
        elif kind == "r2c":
            x, y = dev(nx * ny * units), dev(2 * nx * ny * units)
            plan = m.mufft_create_plan_2d_r2c(nx, ny, flags)
            host = {b: rnd(rng, nx * ny) for b in ends(units)}
            for b, h in host.items():
                x.put(b * nx * ny, h)
            best, med = timed(lambda: m.mufft_execute_plan_2d_batch_async(plan, y.ptr, x.ptr, units), reps)
            w = nx // 2 + 1
            err = max(rel_l2(y.get(b * 2 * nx * ny, 2 * nx * ny).view(np.complex64).reshape(ny, nx)[:, :w],
                             np.fft.rfft2(h.reshape(ny, nx).astype(np.float64))) for b, h in host.items())
            nbytes, flops = 4.0 * nx * ny + 8.0 * w * ny, 2.5 * nx * ny * math.log2(nx * ny)
        else:
            x, y = dev(2 * nx * ny * units), dev(nx * ny * units)
            plan = m.mufft_create_plan_2d_c2r(nx, ny, flags)
            w = nx // 2 + 1
            host = {}
            for b in ends(units):
                real = rnd(rng, nx * ny).reshape(ny, nx)
                spec = np.zeros((ny, nx), np.complex64)
                spec[:, :w] = np.fft.rfft2(real.astype(np.float64))
                x.put(b * 2 * nx * ny, spec.view(np.float32))
                host[b] = real
            best, med = timed(lambda: m.mufft_execute_plan_2d_batch_async(plan, y.ptr, x.ptr, units), reps)
            err = max(rel_l2(y.get(b * nx * ny, nx * ny) / (nx * ny), h) for b, h in host.items())
            nbytes, flops = 4.0 * nx * ny + 8.0 * w * ny, 2.5 * nx * ny * math.log2(nx * ny)
        m.mufft_free_plan_2d(plan)
        tol = 1e-6 * math.log2(nx * ny)
    elif name.startswith("e2e"):
        # the host-pointer path: pinned host buffers through the synchronous batch call (H2D, kernels, D2H pipelined by the
        # library), wall clock; next to it what plain pinned copies of the same buffers reach both ways at once
        n = int(name[3:])
        units = units or max(1, (1 << 31) // (8 * n))
        nbytes_buf = 8 * n * units
        hin, hout = m.mufft_cuda_host_alloc(nbytes_buf), m.mufft_cuda_host_alloc(nbytes_buf)
        assert hin and hout, m.mufft_cuda_last_error()
        vin = np.ctypeslib.as_array((C.c_float * (2 * n * units)).from_address(hin))
        vout = np.ctypeslib.as_array((C.c_float * (2 * n * units)).from_address(hout))
        vin[:] = 0
        host = {b: rnd(rng, 2 * n) for b in ends(units)}
        for b, h in host.items():
            vin[b * 2 * n:(b + 1) * 2 * n] = h
        plan = m.mufft_create_plan_1d_c2c(n, -1, flags)
        times = []
        for i in range(1 + reps):
            t0 = time.perf_counter()
            assert m.mufft_execute_plan_1d_batch(plan, hout, hin, units, 0, 0) == 0, m.mufft_cuda_last_error()
            if i:
                times.append((time.perf_counter() - t0) * 1e3)
        times.sort()
        best, med = times[0], times[len(times) // 2]
        err = max(rel_l2(vout[b * 2 * n:(b + 1) * 2 * n].view(np.complex64), np.fft.fft(h.view(np.complex64).astype(np.complex128))) for b, h in host.items())
        m.mufft_free_plan_1d(plan)
        bus = None
        if EV.rt is not None:  # raw bus rate: the same buffers, plain copies both ways at once on two streams
            d1, d2 = dev(min(2 * n * units, 1 << 28)), dev(min(2 * n * units, 1 << 28))
            cb = min(nbytes_buf, 1 << 30)
            s1, s2 = C.c_void_p(), C.c_void_p()
            EV.rt.cudaStreamCreateWithFlags(C.byref(s1), 1)
            EV.rt.cudaStreamCreateWithFlags(C.byref(s2), 1)
            EV.rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
            rates = []
            for _ in range(3):
                EV.rt.cudaDeviceSynchronize()
                t0 = time.perf_counter()
                EV.rt.cudaMemcpyAsync(d1.ptr, hin, cb, 4, s1)
                EV.rt.cudaMemcpyAsync(hout, d2.ptr, cb, 4, s2)
                EV.rt.cudaDeviceSynchronize()
                rates.append(2 * cb / (time.perf_counter() - t0) * 1e-9)
            bus = max(rates)
            # the timed output was overwritten by the probe copy: fine, parity was taken before
        m.mufft_cuda_host_free(hin)
        m.mufft_cuda_host_free(hout)
        for d in bufs:
            d.free()
        gbs = 2.0 * nbytes_buf / med / 1e6
        return {"workload": name, "units": units, "ms": round(med, 3), "ms_best": round(best, 3), "bus_gbs": round(gbs, 1),
                "raw_both_ways_gbs": round(bus, 1) if bus else None, "over_bus": round(gbs / bus, 3) if bus else None,
                "gflops": round(5.0 * n * math.log2(n) * units / med / 1e6, 1), "rel_l2": err, "ok": bool(err <= 2e-6 * math.log2(n)),
                "gbs": round(gbs, 1), "frac": 0.0}
    elif name.startswith("cols"):
        # column pass of a slab transform: the plan's N-point C2C transform down the columns of ONE [N][W] matrix, in place
        n, w = (int(v) for v in name[4:].split("x"))
        units = 1
        x = dev(2 * n * w)
        probe = sorted({0, w // 2 + 1, w - 1})
        host = {c: rnd(rng, 2 * n) for c in probe}
        # a full host copy is small enough at these sizes (n * w * 8 bytes <= 1 GiB): fill on the host, copy once
        full = np.zeros((n, w, 2), np.float32)
        for c, h in host.items():
            full[:, c, :] = h.reshape(n, 2)
        x.put(0, full.ravel())
        del full
        plan = m.mufft_create_plan_1d_c2c(n, -1, flags)
        # in place and repeated: only the first run transforms the probe columns; time on a second buffer
        y = dev(2 * n * w)
        best, med = timed(lambda: m.mufft_execute_plan_1d_columns(plan, y.ptr, x.ptr, w, w * 8, w * 8), reps)
        got = y.get(0, 2 * n * w).reshape(n, w, 2)
        err = max(rel_l2(got[:, c, :].copy().view(np.complex64), np.fft.fft(h.view(np.complex64).astype(np.complex128))) for c, h in host.items())
        del got
        m.mufft_free_plan_1d(plan)
        nbytes, flops, tol = 16.0 * n * w, 5.0 * n * w * math.log2(n), 1e-6 * math.log2(n)
    elif name.startswith(("c2c", "r2c", "c2r")):
        kind, n = name[:3], int(name[3:])
        if kind == "c2c":
            units = units or max(1, DEFAULT_BYTES["c2c"] // (8 * n))
            x, y = dev(2 * n * units), dev(2 * n * units)
            host = {b: rnd(rng, 2 * n) for b in ends(units)}
            for b, h in host.items():
                x.put(b * 2 * n, h)
            plan = m.mufft_create_plan_1d_c2c(n, -1, flags)
            best, med = timed(lambda: m.mufft_execute_plan_1d_batch_async(plan, y.ptr, x.ptr, units, 0, 0), reps)
            err = max(rel_l2(y.get(b * 2 * n, 2 * n).view(np.complex64), np.fft.fft(h.view(np.complex64).astype(np.complex128))) for b, h in host.items())
            nbytes, flops = 16.0 * n, 5.0 * n * math.log2(n)
        else:
            units = units or max(1, DEFAULT_BYTES[kind] // (4 * n))
            bins = n // 2 + 1
            r, s = dev(n * units), dev(2 * bins * units)
            host = {b: rnd(rng, n) for b in ends(units)}
            for b, h in host.items():
                r.put(b * n, h)
            fwd = m.mufft_create_plan_1d_r2c(n, flags)
            best, med = timed(lambda: m.mufft_execute_plan_1d_batch_async(fwd, s.ptr, r.ptr, units, 0, 0), reps)
            err = max(rel_l2(s.get(b * 2 * bins, 2 * bins).view(np.complex64), np.fft.rfft(h.astype(np.float64))) for b, h in host.items())
            m.mufft_free_plan_1d(fwd)
            plan = None
            if kind == "c2r":
                back = dev(n * units)
                plan = m.mufft_create_plan_1d_c2r(n, flags)
                best, med = timed(lambda: m.mufft_execute_plan_1d_batch_async(plan, back.ptr, s.ptr, units, 0, 0), reps)
                err = max(err, max(rel_l2(back.get(b * n, n) / n, h) for b, h in host.items()))
            nbytes, flops = 4.0 * n + 8.0 * bins, 2.5 * n * math.log2(n)
        if plan:
            m.mufft_free_plan_1d(plan)
        tol = 2e-6 * math.log2(n)
    elif name.startswith("conv"):
        fused = name.endswith("_fused")
        n = int(name[4:].split("_")[0])
        units = units or max(1, DEFAULT_BYTES["conv"] // (4 * n))
        method = m.MUFFT_CONV_METHOD_FLAG_STEREO_MONO | m.MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_FIRST | m.MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_SECOND
        plan = m.mufft_create_plan_conv(n, flags, method)
        blk = m.mufft_conv_get_transformed_block_size(plan)
        half = n // 2
        x, taps_d, fspec, y = dev(2 * half * units), dev(half), dev(blk // 4), dev(2 * n * units)
        taps = rnd(rng, half)
        taps_d.put(0, taps)
        host = {b: rnd(rng, 2 * half) for b in ends(units)}
        for b, h in host.items():
            x.put(b * 2 * half, h)
        assert m.mufft_execute_conv_input_batch(plan, 1, fspec.ptr, taps_d.ptr, 1) == 0
        if fused:
            best, med = timed(lambda: m.mufft_execute_conv_batch_async(plan, y.ptr, x.ptr, fspec.ptr, units, 0), reps)
            nbytes = 12.0 * n
        else:
            spec = dev(units * blk // 4)

            def both():
                return (m.mufft_execute_conv_input_batch_async(plan, 0, spec.ptr, x.ptr, units) |
                        m.mufft_execute_conv_output_batch_async(plan, y.ptr, spec.ptr, fspec.ptr, units, 0))
            best, med = timed(both, reps)
            nbytes = 28.0 * n
        err = 0.0
        for b, h in host.items():
            want = np.zeros(n, np.complex128)
            want[: n - 1] = np.convolve(h.view(np.complex64).astype(np.complex128), taps.astype(np.float64))
            err = max(err, rel_l2(y.get(b * 2 * n, 2 * n).view(np.complex64), want))
        m.mufft_free_plan_conv(plan)
        flops, tol = 2 * 5.0 * n * math.log2(n) + 6.0 * n, 2e-6 * math.log2(n)
    else:
        raise SystemExit(f"unknown workload {name}")
    for d in bufs:
        d.free()
    gbs = nbytes * units / med / 1e6
    return {"workload": name, "units": units, "ms": round(med, 4), "ms_best": round(best, 4), "gbs": round(gbs, 1), "frac": round(gbs / PEAK_GBS, 3),
            "gflops": round(flops * units / med / 1e6, 0), "rel_l2": err, "ok": bool(err <= tol)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="+")
    ap.add_argument("--label", default="")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--flags", type=lambda v: int(v, 0), default=0, help="plan-creation flags (mufft_cuda.h MUFFT_CUDA_FLAG_*)")
    args = ap.parse_args()
    assert m.mufft_cuda_device_count() >= 1 and m.mufft_cuda_set_device(0) == 0
    EV = Events()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.devnull if DRY else os.path.join(ROOT, "gpurun_out", "fastbench.jsonl"), "a") as log:
        for w in args.workloads:
            name, _, u = w.partition(":")
            res = run(name, int(u) if u else 0, args.reps, args.flags)
            res["label"] = args.label
            if args.flags:
                res["flags"] = hex(args.flags)
            line = json.dumps(res)
            print(line, flush=True)
            log.write(line + "\n")
            log.flush()
