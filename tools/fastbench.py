"""Seconds-long GPU measurements without importing torch (a fresh box spends about a minute on that
import): ctypes + numpy on mufft_alloc (managed) memory, wall clock around synchronous batch calls
of a millisecond or more (best of --reps after 2 warm-up calls, ~0.5 % above CUDA-event time),
parity of the first and last unit against numpy float64.  For A/B loops under gpurun; bench.py
stays the measurement of record.

    python tools/fastbench.py c2c4096 r2c4096 c2r4096 conv8192 conv8192_fused c2c2d1024 c2c16384
    python tools/fastbench.py --label tw_all --reps 7 c2c8192:65536      # name[:units]
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
DEFAULT_UNITS = {"c2c": 1 << 30, "r2c": 1 << 30, "c2r": 1 << 30, "conv": 1 << 29, "c2c2d": 1 << 31}  # bytes of input per run


def managed(count_floats):
    ptr = m.mufft_alloc(count_floats * 4)
    assert ptr, "mufft_alloc failed"
    return ptr, np.ctypeslib.as_array((C.c_float * count_floats).from_address(ptr))


def rnd(rng, count):
    return rng.random(count, np.float32) - np.float32(0.5)


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def timed(call, reps):
    for _ in range(2):
        assert call() == 0, m.mufft_cuda_last_error()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        call()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def run(name, units, reps):
    rng = np.random.default_rng(len(name))
    frees = []

    def buf(count):
        p, a = managed(count)
        frees.append(p)
        return p, a

    if name.startswith("c2c2d"):
        n = int(name[5:])
        units = units or max(1, DEFAULT_UNITS["c2c2d"] // (8 * n * n))
        pin, x = buf(2 * n * n * units)
        pout, y = buf(2 * n * n * units)
        x[:2 * n * n] = rnd(rng, 2 * n * n)
        x[-2 * n * n:] = rnd(rng, 2 * n * n)
        plan = m.mufft_create_plan_2d_c2c(n, n, -1, 0)
        ms = timed(lambda: m.mufft_execute_plan_2d_batch(plan, pout, pin, units), reps)
        xc, yc = x.view(np.complex64).reshape(units, n, n), y.view(np.complex64).reshape(units, n, n)
        err = max(rel_l2(yc[b], np.fft.fft2(xc[b].astype(np.complex128))) for b in {0, units - 1})
        m.mufft_free_plan_2d(plan)
        nbytes, flops, tol = 16.0 * n * n, 5.0 * n * n * 2 * math.log2(n), 1e-6 * 2 * math.log2(n)
    elif name.startswith(("c2c", "r2c", "c2r")):
        kind, n = name[:3], int(name[3:])
        if kind == "c2c":
            units = units or max(1, DEFAULT_UNITS["c2c"] // (8 * n))
            pin, x = buf(2 * n * units)
            pout, y = buf(2 * n * units)
            x[:2 * n] = rnd(rng, 2 * n)
            x[-2 * n:] = rnd(rng, 2 * n)
            plan = m.mufft_create_plan_1d_c2c(n, -1, 0)
            ms = timed(lambda: m.mufft_execute_plan_1d_batch(plan, pout, pin, units, 0, 0), reps)
            xc, yc = x.view(np.complex64).reshape(units, n), y.view(np.complex64).reshape(units, n)
            err = max(rel_l2(yc[b], np.fft.fft(xc[b].astype(np.complex128))) for b in {0, units - 1})
            nbytes, flops = 16.0 * n, 5.0 * n * math.log2(n)
        else:
            units = units or max(1, DEFAULT_UNITS[kind] // (4 * n))
            bins = n // 2 + 1
            preal, r = buf(n * units)
            pspec, s = buf(2 * bins * units)
            r[:n] = rnd(rng, n)
            r[-n:] = rnd(rng, n)
            fwd = m.mufft_create_plan_1d_r2c(n, 0)
            ms = timed(lambda: m.mufft_execute_plan_1d_batch(fwd, pspec, preal, units, 0, 0), reps)
            rr, sc = r.reshape(units, n), s.view(np.complex64).reshape(units, bins)
            err = max(rel_l2(sc[b], np.fft.rfft(rr[b].astype(np.float64))) for b in {0, units - 1})
            m.mufft_free_plan_1d(fwd)
            if kind == "c2r":
                pback, back = buf(n * units)
                plan = m.mufft_create_plan_1d_c2r(n, 0)
                ms = timed(lambda: m.mufft_execute_plan_1d_batch(plan, pback, pspec, units, 0, 0), reps)
                bb = back.reshape(units, n)
                err = max(err, max(rel_l2(bb[b] / n, rr[b]) for b in {0, units - 1}))
            else:
                plan = None
            nbytes, flops = 4.0 * n + 8.0 * bins, 2.5 * n * math.log2(n)
        if plan:
            m.mufft_free_plan_1d(plan)
        tol = 2e-6 * math.log2(n)
    elif name.startswith("conv"):
        fused = name.endswith("_fused")
        n = int(name[4:].split("_")[0])
        units = units or max(1, DEFAULT_UNITS["conv"] // (4 * n))
        method = m.MUFFT_CONV_METHOD_FLAG_STEREO_MONO | m.MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_FIRST | m.MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_SECOND
        plan = m.mufft_create_plan_conv(n, 0, method)
        blk = m.mufft_conv_get_transformed_block_size(plan)
        half = n // 2
        pin, x = buf(2 * half * units)
        ptaps, taps = buf(half)
        pf, _ = buf(blk // 4)
        pout, y = buf(2 * n * units)
        x[:2 * half] = rnd(rng, 2 * half)
        x[-2 * half:] = rnd(rng, 2 * half)
        taps[:] = rnd(rng, half)
        m.mufft_execute_conv_input(plan, 1, pf, ptaps)
        if fused:
            ms = timed(lambda: m.mufft_execute_conv_batch(plan, pout, pin, pf, units, 0), reps)
            nbytes = 12.0 * n
        else:
            pspec, _ = buf(units * blk // 4)

            def both():
                return m.mufft_execute_conv_input_batch(plan, 0, pspec, pin, units) | m.mufft_execute_conv_output_batch(plan, pout, pspec, pf, units, 0)
            ms = timed(both, reps)
            nbytes = 28.0 * n
        xc, yc = x.view(np.complex64).reshape(units, half), y.view(np.complex64).reshape(units, n)
        err = 0.0
        for b in {0, units - 1}:
            want = np.zeros(n, np.complex128)
            want[: n - 1] = np.convolve(xc[b].astype(np.complex128), taps.astype(np.float64))
            err = max(err, rel_l2(yc[b], want))
        m.mufft_free_plan_conv(plan)
        flops, tol = 2 * 5.0 * n * math.log2(n) + 6.0 * n, 2e-6 * math.log2(n)
    else:
        raise SystemExit(f"unknown workload {name}")
    for p in frees:
        m.mufft_free(p)
    gbs = nbytes * units / ms / 1e6
    return {"workload": name, "units": units, "ms": round(ms, 4), "gbs": round(gbs, 1), "frac": round(gbs / PEAK_GBS, 3),
            "gflops": round(flops * units / ms / 1e6, 0), "rel_l2": err, "ok": bool(err <= tol)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="+")
    ap.add_argument("--label", default="")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    assert m.mufft_cuda_device_count() >= 1 and m.mufft_cuda_set_device(0) == 0
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    dry = os.environ.get("MUFFT_B200_LIB", "").endswith("libmufft_hostemu.so")  # the emulated host layer: not a measurement
    with open(os.devnull if dry else os.path.join(ROOT, "gpurun_out", "fastbench.jsonl"), "a") as log:
        for w in args.workloads:
            name, _, u = w.partition(":")
            res = run(name, int(u) if u else 0, args.reps)
            res["label"] = args.label
            line = json.dumps(res)
            print(line, flush=True)
            log.write(line + "\n")
            log.flush()
