#!/usr/bin/env python
"""Benchmark of the B200 muFFT hot path (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A step is one pass of the hot path over one batch of synthetic input.  The default workload is
BASELINE.json configs[1]: batched 1D C2C forward, N=4096, batch 262144 PER GPU (weak scaling: the
batch is sharded by transform, no collective on the data path).  Rank 0 prints ONE JSON line.

  value     GFLOP/s (muFFT's 5*N*log2 N convention) with inputs resident in HBM, CUDA-event timed
  e2e       the same metric through the host-pointer C ABI (mufft_execute_plan_1d_batch on pinned
            host buffers): H2D and D2H copies are inside the timed region
  roofline  algorithmic bytes (16*N per transform) / kernel time vs the measured HBM copy peak
  cpu_baseline / --impl reference
            the UNMODIFIED reference (oracle/_ref/ref_bench, all host cores, one plan per thread)
            on a bounded sample of the same workload; falls back to the C oracle port when the
            reference build is absent.  This is the only place oracle/ is executed by bench.py.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched FFT GFLOP/s (5N log2N)"
UNIT = "GFLOP/s"

# name -> (kind, N (Nx), Ny, units per GPU)
WORKLOADS = {
    "c2c4096": ("c2c", 4096, 0, 262144),
    "r2c4096": ("r2c", 4096, 0, 262144),
    "c2r4096": ("c2r", 4096, 0, 262144),
    "conv8192": ("conv_stereo", 8192, 0, 65536),
    "conv8192_fused": ("conv_stereo_fused", 8192, 0, 65536),
    "conv8192_part8": ("conv_stereo_part8", 8192, 0, 32768),
    "c2c2d1024": ("c2c2d", 1024, 1024, 256),
    "c2c64": ("c2c", 64, 0, 1 << 24),
    "c2c256": ("c2c", 256, 0, 1 << 22),
    "c2c1024": ("c2c", 1024, 0, 1 << 20),
    "c2c2048": ("c2c", 2048, 0, 1 << 19),
    "c2c8192": ("c2c", 8192, 0, 1 << 17),
    "r2c2048": ("r2c", 2048, 0, 1 << 19),
    "c2r2048": ("c2r", 2048, 0, 1 << 19),
    "r2c8192": ("r2c", 8192, 0, 1 << 17),
    "c2r8192": ("c2r", 8192, 0, 1 << 17),
    "r2c16384": ("r2c", 16384, 0, 1 << 16),
    "c2r16384": ("c2r", 16384, 0, 1 << 16),
    "r2c2d1024": ("r2c2d", 1024, 1024, 256),
    "c2r2d1024": ("c2r2d", 1024, 1024, 256),
    "c2c2d512": ("c2c2d", 512, 512, 1024),
    "c2c2d2048": ("c2c2d", 2048, 2048, 64),
    "c2c16384": ("c2c", 16384, 0, 32768),
    "c2c32768": ("c2c", 32768, 0, 16384),
    "c2c1m": ("c2c", 1 << 20, 0, 512),
    "slab2d": ("slab2d", 32768, 32768, 1),
    "slab2d8k": ("slab2d", 8192, 8192, 1),
    "slab2d_r2c": ("slabr2d", 32768, 32768, 1),
    "slab2d8k_r2c": ("slabr2d", 8192, 8192, 1),
}
DESCRIPTION = {
    "c2c4096": "batched 1D C2C forward N=4096, batch 262144 per GPU (BASELINE configs[1])",
    "r2c4096": "batched 1D R2C N=4096, batch 262144 per GPU (BASELINE configs[1])",
    "c2r4096": "batched 1D C2R N=4096, batch 262144 per GPU (BASELINE configs[1])",
    "conv8192": "stereo fast convolution FFT N=8192, 65536 blocks per GPU, one filter (BASELINE configs[2])",
    "conv8192_fused": "stereo fast convolution FFT N=8192, 65536 blocks per GPU, one filter, whole convolution in one kernel "
                      "(mufft_execute_conv_batch extension: 12N bytes per block instead of the API sequence's 28N)",
    "conv8192_part8": "stereo fast convolution FFT N=8192 with a LARGE filter (8 partitions = 32768 taps), 32768 blocks per GPU: "
                      "conv_input per block + one partitioned conv_output (8 spectral products summed in front of one inverse FFT)",
    "c2c2d1024": "batched 2D C2C 1024x1024, batch 256 per GPU (BASELINE configs[3])",
    "c2c64": "batched 1D C2C forward N=64, batch 2^24 per GPU (batched form of BASELINE configs[0])",
    "c2c16384": "batched 1D C2C forward N=16384, batch 32768 per GPU (longest single-kernel length: one 512-thread CTA per transform)",
    "c2c32768": "batched 1D C2C forward N=32768, batch 16384 per GPU (single-launch four-step kernel, intermediate in L2; row length of BASELINE configs[4])",
    "c2c256": "batched 1D C2C forward N=256, batch 2^22 per GPU",
    "c2c1024": "batched 1D C2C forward N=1024, batch 2^20 per GPU",
    "c2c2048": "batched 1D C2C forward N=2048, batch 2^19 per GPU",
    "c2c8192": "batched 1D C2C forward N=8192, batch 2^17 per GPU",
    "r2c2048": "batched 1D R2C N=2048, batch 2^19 per GPU",
    "c2r2048": "batched 1D C2R N=2048, batch 2^19 per GPU",
    "r2c8192": "batched 1D R2C N=8192, batch 2^17 per GPU",
    "c2r8192": "batched 1D C2R N=8192, batch 2^17 per GPU",
    "r2c16384": "batched 1D R2C N=16384, batch 2^16 per GPU",
    "c2r16384": "batched 1D C2R N=16384, batch 2^16 per GPU",
    "r2c2d1024": "batched 2D R2C 1024x1024, batch 256 per GPU",
    "c2r2d1024": "batched 2D C2R 1024x1024, batch 256 per GPU",
    "c2c2d512": "batched 2D C2C 512x512, batch 1024 per GPU",
    "c2c2d2048": "batched 2D C2C 2048x2048, batch 64 per GPU",
    "c2c1m": "batched 1D C2C forward N=2^20, batch 512 per GPU",
    "slab2d": "single 2D C2C 32768x32768, row slabs over the GPUs, all-to-all transpose over NVLink, transposed-out (BASELINE configs[4])",
    "slab2d8k": "single 2D C2C 8192x8192, row slabs over the GPUs, all-to-all transpose, transposed-out",
    "slab2d_r2c": "single 2D R2C 32768x32768, real row slabs over the GPUs, R2C rows first (half the exchange volume), all-to-all transpose over NVLink, transposed-out (BASELINE configs[4])",
    "slab2d8k_r2c": "single 2D R2C 8192x8192, real row slabs over the GPUs, all-to-all transpose, transposed-out",
}


def flops_per_unit(kind, n, ny):
    lg = math.log2(n)
    if kind in ("c2c2d", "slab2d"):
        return 5.0 * n * ny * (lg + math.log2(ny))
    if kind in ("r2c2d", "c2r2d", "slabr2d"):
        return 2.5 * n * ny * (lg + math.log2(ny))  # bench.c:546,567
    if kind in ("conv_stereo", "conv_stereo_fused"):
        return 2 * 5.0 * n * lg + 6.0 * n
    if kind == "conv_stereo_part8":
        return 2 * 5.0 * n * lg + 8 * 8.0 * n  # two transforms + 8 complex multiply-adds per bin
    if kind in ("r2c", "c2r"):
        return 2.5 * n * lg  # bench.c:546,567 convention for real transforms
    return 5.0 * n * lg


def bytes_per_unit(kind, n, ny):
    """Algorithmic HBM bytes per unit."""
    if kind == "c2c":
        return 16 * n
    if kind in ("r2c", "c2r"):
        return 4 * n + 8 * (n // 2 + 1)
    if kind == "conv_stereo":
        return 28 * n  # conv_input (4N in, 8N out) + conv_output (8N + 8N), filter spectrum amortised
    if kind == "conv_stereo_part8":
        return 28 * n  # as conv_stereo: every block spectrum is written once and read from HBM once (its 7 re-reads hit L2)
    if kind == "conv_stereo_fused":
        return 12 * n  # zero-padded block in (4N), convolved block out (8N); the spectrum stays on chip
    if kind in ("c2c2d", "slab2d"):
        return 16 * n * ny  # compulsory; the row->column intermediate is meant to stay in L2
    if kind in ("r2c2d", "c2r2d", "slabr2d"):
        return 4 * n * ny + 8 * (n // 2 + 1) * ny  # real image + the N/2+1 meaningful columns of the spectrum
    raise ValueError(kind)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


# ---- clocks sampler -----------------------------------------------------------------------------
class Clocks:
    """SM clock and throttle reasons sampled DURING the timed region.  The timed region of the
    default run is ~50 ms, far shorter than nvidia-smi's start-up, so the sampler is an NVML thread
    (the same counters nvidia-smi prints: clocks.sm, clocks.max.sm, clocks_event_reasons.*) polling
    every 2 ms; `nvidia-smi -lms` is the fallback when the NVML binding is missing."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    # nvml.h nvmlClocksEventReasons bit masks
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.samples = []  # (sm_mhz, reasons bitmask, power_w)
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.proc = None
        self.path = None
        self.source = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        try:
                            mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        except Exception:
                            mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        try:
                            watts = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        except Exception:
                            watts = None
                        self.samples.append((mhz, mask, watts))
                    except Exception:
                        pass
                    time.sleep(0.002)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.source = "nvml thread, 2 ms period"
            return
        except Exception:
            self.thread = None
        self.path = tempfile.mktemp(prefix="mufft_clocks_", suffix=".csv")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            self.source = "nvidia-smi -lms 20"
            # nvidia-smi needs a moment before its first sample; do not start timing without it
            t_end = time.time() + 5.0
            while time.time() < t_end and os.path.getsize(self.path) == 0:
                time.sleep(0.05)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm, mx, reasons, watts = [], [], set(), []
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            for mhz, mask, w in self.samples:
                sm.append(mhz)
                if w is not None:
                    watts.append(w)
                for name, bit in self.REASONS:
                    if mask & bit:
                        reasons.add(name)
            mx = [self.max_mhz]
        elif self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            names = [n for n, _ in self.REASONS]
            try:
                for line in open(self.path):
                    parts = [p.strip() for p in line.split(",")]
                    if len(parts) < 7:
                        continue
                    try:
                        sm.append(float(parts[0]))
                        mx.append(float(parts[1]))
                        watts.append(float(parts[2]))
                    except ValueError:
                        continue
                    for name, val in zip(names, parts[3:7]):
                        if val == "Active":
                            reasons.add(name)
                os.unlink(self.path)
            except OSError:
                pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(watts) if watts else None, source=self.source)
        return out


# ---- CPU reference ---------------------------------------------------------------------------------
REF_BENCH = os.path.join(ROOT, "oracle", "_ref", "ref_bench")
REF_MODE = {"c2c": "c2c", "r2c": "r2c_c2r", "c2r": "r2c_c2r", "conv_stereo": "conv_stereo", "conv_stereo_fused": "conv_stereo", "conv_stereo_part8": "conv_stereo", "c2c2d": "c2c2d"}


def cpu_reference_sample(kind, n, ny, seconds_of_cpu_work):
    """Times the reference CPU implementation on a bounded sample of the workload with every host
    core (one plan + private buffers per thread, transforms not cache-hot: each thread walks its
    own shard like the GPU walks the batch).  Returns the cpu_baseline object."""
    cores = os.cpu_count() or 1
    est_us = flops_per_unit(kind, n, ny) * (2 if kind in ("r2c", "c2r") else 1) / 15e3  # ~15 GFLOP/s per core, out of cache
    est_us = max(est_us, 0.05)
    if kind == "c2c2d":
        per_thread = 4
    else:
        per_thread = max(16, min(1 << 16, (64 << 20) // (16 * n)))  # ~64 MiB of in+out per thread
    batch = per_thread * cores
    repeats = max(1, int(seconds_of_cpu_work * 1e6 / (est_us * batch)))
    if os.path.exists(REF_BENCH):
        res = subprocess.run([REF_BENCH, REF_MODE[kind], str(n), str(ny), str(batch), str(cores), str(repeats)],
                             capture_output=True, text=True, check=True)
        r = json.loads(res.stdout.strip().splitlines()[-1])
        units_per_s = r["transforms_per_s"]
        seconds = r["seconds"]
        if kind in ("r2c", "c2r"):
            # ref_bench times R2C+C2R pairs (bench.c:312-344): a pair is two real transforms
            gflops = units_per_s * 2 * flops_per_unit(kind, n, ny) * 1e-9
            units_per_s *= 2
        else:
            gflops = units_per_s * flops_per_unit(kind, n, ny) * 1e-9
        return {"value": gflops, "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"oracle/_ref/ref_bench {REF_MODE[kind]}: {batch} units ({per_thread}/thread, private buffers) x {repeats} sweeps, {seconds:.2f} s wall",
                "units_per_s": units_per_s, "seconds": seconds}
    return cpu_oracle_sample(kind, n, ny, cores, per_thread, repeats)


def cpu_oracle_sample(kind, n, ny, cores, per_thread, repeats):
    """Fallback when oracle/_ref is absent: our C port (scalar) of the reference algorithm."""
    import numpy as np
    path = os.path.join(ROOT, "oracle", "libmufft_oracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    lib = ctypes.CDLL(path)
    lib.oracle_create_plan_1d_c2c.restype = ctypes.c_void_p
    lib.oracle_create_plan_1d_c2c.argtypes = [ctypes.c_uint, ctypes.c_int, ctypes.c_uint]
    lib.oracle_execute_plan_1d_many.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_size_t] * 3
    nn = n if kind != "c2c2d" else n  # the port baseline only models 1D complex passes of length n
    times = [0.0] * cores

    def work(t):
        plan = lib.oracle_create_plan_1d_c2c(nn, -1, 0)
        x = (np.random.default_rng(t).random((per_thread, 2 * nn), np.float32) - np.float32(0.5))
        y = np.zeros_like(x)
        t0 = time.perf_counter()
        for _ in range(repeats):
            lib.oracle_execute_plan_1d_many(plan, y.ctypes.data, x.ctypes.data, per_thread, 8 * nn, 8 * nn)
        times[t] = time.perf_counter() - t0

    threads = [threading.Thread(target=work, args=(t,)) for t in range(cores)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    seconds = max(times)
    units = per_thread * cores * repeats
    scale = {"c2c": 1.0, "r2c": 0.5, "c2r": 0.5, "conv_stereo": 2.0, "conv_stereo_fused": 2.0, "conv_stereo_part8": 2.0, "c2c2d": float(max(ny, 1)) * 2}[kind]
    units_per_s = units / seconds / scale
    return {"value": units_per_s * flops_per_unit(kind, n, ny) * 1e-9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port, {units} complex N={nn} transforms over {cores} threads, {seconds:.2f} s wall",
            "units_per_s": units_per_s, "seconds": seconds}


def run_reference_arm(args, rank):
    kind, n, ny, units = WORKLOADS[args.workload]
    if rank != 0:
        return
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_reference_sample(kind, n, ny, seconds_of_cpu_work=1.0 * (os.cpu_count() or 1))
        if i >= args.warmup:
            vals.append(base)
    total_s = sum(v["seconds"] for v in vals)
    value = sum(v["value"] * v["seconds"] for v in vals) / total_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / len(vals), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": DESCRIPTION[args.workload], "n": n, "ny": ny, "step": "bounded sample, see cpu_baseline.sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": base["cores"], "kind": base["kind"], "sample": base["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- the one JSON line ---------------------------------------------------------------------------------
MAX_LINE_BYTES = 4000  # the driver keeps only the tail of stdout: the line of record must stay far below that


def format_line(line):
    """JSON text of the line of record, guaranteed compact: floats rounded to 6 significant digits, and if it
    still exceeds MAX_LINE_BYTES the optional parts go first (`also` entries, longest first, then `also` itself)."""
    def rnd(o):
        if isinstance(o, float):
            return float(f"{o:.6g}")
        if isinstance(o, dict):
            return {k: rnd(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return [rnd(v) for v in o]
        return o
    line = rnd(line)
    text = json.dumps(line, separators=(",", ":"))
    also = line.get("also")
    while len(text) > MAX_LINE_BYTES and isinstance(also, dict) and also:
        also.pop(max(also, key=lambda k: len(json.dumps(also[k]))))
        line["also_truncated"] = True
        text = json.dumps(line, separators=(",", ":"))
    if len(text) > MAX_LINE_BYTES:
        line.pop("also", None)
        text = json.dumps(line, separators=(",", ":"))
    return text


# ---- GPU arm ---------------------------------------------------------------------------------------
class Job:
    """Device-resident buffers + the batched async C-ABI call for one workload."""

    def __init__(self, m, torch, kind, n, ny, units, dev):
        self.m, self.torch, self.kind, self.n, self.ny, self.units = m, torch, kind, n, ny, units
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234 + dev.index)

        def rnd(*shape):
            out = torch.empty(*shape, dtype=torch.float32, device=dev)
            step = max(1, shape[0] // 8)
            for lo in range(0, shape[0], step):  # chunked to bound the temporary
                out[lo:lo + step].uniform_(-0.5, 0.5, generator=gen)
            return out

        self.extra = []
        if kind == "c2c":
            self.plan = m.mufft_create_plan_1d_c2c(n, m.MUFFT_FORWARD, 0)
            self.inp = rnd(units, 2 * n)
            self.out = torch.empty_like(self.inp)
            self.call = lambda o, i, cnt: m.mufft_execute_plan_1d_batch_async(self.plan, o, i, cnt, 0, 0)
            self.call_sync = lambda o, i, cnt: m.mufft_execute_plan_1d_batch(self.plan, o, i, cnt, 0, 0)
            self.free = lambda: m.mufft_free_plan_1d(self.plan)
        elif kind == "r2c":
            # spectra rows: N/2+1 bins at a pitch of N/2+4 complex (16-byte aligned rows; the same
            # padded pitch the CPU reference driver oracle/ref_bench.c uses)
            pitch = 8 * (n // 2 + 4)
            self.plan = m.mufft_create_plan_1d_r2c(n, 0)
            self.inp = rnd(units, n)
            self.out = torch.empty(units, pitch // 4, dtype=torch.float32, device=dev)
            self.call = lambda o, i, cnt: m.mufft_execute_plan_1d_batch_async(self.plan, o, i, cnt, pitch, 0)
            self.call_sync = lambda o, i, cnt: m.mufft_execute_plan_1d_batch(self.plan, o, i, cnt, pitch, 0)
            self.free = lambda: m.mufft_free_plan_1d(self.plan)
        elif kind == "c2r":
            pitch = 8 * (n // 2 + 4)
            self.plan = m.mufft_create_plan_1d_c2r(n, 0)
            self.inp = rnd(units, pitch // 4)
            self.out = torch.empty(units, n, dtype=torch.float32, device=dev)
            self.call = lambda o, i, cnt: m.mufft_execute_plan_1d_batch_async(self.plan, o, i, cnt, 0, pitch)
            self.call_sync = lambda o, i, cnt: m.mufft_execute_plan_1d_batch(self.plan, o, i, cnt, 0, pitch)
            self.free = lambda: m.mufft_free_plan_1d(self.plan)
        elif kind == "conv_stereo":
            method = (m.MUFFT_CONV_METHOD_FLAG_STEREO_MONO | m.MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_FIRST |
                      m.MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_SECOND)
            self.plan = m.mufft_create_plan_conv(n, 0, method)
            blk = m.mufft_conv_get_transformed_block_size(self.plan) // 4
            self.inp = rnd(units, n)                      # n/2 stereo samples per block, zero padded to n
            filt = rnd(1, n // 2)
            self.fspec = torch.empty(blk, dtype=torch.float32, device=dev)
            self.spec = torch.empty(units, blk, dtype=torch.float32, device=dev)
            self.out = torch.empty(units, 2 * n, dtype=torch.float32, device=dev)
            assert m.mufft_execute_conv_input_batch(self.plan, 1, self.fspec.data_ptr(), filt.data_ptr(), 1) == 0

            def call(o, i, cnt, sync=False):
                f_in = m.mufft_execute_conv_input_batch if sync else m.mufft_execute_conv_input_batch_async
                f_out = m.mufft_execute_conv_output_batch if sync else m.mufft_execute_conv_output_batch_async
                return (f_in(self.plan, 0, self.spec.data_ptr(), i, cnt) or
                        f_out(self.plan, o, self.spec.data_ptr(), self.fspec.data_ptr(), cnt, 0))
            self.call = call
            self.call_sync = lambda o, i, cnt: call(o, i, cnt, True)
            self.free = lambda: m.mufft_free_plan_conv(self.plan)
        elif kind == "conv_stereo_part8":
            parts = 8
            method = (m.MUFFT_CONV_METHOD_FLAG_STEREO_MONO | m.MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_FIRST |
                      m.MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_SECOND)
            self.plan = m.mufft_create_plan_conv(n, 0, method)
            blk = m.mufft_conv_get_transformed_block_size(self.plan) // 4
            self.inp = rnd(units, n)
            taps = rnd(parts, n // 2)
            self.fspec = torch.empty(parts, blk, dtype=torch.float32, device=dev)
            self.spec = torch.zeros(units + parts - 1, blk, dtype=torch.float32, device=dev)  # parts-1 blocks of (zero) history in front
            self.out = torch.empty(units, 2 * n, dtype=torch.float32, device=dev)
            assert m.mufft_execute_conv_input_batch(self.plan, 1, self.fspec.data_ptr(), taps.data_ptr(), parts) == 0
            first = self.spec.data_ptr() + (parts - 1) * blk * 4

            def call(o, i, cnt, sync=False):
                f_in = m.mufft_execute_conv_input_batch if sync else m.mufft_execute_conv_input_batch_async
                f_out = m.mufft_execute_conv_output_partitioned_batch if sync else m.mufft_execute_conv_output_partitioned_batch_async
                return f_in(self.plan, 0, first, i, cnt) or f_out(self.plan, o, first, self.fspec.data_ptr(), parts, cnt)
            self.call = call
            self.call_sync = lambda o, i, cnt: call(o, i, cnt, True)
            self.free = lambda: m.mufft_free_plan_conv(self.plan)
        elif kind == "conv_stereo_fused":
            method = (m.MUFFT_CONV_METHOD_FLAG_STEREO_MONO | m.MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_FIRST |
                      m.MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_SECOND)
            self.plan = m.mufft_create_plan_conv(n, 0, method)
            blk = m.mufft_conv_get_transformed_block_size(self.plan) // 4
            self.inp = rnd(units, n)
            filt = rnd(1, n // 2)
            self.fspec = torch.empty(blk, dtype=torch.float32, device=dev)
            self.out = torch.empty(units, 2 * n, dtype=torch.float32, device=dev)
            assert m.mufft_execute_conv_input_batch(self.plan, 1, self.fspec.data_ptr(), filt.data_ptr(), 1) == 0
            self.call = lambda o, i, cnt: m.mufft_execute_conv_batch_async(self.plan, o, i, self.fspec.data_ptr(), cnt, 0)
            self.call_sync = lambda o, i, cnt: m.mufft_execute_conv_batch(self.plan, o, i, self.fspec.data_ptr(), cnt, 0)
            self.free = lambda: m.mufft_free_plan_conv(self.plan)
        elif kind == "slab2d":
            import torch.distributed as dist_mod
            from mufft_b200.slab import SlabFFT2D
            if not dist_mod.is_initialized():  # single GPU: a one-rank group
                fd, rdzv = tempfile.mkstemp(prefix="mufft_bench_rdzv_")  # a file store: no port to collide on
                os.close(fd)
                os.unlink(rdzv)
                dist_mod.init_process_group("nccl", init_method="file://" + rdzv, rank=0, world_size=1, device_id=dev)
            world = dist_mod.get_world_size()
            self.slab = SlabFFT2D(n, ny, m.MUFFT_FORWARD, device=dev, chunks=int(os.environ.get("MUFFT_SLAB_CHUNKS", "8")),
                                  exchange=os.environ.get("MUFFT_SLAB_EXCHANGE") or None)
            self.inp = torch.view_as_complex(rnd(ny // world, n * 2).view(ny // world, n, 2))
            self.out = torch.empty((ny, n // world), dtype=torch.complex64, device=dev)
            self.buf = self.slab.buffers()
            self.plan = 1
            self.units = 1.0 / world  # this rank's share of the one transform
            self.call = None
            self.call_sync = None
            self.free = self.slab.close
        elif kind == "slabr2d":
            import torch.distributed as dist_mod
            from mufft_b200.slab import SlabRFFT2D
            if not dist_mod.is_initialized():
                fd, rdzv = tempfile.mkstemp(prefix="mufft_bench_rdzv_")  # a file store: no port to collide on
                os.close(fd)
                os.unlink(rdzv)
                dist_mod.init_process_group("nccl", init_method="file://" + rdzv, rank=0, world_size=1, device_id=dev)
            world = dist_mod.get_world_size()
            # MUFFT_SLAB_R2C_EXCHANGE=p2p: the R2C row kernel scatters the half spectrum into the peers' slabs (opt-in)
            self.slab = SlabRFFT2D(n, ny, device=dev, chunks=int(os.environ.get("MUFFT_SLAB_CHUNKS", "8")),
                                   exchange=os.environ.get("MUFFT_SLAB_R2C_EXCHANGE") or None)
            self.inp = rnd(ny // world, n)
            self.out = torch.empty((ny, self.slab.cols_local), dtype=torch.complex64, device=dev)
            self.buf = self.slab.buffers()
            self.plan = 1
            self.units = 1.0 / world
            self.call = None
            self.call_sync = None
            self.free = self.slab.close
        elif kind in ("r2c2d", "c2r2d"):
            # spectrum images keep the reference's pitch of Nx complex per row (fft.h:238-257)
            if kind == "r2c2d":
                self.plan = m.mufft_create_plan_2d_r2c(n, ny, 0)
                self.inp = rnd(units, ny * n)
                self.out = torch.empty(units, ny * n * 2, dtype=torch.float32, device=dev)
            else:
                self.plan = m.mufft_create_plan_2d_c2r(n, ny, 0)
                self.inp = rnd(units, ny * n * 2)
                self.out = torch.empty(units, ny * n, dtype=torch.float32, device=dev)
            self.call = lambda o, i, cnt: m.mufft_execute_plan_2d_batch_async(self.plan, o, i, cnt)
            self.call_sync = lambda o, i, cnt: m.mufft_execute_plan_2d_batch(self.plan, o, i, cnt)
            self.free = lambda: m.mufft_free_plan_2d(self.plan)
        elif kind == "c2c2d":
            self.plan = m.mufft_create_plan_2d_c2c(n, ny, m.MUFFT_FORWARD, 0)
            self.inp = rnd(units, ny * n * 2)
            self.out = torch.empty_like(self.inp)
            self.call = lambda o, i, cnt: m.mufft_execute_plan_2d_batch_async(self.plan, o, i, cnt)
            self.call_sync = lambda o, i, cnt: m.mufft_execute_plan_2d_batch(self.plan, o, i, cnt)
            self.free = lambda: m.mufft_free_plan_2d(self.plan)
        else:
            raise ValueError(kind)
        if not self.plan:
            raise RuntimeError("plan creation failed: " + m.mufft_cuda_last_error().decode())

    def step(self):
        if self.kind == "slab2d":
            self.slab.execute(self.out, self.inp, buf=self.buf)
            return
        if self.kind == "slabr2d":
            self.slab.forward(self.out, self.inp, buf=self.buf)
            return
        err = self.call(self.out.data_ptr(), self.inp.data_ptr(), self.units)
        if err:
            raise RuntimeError("execute failed: " + self.m.mufft_cuda_last_error().decode())


def slab_also(m, ndev, nx=32768, ny=32768, reps=5, real=False):
    """One nx x ny C2C (real: R2C) transform over ndev devices of this process (include/mufft_cuda.h slab plan), data
    resident in the plan's device slabs; device-event time (max over devices), median of `reps` after 2 warm-up runs."""
    if real:
        plan = m.mufft_cuda_create_plan_2d_slab_r2c(nx, ny, ndev, None, 0)
    else:
        plan = m.mufft_cuda_create_plan_2d_slab_c2c(nx, ny, m.MUFFT_FORWARD, ndev, None, 0)
    if not plan:
        raise RuntimeError(m.mufft_cuda_last_error().decode())
    try:
        for g in range(ndev):  # any finite data: timing does not depend on values (zeros would do, but keep it honest)
            assert m.mufft_cuda_memset(m.mufft_cuda_slab_input(plan, g), 0x3c, m.mufft_cuda_slab_input_bytes(plan, g)) == 0
        times = []
        for i in range(2 + reps):
            if m.mufft_cuda_execute_plan_2d_slab(plan) != 0:
                raise RuntimeError(m.mufft_cuda_last_error().decode())
            if i >= 2:
                times.append((m.mufft_cuda_slab_last_ms(plan), m.mufft_cuda_slab_last_rows_ms(plan)))
        ms, rows_ms = sorted(times)[len(times) // 2]
    finally:
        m.mufft_cuda_free_plan_2d_slab(plan)
        m.mufft_cuda_set_device(int(os.environ.get("LOCAL_RANK", "0")))
    flops = (2.5 if real else 5.0) * nx * ny * (math.log2(nx) + math.log2(ny))
    per_gpu_out = (4.0 if real else 8.0) * nx * ny * (ndev - 1) / ndev / ndev  # bytes leaving each GPU over NVLink
    return {"workload": "single 2D %s %dx%d over %d GPUs, C-ABI slab plan, peer-store exchange, transposed-out" % ("R2C" if real else "C2C", nx, ny, ndev),
            "ms": round(ms, 4), "rows_and_exchange_ms": round(rows_ms, 4), "gflops": round(flops / ms * 1e-6, 1),
            "nvlink_gbs_per_gpu_per_dir_in_row_pass": round(per_gpu_out / rows_ms * 1e-6, 1) if rows_ms > 0 else None,
            "hbm_gbs_per_gpu_algorithmic": round((8.0 if real else 16.0) * nx * ny / ndev / ms * 1e-6, 1)}


def time_steps(torch, dist, job, warmup, steps):
    """CUDA-event timing on the launching stream; returns (total ms max over ranks, per-step ms list)."""
    for _ in range(warmup):
        job.step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for k in range(steps):
        job.step()
        ev[k + 1].record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    total = ev[0].elapsed_time(ev[steps])
    per = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    if dist is not None:
        t = torch.tensor([total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = float(t.item())
    return total, per


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2c4096", choices=sorted(WORKLOADS))
    ap.add_argument("--units", type=int, default=0, help="override units per GPU (debug; invalidates the headline)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--also", default="r2c4096,c2r4096,conv8192,conv8192_fused,c2c2d1024,r2c2d1024,c2c64,c2c8192,c2c16384,c2c32768",
                    help="comma separated extra workloads reported under 'also' (N=1 only); '' disables")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the muFFT B200 library has no CPU path "
                         "(use --impl reference for the CPU reference arm)")
    import mufft_b200 as m

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    assert m.mufft_cuda_set_device(local) == 0
    m.mufft_cuda_set_stream(torch.cuda.current_stream().cuda_stream or None)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod
    n_gpus = world

    kind, n, ny, units = WORKLOADS[args.workload]
    if args.units:
        units = args.units
    job = Job(m, torch, kind, n, ny, units, dev)
    if kind == "conv_stereo_part8":
        args.no_e2e = True  # the partitioned output stage takes device-accessible memory only
    if kind in ("slab2d", "slabr2d"):
        units = job.units
        args.no_e2e = True
        args.no_cpu = True  # the CPU reference for one 32768^2 image is a 33 s, 24 GiB single-thread job

    # Untimed, before the W warm-up steps of the contract: the first launches after the buffers were filled run
    # ~2 % slower than the steady state (r01c_bench.json: mean 2.725 ms against a median of 2.681 ms over 20
    # steps); a throughput number should not depend on how long the GPU idled while the inputs were generated.
    pre_warm = 10
    for _ in range(pre_warm):
        job.step()
    torch.cuda.synchronize()
    clocks = Clocks(local) if rank == 0 else None
    m.mufft_cuda_launch_count(1)
    total_ms, per = time_steps(torch, dist, job, args.warmup, args.steps)
    launches_total = m.mufft_cuda_launch_count(1)
    launches = launches_total * args.steps // (args.steps + args.warmup)
    clock_info = clocks.stop() if clocks else None

    ms_per_step = total_ms / args.steps
    fl, by = flops_per_unit(kind, n, ny), bytes_per_unit(kind, n, ny)
    value = n_gpus * units * fl / (ms_per_step * 1e-3) * 1e-9
    peak, peak_src = measured_peak()
    # dominant kernel: at N<=8192 a step is one fused kernel launch (two for conv / 2D), so the
    # kernel's mean duration is the step's event time divided by the launches it contains
    achieved = units * by / (ms_per_step * 1e-3) * 1e-9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(args.workload),
                "traffic_source": "profiles/traffic.json: ncu --set full capture of this kernel at this size, not re-measured in this run",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": units * by, "kernel_ms": ms_per_step,
                "median_step_ms": sorted(per)[len(per) // 2], "min_step_ms": min(per)}

    # ---- e2e: host-pointer C ABI, pinned host buffers, copies inside the timed region -----------
    e2e = None
    if not args.no_e2e:
        host_in = torch.empty(job.inp.shape, dtype=torch.float32, pin_memory=True)
        host_out = torch.empty(job.out.shape, dtype=torch.float32, pin_memory=True)
        host_in.copy_(job.inp)
        torch.cuda.synchronize()
        e2e_steps = max(1, min(args.steps, 5))
        for _ in range(1):
            assert job.call_sync(host_out.data_ptr(), host_in.data_ptr(), units) == 0, m.mufft_cuda_last_error()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            assert job.call_sync(host_out.data_ptr(), host_in.data_ptr(), units) == 0
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # result check of the host path against the device-resident path (same kernels, same input)
        job.step()
        torch.cuda.synchronize()
        probe = slice(0, min(units, 64))
        same = bool(torch.equal(host_out[probe].to(dev), job.out[probe]))
        e2e = {"value": n_gpus * units * fl / (dt / e2e_steps) * 1e-9, "unit": UNIT,
               "h2d_bytes_per_step": host_in.numel() * 4, "d2h_bytes_per_step": host_out.numel() * 4,
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "api": "mufft_execute_plan_1d_batch(host pinned in/out)" if kind in ("c2c", "r2c", "c2r") else "host-pointer batch API",
               "matches_device_path": same}
        # What the bus itself does on this box with the same pinned buffers (plain cudaMemcpyAsync through torch,
        # up to 1 GiB each way per GPU): the ceiling of any host-buffer path, reported next to the e2e number it
        # bounds.  At N > 1 every rank copies at the same time (barrier-aligned) and the aggregate is reported.
        try:
            rows = max(1, min(host_in.shape[0], host_out.shape[0], (1 << 30) // max(1, 4 * max(host_in.shape[1], host_out.shape[1]))))
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

            def copy_in():
                with torch.cuda.stream(s_in):
                    job.inp[:rows].copy_(host_in[:rows], non_blocking=True)

            def copy_out():
                with torch.cuda.stream(s_out):
                    host_out[:rows].copy_(job.out[:rows], non_blocking=True)

            def rate(fns, nbytes):
                best = 0.0
                for _ in range(3):
                    torch.cuda.synchronize()
                    if dist is not None:
                        dist.barrier()
                        torch.cuda.synchronize()
                    t1 = time.perf_counter()
                    for fn in fns:
                        fn()
                    torch.cuda.synchronize()
                    el = time.perf_counter() - t1
                    if dist is not None:
                        tt = torch.tensor([el], dtype=torch.float64, device="cuda")
                        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                        el = float(tt.item())
                    best = max(best, n_gpus * nbytes / el * 1e-9)
                return best
            b_in, b_out = rows * host_in.shape[1] * 4, rows * host_out.shape[1] * 4
            copy_in(); copy_out(); torch.cuda.synchronize()
            both = rate([copy_in, copy_out], b_in + b_out)
            e2e_gbs = n_gpus * (host_in.numel() + host_out.numel()) * 4 / (dt / e2e_steps) * 1e-9
            e2e["pcie"] = {"h2d_alone_gbs": round(rate([copy_in], b_in), 1), "d2h_alone_gbs": round(rate([copy_out], b_out), 1),
                           "both_ways_gbs": round(both, 1), "e2e_gbs": round(e2e_gbs, 1), "e2e_over_bus": round(e2e_gbs / both, 3),
                           "how": "aggregate over %d GPU(s): plain pinned<->device copies of %.2f GiB each way per GPU on two streams, all ranks at once, best of 3" % (n_gpus, b_in / 2**30)}
        except Exception as exc:
            e2e["pcie"] = {"error": repr(exc)[:200]}
        del host_in, host_out

    also = {}
    job.free()
    del job
    torch.cuda.empty_cache()
    if world == 1 and args.also and not args.units:
        for name in [w for w in args.also.split(",") if w and w != args.workload]:
            k2, n2, ny2, u2 = WORKLOADS[name]
            j2 = Job(m, torch, k2, n2, ny2, u2, dev)
            m.mufft_cuda_launch_count(1)
            t_ms, per2 = time_steps(torch, None, j2, 10, 10)
            l2 = m.mufft_cuda_launch_count(1) * 10 // 20
            ms2 = t_ms / 10
            gbs2 = u2 * bytes_per_unit(k2, n2, ny2) / (ms2 * 1e-3) * 1e-9
            also[name] = {"ms": round(ms2, 4), "gflops": round(u2 * flops_per_unit(k2, n2, ny2) / (ms2 * 1e-3) * 1e-9, 1),
                          "gbs": round(gbs2, 1), "frac": round(gbs2 / peak, 3), "launches": l2}
            j2.free()
            del j2
            torch.cuda.empty_cache()

    if world == 1 and args.also and not args.units and "c2c64" in args.also:
        # BASELINE configs[0] as the reference runs it (`muFFT-bench 1000000 64`): ONE 64-point transform
        # per mufft_execute_plan_1d call on mufft_alloc'd memory; each call is a launch + a sync
        plan = m.mufft_create_plan_1d_c2c(64, m.MUFFT_FORWARD, 0)
        a_in, a_out = m.mufft_calloc(64 * 8), m.mufft_alloc(64 * 8)
        for _ in range(200):
            m.mufft_execute_plan_1d(plan, a_out, a_in)
        iters = 5000
        t0 = time.perf_counter()
        for _ in range(iters):
            m.mufft_execute_plan_1d(plan, a_out, a_in)
        us = (time.perf_counter() - t0) / iters * 1e6
        # launch + synchronise per call; the batched form of the same transform is 'c2c64'
        also["c2c64_single_call"] = {"us_per_call": round(us, 3), "gflops": round(5 * 64 * 6 / us * 1e-3, 4)}
        m.mufft_free(a_in); m.mufft_free(a_out); m.mufft_free_plan_1d(plan)

    # BASELINE configs[4] next to the scaling curve: ONE 32768 x 32768 C2C over all N GPUs through the C-ABI slab plan
    # (mufft_cuda_create_plan_2d_slab_c2c: one process - rank 0 - drives every device; the row kernels' peer stores
    # over NVLink are the all-to-all).  The other ranks stay off their GPUs meanwhile (no NCCL kernel spinning there).
    if world > 1 and args.also and not args.units and args.workload == "c2c4096":
        flag = os.path.join(tempfile.gettempdir(), "mufft_slab_done_%s" % os.environ.get("MASTER_PORT", "0"))
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if rank == 0:
            for key, real in (("slab2d", False), ("slab2d_r2c", True)):
                try:
                    also[key] = slab_also(m, world, real=real)
                except Exception as exc:
                    also[key] = {"error": repr(exc)[:200]}
            open(flag, "w").close()
        else:
            t_end = time.time() + 300
            while not os.path.exists(flag) and time.time() < t_end:
                time.sleep(0.05)
        dist.barrier()
        if rank == 0:
            os.unlink(flag)

    cpu = None
    if rank == 0 and not args.no_cpu:
        # the other ranks wait in the barrier below; a bounded sample (~20 s of CPU work over all cores)
        cpu = cpu_reference_sample(kind, n, ny, seconds_of_cpu_work=20.0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if kind in ("slab2d", "slabr2d") else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": DESCRIPTION[args.workload], "n": n, "ny": ny, "units_per_gpu": units,
                       "global_units": units * n_gpus, "parallelism": f"batch-sharded x{n_gpus}, no collective",
                       "l2": "inputs+outputs per step (%.1f GiB per GPU) exceed the 126 MB L2; no flush needed" %
                             (units * by / 2**30),
                       "pre_warm_steps": pre_warm},
            "transforms_per_s": n_gpus * units / (ms_per_step * 1e-3),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clock_info, "also": also,
        }
        print(format_line(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    elif kind in ("slab2d", "slabr2d"):
        import torch.distributed as dist_mod
        if dist_mod.is_initialized():  # the one-rank group a single-GPU slab job created
            dist_mod.destroy_process_group()


if __name__ == "__main__":
    main()
