"""Dry runs of the GPU measurement tools (tools/fastbench.py, tools/probe_optin.py - what bench.py starts as
child processes for also.optin_probes) without a GPU: MUFFT_B200_LIB points the ctypes binding at the C host
layer on the emulated CUDA side (tests/emu: libmufft_hostemu.so exports the same fft.h / mufft_cuda.h
surface), sizes are tiny.  Only the tools' own logic and their parity checks are under test here; the times
they print mean nothing."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def env():
    subprocess.run(["make", "-s", "-C", EMU_DIR, "-j", str(min(8, os.cpu_count() or 1))], check=True)
    return dict(os.environ, MUFFT_B200_LIB=os.path.join(EMU_DIR, "_build", "libmufft_hostemu.so"))


def json_lines(text):
    return [json.loads(ln) for ln in text.splitlines() if ln.startswith("{")]


def test_fastbench_workloads(env):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fastbench.py"), "--label", "dry", "--reps", "1",
                          "c2c64:8", "r2c64:8", "c2r64:8", "conv64:4", "conv64_fused:4", "c2c2d256:2"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = json_lines(res.stdout)
    assert [d["workload"] for d in lines] == ["c2c64", "r2c64", "c2r64", "conv64", "conv64_fused", "c2c2d256"]
    assert all(d["ok"] and d["label"] == "dry" for d in lines), lines


def test_probe_2d_and_the_knobs_it_turns(env):
    code = (
        "import sys; sys.argv = ['probe_optin.py', 'dry']; sys.path.insert(0, 'tools')\n"
        "import probe_optin as p\n"
        "p.probe_2d(256, 256, 2, modes=(1, 9, 11, 12, 16), tag='_dry')\n"
        "p.m.mufft_cuda_set_fused_2d_lag(2); p.probe_2d(256, 256, 3, modes=(1,), tag='_lag2'); p.m.mufft_cuda_set_fused_2d_lag(0)\n"
        "p.probe_partitioned(64, 3, 1)\n"
        "p.probe_managed(cases=(('c2c1d', 1024, 1), ('c2c2d', 256, 256)), rounds=2)\n")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = json_lines(res.stdout)
    assert len(lines) == 7 + 12 and all(d["ok"] for d in lines), lines
    assert [d["prefetch"] for d in lines[7:]] == [0, 1, 2, 0, 1, 2] * 2
    assert [d.get("fused_mode") for d in lines[:5]] == [1, 9, 11, 12, 16]


def test_probe_real2d_latency_cluster(env):
    """The probes bench.py starts first at round end: 2D R2C with and without the fused kernel (512 reals x 256
    rows = a 256 x 256 complex shape the emulation has), single-call latency with both synchronise flavours,
    32768-point rows on the two-kernel split and on both cluster flavours."""
    code = (
        "import sys; sys.argv = ['probe_optin.py', 'dry']; sys.path.insert(0, 'tools')\n"
        "import probe_optin as p\n"
        "p.REPS = 1\n"
        "p.probe_real2d(shapes=((512, 256),), bytes_in=1)\n"
        "p.probe_latency(iters=20, warm=2)\n"
        "p.probe_cluster(sizes=(32768,), bytes_in=1, flavours=(0, 1, 2))\n")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1500, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = json_lines(res.stdout)
    assert [d["what"] for d in lines] == ["r2c2d"] * 4 + ["single_call_latency"] * 4 + ["c2c1d_cluster"] * 6
    assert all(d["ok"] for d in lines), lines
    assert [d["launches_per_call"] for d in lines[:4]] == [2.0, 1.0, 2.0, 1.0]          # fused real 2D is one launch
    assert [d["launches_per_call"] for d in lines[8:]] == [2.0, 2.0, 1.0, 1.0, 1.0, 1.0]  # the cluster kernel replaces two
