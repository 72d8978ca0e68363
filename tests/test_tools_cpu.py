"""Dry run of the GPU measurement tool tools/fastbench.py without a GPU: MUFFT_B200_LIB points the ctypes
binding at the C host layer on the emulated CUDA side (tests/emu: libmufft_hostemu.so exports the same fft.h /
mufft_cuda.h surface), sizes are tiny.  Only the tool's own logic and its parity checks are under test here; the
times it prints mean nothing."""
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
                          "c2c64:8", "r2c64:8", "c2r64:8", "conv64:4", "conv64_fused:4", "c2c2d256:2", "r2c2d512x256:2", "c2r2d512x256:3"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = json_lines(res.stdout)
    assert [d["workload"] for d in lines] == ["c2c64", "r2c64", "c2r64", "conv64", "conv64_fused", "c2c2d256", "r2c2d512x256", "c2r2d512x256"]
    assert all(d["ok"] and d["label"] == "dry" for d in lines), lines
