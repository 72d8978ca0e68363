"""The reference's OWN executables as acceptance tools.

oracle/Makefile `consumers` compiles the reference's test.c and bench.c UNCHANGED (from
/root/reference, in the build container) and links them against libmufft_b200.so; the only thing
they lack, FFTW3f as ground truth, is supplied by oracle/fftw3_shim (float64).  The binaries are
prebuilt artefacts under oracle/_ref/ (git-ignored, shipped to the GPU box); where they are absent
the tests skip.

* CPU:  muFFT-test-ref = test.c + shim + the reference library: proves the shim is a faithful
        stand-in (the reference passes its own sweep against it).
* GPU:  muFFT-test = test.c + shim + OUR library: every size/flag combination of ref test.c:493-579
        (1D C2C 2..65536, R2C/C2R incl. zero padding and FULL_R2C symmetry, mono/stereo
        convolution vs the O(N^2) sum, 2D C2C/R2C/C2R up to 512x512; CPU flags 0..7 are accepted and
        ignored) with the reference's own thresholds (1e-6 sqrt N absolute, test.c:58,195,431).
        muFFT-bench = bench.c against our library, one 1D and one 2D invocation.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _binary(name):
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        pytest.skip(f"oracle/_ref/{name} not built (needs /root/reference: `make -C oracle consumers`)")
    return path


def test_shim_is_a_valid_ground_truth_for_the_reference():
    res = subprocess.run([_binary("muFFT-test-ref")], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.rstrip().endswith("All tests passed!")
    assert res.stdout.count("... Passed") > 2500


@pytest.mark.gpu
def test_reference_test_suite_passes_against_libmufft_b200():
    res = subprocess.run([_binary("muFFT-test")], capture_output=True, text=True, timeout=1500)
    tail = res.stdout[-1500:] + res.stderr[-1500:]
    assert res.returncode == 0, f"muFFT-test aborted (rc={res.returncode}) after:\n{tail}"
    assert res.stdout.rstrip().endswith("All tests passed!")
    assert res.stdout.count("... Passed") > 2500


@pytest.mark.gpu
@pytest.mark.parametrize("argv", [["2000", "64"], ["200", "4096"], ["20", "256", "256"]], ids=["1d-64", "1d-4096", "2d-256"])
def test_reference_bench_runs_against_libmufft_b200(argv):
    """bench.c has no pass/fail; it must run to completion and print a muFFT line per benchmark
    (bench.c:430-583).  Its numbers are per-call latencies (one transform, one launch + sync per
    call), not the batched throughput bench.py reports."""
    res = subprocess.run([_binary("muFFT-bench"), *argv], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    lines = [line for line in res.stdout.splitlines() if line.startswith("muFFT")]
    assert len(lines) >= 3, res.stdout
    print("\n".join(lines))
