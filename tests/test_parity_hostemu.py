"""The GPU parity tests of tests/test_parity_gpu.py, unchanged, run a second time without a GPU: their
`gpu` fixture is replaced by the C host layer on the emulated CUDA side (tests/emu:
libmufft_hostemu.so).  Same inputs (the reference tests' srand(0) sequences), same oracle, same
tolerances - only the device is emulated.  Test infrastructure; see tests/test_hostemu.py."""
import os
import subprocess

import pytest

import cpu_libs
import test_parity_gpu as P

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


@pytest.fixture(scope="module")
def gpu():
    subprocess.run(["make", "-s", "-C", EMU_DIR, "-j", str(min(8, os.cpu_count() or 1))], check=True)
    return cpu_libs._CpuLib(os.path.join(EMU_DIR, "_build", "libmufft_hostemu.so"), "mufft_")


test_c2c_1d = P.test_c2c_1d
test_c2c_1d_zero_pad = P.test_c2c_1d_zero_pad
test_r2c_1d = P.test_r2c_1d
test_c2r_1d = P.test_c2r_1d
test_conv = P.test_conv
test_c2c_2d = P.test_c2c_2d
test_r2c_2d = P.test_r2c_2d
test_c2r_2d = P.test_c2r_2d
test_constructor_errors = P.test_constructor_errors
test_flags_cpu_bits_ignored = P.test_flags_cpu_bits_ignored
