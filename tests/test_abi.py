"""CPU tests (no GPU): the C-ABI shared library loads, exports every symbol that include/fft.h and
include/mufft_cuda.h declare, and fails loudly - never falls back - when there is no CUDA device.
No compute entry point is called here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    return sorted(set(re.findall(r"\b(mufft_[a-z0-9_]+)\s*\(", text)) - {"mufft_convolve_func"})


@pytest.mark.parametrize("header", ["fft.h", "mufft_cuda.h"])
def test_library_exports_every_declared_symbol(header):
    import mufft_b200
    names = declared_functions(header)
    assert len(names) >= (19 if header == "fft.h" else 10)
    for name in names:
        assert hasattr(mufft_b200.lib, name), f"{name} declared in include/{header} but not exported"
        assert name in mufft_b200.SIGNATURES, f"{name} has no ctypes signature in mufft_b200/api.py"


def test_fft_h_surface_is_the_reference_surface():
    """The 19 entry points and the constants of ref fft.h:44-297."""
    names = declared_functions("fft.h")
    expected = """mufft_create_plan_1d_c2c mufft_create_plan_1d_r2c mufft_create_plan_1d_c2r mufft_execute_plan_1d
        mufft_free_plan_1d mufft_create_plan_conv mufft_execute_conv_input mufft_conv_get_transformed_block_size
        mufft_execute_conv_output mufft_free_plan_conv mufft_get_convolve_func mufft_create_plan_2d_c2c
        mufft_create_plan_2d_r2c mufft_create_plan_2d_c2r mufft_execute_plan_2d mufft_free_plan_2d mufft_alloc
        mufft_calloc mufft_free""".split()
    assert sorted(expected) == names
    text = open(os.path.join(ROOT, "include", "fft.h")).read()

    def macro(name):
        m = re.search(r"#define\s+%s\s+(.+)" % name, text)
        assert m, name
        return eval(m.group(1).split("/*")[0].strip().replace("u", ""))  # noqa: S307 - constant expressions only
    assert macro("MUFFT_FORWARD") == -1 and macro("MUFFT_INVERSE") == 1
    assert macro("MUFFT_FLAG_CPU_NO_SIMD") == 0xFFFF
    assert macro("MUFFT_FLAG_FULL_R2C") == 1 << 16 and macro("MUFFT_FLAG_ZERO_PAD_UPPER_HALF") == 1 << 17
    assert macro("MUFFT_CONV_METHOD_FLAG_STEREO_MONO") == 1 and macro("MUFFT_CONV_METHOD_FLAG_MONO_MONO") == 0
    assert macro("MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_FIRST") == 4
    assert macro("MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_SECOND") == 8


def test_python_constants_match_header():
    import mufft_b200 as m
    assert (m.MUFFT_FORWARD, m.MUFFT_INVERSE) == (-1, 1)
    assert m.MUFFT_FLAG_FULL_R2C == 1 << 16 and m.MUFFT_FLAG_ZERO_PAD_UPPER_HALF == 1 << 17


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device plan creation and allocation return NULL (and say why); they never
    compute on the CPU.  Argument validation (ref fft.c:628-631) still answers NULL first."""
    import mufft_b200 as m
    for n in (0, 1, 3, 100):
        assert not m.mufft_create_plan_1d_c2c(n, m.MUFFT_FORWARD, 0)
    assert not m.mufft_create_plan_1d_r2c(2, 0)
    if m.mufft_cuda_device_count() > 0:
        pytest.skip("a CUDA device is visible: the no-device behaviour cannot be observed here")
    assert not m.mufft_create_plan_1d_c2c(64, m.MUFFT_FORWARD, 0)
    assert b"no usable CUDA device" in m.mufft_cuda_last_error()
    assert not m.mufft_create_plan_2d_c2c(8, 8, m.MUFFT_FORWARD, 0)
    assert not m.mufft_create_plan_conv(64, 0, 0)
    assert not m.mufft_alloc(1024)
    assert not m.mufft_cuda_device_alloc(1024) and not m.mufft_cuda_host_alloc(1024)
    assert not m.mufft_cuda_create_plan_2d_slab_c2c(64, 64, m.MUFFT_FORWARD, 1, None, 0)
    assert not m.mufft_cuda_create_plan_2d_slab_r2c(64, 64, 1, None, 0)
    m.mufft_cuda_free_plan_2d_slab(None)
    m.mufft_cuda_device_free(None)
    m.mufft_cuda_host_free(None)
    m.mufft_free(None)
    m.mufft_free_plan_1d(None)
    m.mufft_free_plan_2d(None)
    m.mufft_free_plan_conv(None)


def test_product_does_not_link_the_oracle():
    """libmufft_b200.so must not depend on anything under oracle/ (nor on cuFFT)."""
    import subprocess
    import mufft_b200
    out = subprocess.run(["ldd", mufft_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "mufft_ref" not in out and "cufft" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", mufft_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in syms


def test_engine_twiddle_tables_match_double_precision():
    """Host-side table construction (the CUDA engine's analogue of ref fft.c:102-134): every entry
    is exp(dir*2*pi*i*k*m/(r*p)) evaluated in double and rounded once."""
    import mufft_b200
    lib = mufft_b200.lib
    lib.mcu_engine_table_size.restype = ctypes.c_size_t
    lib.mcu_engine_table_size.argtypes = [ctypes.c_uint]
    lib.mcu_engine_table_fill.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_int]
    for n in (16, 64, 256, 4096, 8192):
        size = lib.mcu_engine_table_size(n)
        assert size >= 1
        for direction in (-1, 1):
            t = np.zeros(size, np.complex64)
            lib.mcu_engine_table_fill(t.ctypes.data, n, direction)
            mag = np.abs(t[: size - 1].astype(np.complex128))
            assert np.all(np.abs(mag - 1.0) < 1e-6)
