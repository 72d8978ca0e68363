"""ctypes binding of libmufft_b200.so.  Names, argument order and meaning follow include/fft.h
(= ref fft.h:44-297) and include/mufft_cuda.h; pointers are passed as integers (numpy
``arr.ctypes.data``, torch ``tensor.data_ptr()``) or anything ctypes converts to ``void *``.
"""
import ctypes as C
import os

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmufft_b200.so")
# A/B measurements only (tools/ab_build.sh builds the same sources with other compile-time choices into
# ab_variants/<label>/libmufft_b200.so): MUFFT_B200_LIB selects such a build of THIS library instead.
if os.environ.get("MUFFT_B200_LIB"):
    LIB_PATH = os.path.abspath(os.environ["MUFFT_B200_LIB"])
if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C mufft_b200/csrc -j` or "
        "`python -c 'import __graft_entry__ as g; g.build()'`. There is no Python or CPU fallback.")
lib = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL)

# ---- constants (ref fft.h:44-65, 131-150) --------------------------------------------------------
MUFFT_FORWARD = -1
MUFFT_INVERSE = 1
MUFFT_FLAG_CPU_ANY = 0
MUFFT_FLAG_CPU_NO_SIMD = (1 << 16) - 1
MUFFT_FLAG_CPU_NO_AVX = 1 << 0
MUFFT_FLAG_CPU_NO_SSE3 = 1 << 1
MUFFT_FLAG_CPU_NO_SSE = 1 << 2
MUFFT_FLAG_FULL_R2C = 1 << 16
MUFFT_FLAG_ZERO_PAD_UPPER_HALF = 1 << 17
MUFFT_CONV_METHOD_FLAG_MONO_MONO = 0
MUFFT_CONV_METHOD_FLAG_STEREO_MONO = 1
MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_FIRST = 1 << 2
MUFFT_CONV_METHOD_FLAG_ZERO_PAD_UPPER_HALF_SECOND = 1 << 3
MUFFT_CONV_BLOCK_FIRST = 0
MUFFT_CONV_BLOCK_SECOND = 1
# plan-creation flags of include/mufft_cuda.h (slower fallback paths, for measurement and tests)
MUFFT_CUDA_FLAG_SEPARATE_2D = 1 << 20
MUFFT_CUDA_FLAG_TWO_KERNEL = 1 << 21
MUFFT_CUDA_FLAG_GENERIC_PASSES = 1 << 22
MUFFT_CUDA_FLAG_FOUR_STEP_16K = 1 << 23

_vp, _u, _i, _sz = C.c_void_p, C.c_uint, C.c_int, C.c_size_t
mufft_convolve_func = C.CFUNCTYPE(None, _vp, _vp, _vp, C.c_float, _u)

# Every symbol declared by include/fft.h and include/mufft_cuda.h: (restype, argtypes).
SIGNATURES = {
    # fft.h
    "mufft_create_plan_1d_c2c": (_vp, [_u, _i, _u]),
    "mufft_create_plan_1d_r2c": (_vp, [_u, _u]),
    "mufft_create_plan_1d_c2r": (_vp, [_u, _u]),
    "mufft_execute_plan_1d": (None, [_vp, _vp, _vp]),
    "mufft_free_plan_1d": (None, [_vp]),
    "mufft_create_plan_conv": (_vp, [_u, _u, _u]),
    "mufft_execute_conv_input": (None, [_vp, _u, _vp, _vp]),
    "mufft_conv_get_transformed_block_size": (_sz, [_vp]),
    "mufft_execute_conv_output": (None, [_vp, _vp, _vp, _vp]),
    "mufft_free_plan_conv": (None, [_vp]),
    "mufft_get_convolve_func": (mufft_convolve_func, [_u]),
    "mufft_create_plan_2d_c2c": (_vp, [_u, _u, _i, _u]),
    "mufft_create_plan_2d_r2c": (_vp, [_u, _u, _u]),
    "mufft_create_plan_2d_c2r": (_vp, [_u, _u, _u]),
    "mufft_execute_plan_2d": (None, [_vp, _vp, _vp]),
    "mufft_free_plan_2d": (None, [_vp]),
    "mufft_alloc": (_vp, [_sz]),
    "mufft_calloc": (_vp, [_sz]),
    "mufft_free": (None, [_vp]),
    # mufft_cuda.h
    "mufft_cuda_device_count": (_i, []),
    "mufft_cuda_set_device": (_i, [_i]),
    "mufft_cuda_set_stream": (None, [_vp]),
    "mufft_cuda_get_stream": (_vp, []),
    "mufft_cuda_synchronize": (_i, []),
    "mufft_cuda_last_error": (C.c_char_p, []),
    "mufft_cuda_launch_count": (C.c_ulonglong, [_i]),
    "mufft_execute_plan_1d_batch": (_i, [_vp, _vp, _vp, _sz, _sz, _sz]),
    "mufft_execute_plan_1d_batch_async": (_i, [_vp, _vp, _vp, _sz, _sz, _sz]),
    "mufft_execute_plan_1d_columns": (_i, [_vp, _vp, _vp, _sz, _sz, _sz]),
    "mufft_execute_plan_1d_batch_scatter_async": (_i, [_vp, C.POINTER(C.c_void_p), _u, _sz, _vp, _sz, _sz]),
    "mufft_cuda_device_alloc": (_vp, [_sz]),
    "mufft_cuda_device_free": (None, [_vp]),
    "mufft_cuda_host_alloc": (_vp, [_sz]),
    "mufft_cuda_host_free": (None, [_vp]),
    "mufft_cuda_copy": (_i, [_vp, _vp, _sz]),
    "mufft_cuda_memset": (_i, [_vp, _i, _sz]),
    "mufft_cuda_ipc_export": (_i, [_vp, _vp]),
    "mufft_cuda_ipc_import": (_vp, [_vp]),
    "mufft_cuda_ipc_close": (_i, [_vp]),
    "mufft_execute_plan_2d_batch": (_i, [_vp, _vp, _vp, _sz]),
    "mufft_execute_plan_2d_batch_async": (_i, [_vp, _vp, _vp, _sz]),
    "mufft_execute_conv_input_batch": (_i, [_vp, _u, _vp, _vp, _sz]),
    "mufft_execute_conv_output_batch": (_i, [_vp, _vp, _vp, _vp, _sz, _sz]),
    "mufft_execute_conv_input_batch_async": (_i, [_vp, _u, _vp, _vp, _sz]),
    "mufft_execute_conv_output_batch_async": (_i, [_vp, _vp, _vp, _vp, _sz, _sz]),
    "mufft_execute_conv_batch": (_i, [_vp, _vp, _vp, _vp, _sz, _sz]),
    "mufft_execute_conv_batch_async": (_i, [_vp, _vp, _vp, _vp, _sz, _sz]),
    "mufft_execute_conv_output_partitioned_batch": (_i, [_vp, _vp, _vp, _vp, _u, _sz]),
    "mufft_execute_conv_output_partitioned_batch_async": (_i, [_vp, _vp, _vp, _vp, _u, _sz]),
    "mufft_cuda_create_plan_2d_slab_c2c": (_vp, [_u, _u, _i, _u, C.POINTER(_i), _u]),
    "mufft_cuda_create_plan_2d_slab_r2c": (_vp, [_u, _u, _u, C.POINTER(_i), _u]),
    "mufft_cuda_free_plan_2d_slab": (None, [_vp]),
    "mufft_cuda_slab_devices": (_u, [_vp]),
    "mufft_cuda_slab_input": (_vp, [_vp, _u]),
    "mufft_cuda_slab_output": (_vp, [_vp, _u]),
    "mufft_cuda_slab_input_bytes": (_sz, [_vp, _u]),
    "mufft_cuda_slab_output_bytes": (_sz, [_vp, _u]),
    "mufft_cuda_slab_output_columns": (_u, [_vp, _u, C.POINTER(_u)]),
    "mufft_cuda_execute_plan_2d_slab": (_i, [_vp]),
    "mufft_cuda_slab_last_ms": (C.c_float, [_vp]),
    "mufft_cuda_slab_last_rows_ms": (C.c_float, [_vp]),
    "mufft_cuda_execute_plan_2d_slab_host": (_i, [_vp, _vp, _vp]),
    "mufft_cuda_plan_1d_uses_engine": (_i, [_vp]),
    "mufft_cuda_plan_2d_paths": (_i, [_vp]),
}

__all__ = [n for n in dir() if n.startswith("MUFFT_")] + list(SIGNATURES) + ["mufft_convolve_func", "SIGNATURES", "ptr"]

for _name, (_res, _args) in SIGNATURES.items():
    try:
        _fn = getattr(lib, _name)  # AttributeError here = the library does not export a declared symbol
    except AttributeError:
        if not os.environ.get("MUFFT_B200_LIB"):
            raise
        __all__.remove(_name)  # an A/B build made before the symbol existed: the measurement tools do not need it
        continue
    _fn.restype, _fn.argtypes = _res, _args
    globals()[_name] = _fn


def ptr(obj):
    """void* of a numpy array, torch tensor or plain integer address."""
    if obj is None:
        return None
    if isinstance(obj, int):
        return obj
    if hasattr(obj, "data_ptr"):
        return obj.data_ptr()
    return obj.ctypes.data
