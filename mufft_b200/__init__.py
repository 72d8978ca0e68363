"""muFFT for B200: Python mirror of the C API in include/fft.h and include/mufft_cuda.h.

The product is ``libmufft_b200.so`` (C host layer + hand-written sm_100a CUDA kernels, built from
``mufft_b200/csrc`` by ``make`` or ``__graft_entry__.build()``).  This package only binds it with
ctypes, one Python function per C entry point, same names and argument order, so tests read like
the reference's test.c.  There is no fallback: importing fails loudly when the shared library has
not been built, and plan creation returns NULL/None when no CUDA device is usable.
"""
from .api import *  # noqa: F401,F403
from .api import lib, LIB_PATH  # noqa: F401
