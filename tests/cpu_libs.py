"""ctypes bindings for the two CPU checkers used by the tests (never by the product):

* ``Oracle``  - oracle/libmufft_oracle.so, our plain-C restatement (symbols ``oracle_*``);
* ``Ref``     - oracle/_ref/libmufft_ref.so, the unmodified reference compiled by oracle/Makefile
                (symbols ``mufft_*``; present in the build container and shipped to the GPU box as
                a prebuilt file, absent from git).

Both expose the same method names as include/fft.h without the prefix, working on numpy arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

FORWARD, INVERSE = -1, 1
FLAG_NO_SIMD = (1 << 16) - 1
FLAG_FULL_R2C = 1 << 16
FLAG_ZERO_PAD = 1 << 17
CONV_MONO, CONV_STEREO = 0, 1
CONV_ZP_FIRST, CONV_ZP_SECOND = 1 << 2, 1 << 3


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def aligned_zeros(shape, dtype, align=64):
    """numpy array whose data pointer is `align`-byte aligned (the reference's SIMD kernels use
    aligned loads/stores, ref fft.h:115-116)."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    raw = np.zeros(count * dtype.itemsize + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + count * dtype.itemsize].view(dtype).reshape(shape)


def aligned_copy(x, dtype):
    x = np.asarray(x, dtype)
    out = aligned_zeros(x.shape, dtype)
    out[...] = x
    return out


class _CpuLib:
    """Thin object wrapper over a fft.h-shaped C library with symbol prefix ``prefix``."""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL)
        self.prefix = prefix
        f = self._f
        vp, u, i = C.c_void_p, C.c_uint, C.c_int
        f("create_plan_1d_c2c", vp, [u, i, u])
        f("create_plan_1d_r2c", vp, [u, u])
        f("create_plan_1d_c2r", vp, [u, u])
        f("execute_plan_1d", None, [vp, vp, vp])
        f("free_plan_1d", None, [vp])
        f("create_plan_conv", vp, [u, u, u])
        f("execute_conv_input", None, [vp, u, vp, vp])
        f("conv_get_transformed_block_size", C.c_size_t, [vp])
        f("execute_conv_output", None, [vp, vp, vp, vp])
        f("free_plan_conv", None, [vp])
        f("create_plan_2d_c2c", vp, [u, u, i, u])
        f("create_plan_2d_r2c", vp, [u, u, u])
        f("create_plan_2d_c2r", vp, [u, u, u])
        f("execute_plan_2d", None, [vp, vp, vp])
        f("free_plan_2d", None, [vp])

    def _f(self, name, restype, argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype, fn.argtypes = restype, argtypes
        setattr(self, "_" + name, fn)

    # ---- 1D -------------------------------------------------------------------------------
    def c2c(self, x, direction, flags=0, n=None):
        x = aligned_copy(x, np.complex64)
        n = n or x.size
        plan = self._create_plan_1d_c2c(n, direction, flags)
        assert plan, "plan creation failed"
        out = aligned_zeros(n, np.complex64)
        self._execute_plan_1d(plan, _ptr(out), _ptr(x))
        self._free_plan_1d(plan)
        return out

    def r2c(self, x, flags=0, n=None):
        x = aligned_copy(x, np.float32)
        n = n or x.size
        plan = self._create_plan_1d_r2c(n, flags)
        assert plan, "plan creation failed"
        out = aligned_zeros(n if flags & FLAG_FULL_R2C else n // 2 + 1, np.complex64)
        self._execute_plan_1d(plan, _ptr(out), _ptr(x))
        self._free_plan_1d(plan)
        return out

    def c2r(self, x, n, flags=0):
        x = aligned_copy(x, np.complex64)
        assert x.size >= n // 2 + 1
        plan = self._create_plan_1d_c2r(n, flags)
        assert plan, "plan creation failed"
        out = aligned_zeros(n, np.float32)
        self._execute_plan_1d(plan, _ptr(out), _ptr(x))
        self._free_plan_1d(plan)
        return out

    def plan_1d_is_null(self, kind, n, direction=FORWARD, flags=0):
        if kind == "c2c":
            p = self._create_plan_1d_c2c(n, direction, flags)
        elif kind == "r2c":
            p = self._create_plan_1d_r2c(n, flags)
        else:
            p = self._create_plan_1d_c2r(n, flags)
        if p:
            self._free_plan_1d(p)
        return not p

    # ---- convolution -------------------------------------------------------------------------
    def conv(self, a, b, n, method, flags=0):
        """a, b: first/second block inputs (already of the length the method expects)."""
        plan = self._create_plan_conv(n, flags, method)
        assert plan, "plan creation failed"
        size = self._conv_get_transformed_block_size(plan)
        stereo = method & 1
        a = aligned_copy(a, np.complex64 if stereo else np.float32)
        b = aligned_copy(b, np.float32)
        s0 = aligned_zeros(size // 8, np.complex64)
        s1 = aligned_zeros(size // 8, np.complex64)
        self._execute_conv_input(plan, 0, _ptr(s0), _ptr(a))
        self._execute_conv_input(plan, 1, _ptr(s1), _ptr(b))
        out = aligned_zeros(n, np.complex64 if stereo else np.float32)
        self._execute_conv_output(plan, _ptr(out), _ptr(s0), _ptr(s1))
        self._free_plan_conv(plan)
        return out, s0, s1

    # ---- 2D ----------------------------------------------------------------------------------
    def c2c_2d(self, x, direction, flags=0):
        x = aligned_copy(x, np.complex64)
        ny, nx = x.shape
        plan = self._create_plan_2d_c2c(nx, ny, direction, flags)
        assert plan, "plan creation failed"
        out = aligned_zeros((ny, nx), np.complex64)
        self._execute_plan_2d(plan, _ptr(out), _ptr(x))
        self._free_plan_2d(plan)
        return out

    def r2c_2d(self, x, flags=0):
        x = aligned_copy(x, np.float32)
        ny, nx = x.shape
        plan = self._create_plan_2d_r2c(nx, ny, flags)
        assert plan, "plan creation failed"
        out = aligned_zeros((ny, nx), np.complex64)  # pitch Nx complex (fft.h: 2D R2C layout)
        self._execute_plan_2d(plan, _ptr(out), _ptr(x))
        self._free_plan_2d(plan)
        return out

    def c2r_2d(self, x, flags=0):
        """x: [Ny][Nx] complex, pitch Nx, columns 0..Nx/2 meaningful."""
        x = aligned_copy(x, np.complex64)
        ny, nx = x.shape
        plan = self._create_plan_2d_c2r(nx, ny, flags)
        assert plan, "plan creation failed"
        out = aligned_zeros(2 * nx * ny, np.float32)  # output doubles as scratch
        self._execute_plan_2d(plan, _ptr(out), _ptr(x))
        self._free_plan_2d(plan)
        return out[: nx * ny].reshape(ny, nx).copy()


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)


def load_oracle():
    path = os.path.join(ORACLE_DIR, "libmufft_oracle.so")
    if not os.path.exists(path):
        build_oracle()
    return _CpuLib(path, "oracle_")


def ref_path():
    return os.path.join(ORACLE_DIR, "_ref", "libmufft_ref.so")


def load_ref():
    """The unmodified reference, or None when oracle/_ref has not been built."""
    path = ref_path()
    if not os.path.exists(path):
        if os.path.isdir("/root/reference"):
            subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)
        else:
            return None
    return _CpuLib(path, "mufft_")


def c_rand(count, seed=0):
    """Replica of the reference tests' generator: srand(seed); (float)rand()/RAND_MAX - 0.5f
    (test.c:38-44).  Uses glibc's rand() so the sequence is the one the reference sees."""
    libc = C.CDLL(None)
    libc.srand(seed)
    libc.rand.restype = C.c_int
    rmax = np.float32(2147483647.0)
    return np.array([np.float32(libc.rand()) / rmax - np.float32(0.5) for _ in range(count)], np.float32)


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128).ravel()
    b = np.asarray(b).astype(np.complex128).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))
