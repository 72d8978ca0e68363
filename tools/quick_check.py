"""Parity spot check of the CUDA library against the oracle for one complex length N (both
directions, R2C/C2R of 2N, a 2D case with N columns/rows when small enough).  Prints OK or the failure."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cpu_libs  # noqa: E402
import mufft_b200  # noqa: E402

n = int(sys.argv[1])
gpu = cpu_libs._CpuLib(mufft_b200.LIB_PATH, "mufft_")
oracle = cpu_libs.load_oracle()
rng = np.random.default_rng(n)
worst = 0.0
x = (rng.random(n, np.float32) - 0.5 + 1j * (rng.random(n, np.float32) - 0.5)).astype(np.complex64)
for d in (-1, 1):
    worst = max(worst, cpu_libs.rel_l2(gpu.c2c(x, d), oracle.c2c(x, d)))
worst = max(worst, cpu_libs.rel_l2(gpu.c2c(x[: n // 2], -1, cpu_libs.FLAG_ZERO_PAD, n=n), oracle.c2c(x[: n // 2], -1, cpu_libs.FLAG_ZERO_PAD, n=n)))
r = rng.random(2 * n, np.float32) - np.float32(0.5)
for flags in (0, cpu_libs.FLAG_FULL_R2C):
    worst = max(worst, cpu_libs.rel_l2(gpu.r2c(r, flags), oracle.r2c(r, flags)))
spec = oracle.r2c(r)
worst = max(worst, cpu_libs.rel_l2(gpu.c2r(spec, 2 * n), oracle.c2r(spec, 2 * n)))
if n <= 2048:
    img = (rng.random((n, 32), np.float32) - 0.5 + 1j * (rng.random((n, 32), np.float32) - 0.5)).astype(np.complex64)
    worst = max(worst, cpu_libs.rel_l2(gpu.c2c_2d(img, -1), oracle.c2c_2d(img, -1)))
    worst = max(worst, cpu_libs.rel_l2(gpu.c2c_2d(img.T.copy(), 1), oracle.c2c_2d(img.T.copy(), 1)))
tol = 1e-6 * math.log2(max(n, 2))
print(("OK" if worst <= tol else "FAIL") + f" worst rel_l2 {worst:.2e} (tol {tol:.1e})")
