"""Offline model of the in-CTA Stockham engine (mufft_b200/csrc/fft_engine.cuh): checks the
index maps against numpy's FFT and counts shared-memory bank conflicts of every exchange for a
given (N, R, radices, padding) so that layouts can be chosen without spending GPU time.

Engine recap: thread t of a transform holds R values; register slot e <-> element t + e*T
(T = N/R threads per transform).  A pass of radix r at span p lets the thread do R/r
butterflies u: i = t + u*T, inputs are slots u + m*(R/r); output q goes to position
(i-k)*r + k + q*p with k = i mod p.  Between passes the outputs go through shared memory at
padded position pos + (pos >> pad_shift).
"""
import sys
import numpy as np


def config(n):
    """(R, radices) as chosen by the CUDA engine."""
    if n >= 256:
        R = 16
    elif n >= 64:
        R = 8
    elif n >= 16:
        R = 4
    else:
        R = 2
    radices = []
    rem = n
    while rem >= R:
        radices.append(R)
        rem //= R
    if rem > 1:
        radices.append(rem)
    return R, radices


def simulate(n, R, radices, sign=-1, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    T = n // R
    regs = np.array([[x[t + e * T] for e in range(R)] for t in range(T)])
    p = 1
    for r in radices:
        nb = R // r
        smem = np.zeros(n, complex)
        for t in range(T):
            for u in range(nb):
                i = t + u * T
                k = i % p
                vin = np.array([regs[t][u + m * nb] for m in range(r)])
                tw = np.exp(sign * 2j * np.pi * k * np.arange(r) / (r * p))
                vin = vin * tw
                for q in range(r):
                    o = np.sum(vin * np.exp(sign * 2j * np.pi * q * np.arange(r) / r))
                    smem[(i - k) * r + k + q * p] = o
        regs = np.array([[smem[t + e * T] for e in range(R)] for t in range(T)])
        p *= r
    y = smem
    ref = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * n
    return np.max(np.abs(y - ref)) / np.max(np.abs(ref))


def conflicts(n, R, radices, pad_shift, tpc_threads=256):
    """Worst 64-bit bank conflict degree (per half-warp) for the scatter writes and the reads."""
    T = n // R
    per_cta = max(1, tpc_threads // T)
    stride = n + (n >> pad_shift) if pad_shift else n
    nthreads = T * per_cta
    worst_w, worst_r = 1, 1
    p = 1
    for r in radices[:-1] if len(radices) > 1 else []:
        nb = R // r
        for u in range(nb):
            for q in range(r):
                for w0 in range(0, nthreads, 16):
                    banks = {}
                    for tid in range(w0, min(w0 + 16, nthreads)):
                        b, t = divmod(tid, T)
                        i = t + u * T
                        k = i % p
                        pos = (i - k) * r + k + q * p
                        a = b * stride + pos + ((pos >> pad_shift) if pad_shift else 0)
                        banks.setdefault(a % 16, set()).add(a)
                    worst_w = max(worst_w, max(len(v) for v in banks.values()))
        for e in range(R):
            for w0 in range(0, nthreads, 16):
                banks = {}
                for tid in range(w0, min(w0 + 16, nthreads)):
                    b, t = divmod(tid, T)
                    pos = t + e * T
                    a = b * stride + pos + ((pos >> pad_shift) if pad_shift else 0)
                    banks.setdefault(a % 16, set()).add(a)
                worst_r = max(worst_r, max(len(v) for v in banks.values()))
        p *= r
    return worst_w, worst_r


if __name__ == "__main__":
    for lg in range(1, 14):
        n = 1 << lg
        R, radices = config(n)
        err = simulate(n, R, radices) if n <= 1024 else float("nan")
        line = f"N={n:5d} R={R:2d} radices={radices} err={err:.1e}"
        for s in (0, 3, 4, 5):
            line += f"  pad>>{s}: w{conflicts(n, R, radices, s)[0]} r{conflicts(n, R, radices, s)[1]}"
        print(line)
