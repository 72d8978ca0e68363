"""Offline numpy model of the mirrored-pass variants of the in-CTA engine (fft_engine.cuh,
variants 1 and 2) used by the R2C epilogue and the C2R prologue of rows_kernel.cuh.

Variant 1 (R2C): in the LAST pass thread t owns butterflies i and q0 - i (q0 = N/r butterflies per
pass), so that Z[k] and Z[N-k] end up in the same thread and the resolve needs no exchange.
Variant 2 (C2R): the same pairing in the FIRST pass, so X[k] and X[N-k] are loaded by the thread
that needs both.

Checks the index maps against numpy.fft and counts shared-memory bank conflicts (64-bit accesses,
per half-warp) of every scatter / gather for a given shape and pad shift.

    python tools/mirror_model.py            # self-test over the shapes of engine_shape(n, 1|2)
"""
import sys
import numpy as np


def bfly_index(t, u, T, nb, q0, mirrored):
    """Butterfly handled by thread t in register group u of a pass with q0 butterflies."""
    if not mirrored:
        return t + u * T
    h = nb // 2
    if u < h:
        return t + u * T
    i0 = t + (u - h) * T
    return q0 // 2 if i0 == 0 else q0 - i0


def run_passes(regs_in, n, R, radices, sign, variant, load_stage=None):
    """regs_in: None (then load_stage(t, slot->value) fills pass-0 inputs) or natural array x.
    Returns the per-thread output registers of the last pass and the mapping slot -> element."""
    T = n // R
    P = len(radices)
    p = 1
    smem = None
    out_regs = None
    for j, r in enumerate(radices):
        nb = R // r
        q0 = n // r
        mirrored = (variant == 1 and j == P - 1) or (variant == 2 and j == 0)
        nxt = np.zeros(n, complex)
        out_regs = {}
        for t in range(T):
            for u in range(nb):
                i = bfly_index(t, u, T, nb, q0, mirrored)
                src = regs_in if j == 0 else smem
                vin = np.array([src[i + m * q0] for m in range(r)])
                k = i % p
                vin = vin * np.exp(sign * 2j * np.pi * k * np.arange(r) / (r * p))
                for q in range(r):
                    o = np.sum(vin * np.exp(sign * 2j * np.pi * q * np.arange(r) / r))
                    pos = (i - k) * r + k + q * p
                    nxt[pos] = o
                    out_regs[(t, u + q * nb)] = pos
        smem = nxt
        p *= r
    return smem, out_regs


def check_variant(n, R, radices, variant):
    rng = np.random.default_rng(n + variant)
    T = n // R
    P = len(radices)
    if variant == 1:
        # R2C of 2n reals: Z = FFT_n(x[2j] + i x[2j+1]); every pair (k, n-k) must sit in one thread
        x = rng.standard_normal(2 * n)
        z = x[0::2] + 1j * x[1::2]
        Z, slots = run_passes(z, n, R, radices, -1, 1)
        assert np.allclose(Z, np.fft.fft(z))
        owner = {}
        for (t, s), pos in slots.items():
            owner[pos] = t
        for k in range(1, n):
            assert owner[k] == owner[n - k], (n, k)
        assert owner[0] == 0
        # resolve
        rtw = -1j * np.exp(-2j * np.pi * np.arange(n) / (2 * n))
        X = np.zeros(n + 1, complex)
        X[0] = Z[0].real + Z[0].imag
        X[n] = Z[0].real - Z[0].imag
        for k in range(1, n):
            fe = Z[k] + np.conj(Z[n - k])
            fo = rtw[k] * (Z[k] - np.conj(Z[n - k]))
            X[k] = 0.5 * (fe + fo)
            assert np.allclose(np.conj(0.5 * (fe - fo)), 0.5 * ((Z[n - k] + np.conj(Z[k])) + rtw[n - k] * (Z[n - k] - np.conj(Z[k]))))
        assert np.allclose(X, np.fft.rfft(x))
    else:
        # C2R: z[k] = (X[k] + conj X[n-k]) + (X[k] - conj X[n-k]) rtw[k]; thread loading X[k] for a
        # first-pass input must also be the one loading X[n-k]
        x = rng.standard_normal(2 * n)
        X = np.fft.rfft(x)
        rtw = 1j * np.exp(2j * np.pi * np.arange(n) / (2 * n))
        z = np.array([(X[k] + np.conj(X[n - k])) + (X[k] - np.conj(X[n - k])) * rtw[k] for k in range(n)])
        r = radices[0]
        nb = R // r
        q0 = n // r
        owner = {}
        for t in range(T):
            for u in range(nb):
                i = bfly_index(t, u, T, nb, q0, True)
                for m in range(r):
                    owner[i + m * q0] = t
        for k in range(1, n):
            assert owner[k] == owner[n - k], (n, k)
        y, _ = run_passes(z, n, R, radices, +1, 2)
        want = np.fft.ifft(z) * n
        assert np.allclose(y, want)
        assert np.allclose(y.real, 2 * n * x[0::2]) and np.allclose(y.imag, 2 * n * x[1::2])  # C2R(R2C(x)) = N x, N = 2n


def conflicts(n, R, radices, pad_shift, variant, cta_threads=256):
    """(worst scatter degree, worst gather degree, mean wavefronts per 64-bit access / ideal)."""
    T = n // R
    per_cta = max(1, cta_threads // T)
    stride = n + (n >> pad_shift) + 1
    nthreads = T * per_cta
    P = len(radices)
    worst_w = worst_r = 1
    wave = ideal = 0

    def degree(addrs):
        banks = {}
        for a in addrs:
            banks.setdefault(a % 16, set()).add(a)
        return max(len(v) for v in banks.values())

    p = 1
    for j, r in enumerate(radices):
        nb = R // r
        q0 = n // r
        mir = (variant == 1 and j == P - 1) or (variant == 2 and j == 0)
        if j > 0:  # gather of this pass's inputs
            for u in range(nb):
                for m in range(r):
                    for w0 in range(0, nthreads, 16):
                        addrs = []
                        for tid in range(w0, min(w0 + 16, nthreads)):
                            b, t = divmod(tid, T)
                            pos = bfly_index(t, u, T, nb, q0, mir) + m * q0
                            addrs.append(b * stride + pos + (pos >> pad_shift))
                        d = degree(addrs)
                        worst_r = max(worst_r, d)
                        wave += d
                        ideal += 1
        if j < P - 1:  # scatter of this pass's outputs
            for u in range(nb):
                for q in range(r):
                    for w0 in range(0, nthreads, 16):
                        addrs = []
                        for tid in range(w0, min(w0 + 16, nthreads)):
                            b, t = divmod(tid, T)
                            i = bfly_index(t, u, T, nb, q0, mir)
                            k = i % p
                            pos = (i - k) * r + k + q * p
                            addrs.append(b * stride + pos + (pos >> pad_shift))
                        d = degree(addrs)
                        worst_w = max(worst_w, d)
                        wave += d
                        ideal += 1
        p *= r
    return worst_w, worst_r, wave / max(ideal, 1)


SHAPES = {
    # n: {variant: (R, radices)}
    512: {1: (16, [16, 16, 2]), 2: (16, [2, 16, 16])},
    1024: {1: (16, [16, 16, 4]), 2: (16, [4, 16, 16])},
    2048: {1: (32, [32, 32, 2]), 2: (32, [2, 32, 32])},
    4096: {1: (32, [16, 16, 16]), 2: (32, [16, 16, 16])},
    8192: {1: (32, [32, 16, 16]), 2: (32, [16, 16, 32])},
}

if __name__ == "__main__":
    extra = {2048: {1: (16, [16, 16, 8]), 2: (16, [8, 16, 16])}}
    for table in (SHAPES, extra):
        for n, vs in table.items():
            for v, (R, radices) in vs.items():
                if n <= 2048:
                    check_variant(n, R, radices, v)
                line = f"N={n:5d} V={v} R={R:2d} radices={radices}"
                for s in (3, 4, 5):
                    w, r_, mean = conflicts(n, R, radices, s, v)
                    line += f"  pad>>{s}: w{w} r{r_} x{mean:.2f}"
                print(line, flush=True)
    # small shapes: index maps only
    for n, R, radices in ((8, 4, [4, 2]), (32, 4, [4, 4, 2]), (128, 8, [8, 8, 2]), (16, 4, [2, 2, 4]), (64, 8, [4, 4, 4])):
        for v in (1, 2):
            rr = radices if v == 1 else radices[::-1]
            nbk = R // (rr[-1] if v == 1 else rr[0])
            if nbk >= 2:
                check_variant(n, R, rr, v)
                print(f"N={n} V={v} R={R} radices={rr} ok")
