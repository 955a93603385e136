"""Prototype (NumPy, CPU) of the algebra of K1 as built (spikesort_b200/csrc/filtfilt.cu):
second-order-section cascade, constant left padding to whole tiles, sweep 1 (tile summaries
F, Bz, ylast), the tile scan (Zf, y[-1], Zb with the M term), and sweep 2 SECTION BY SECTION
(lane-local pass, 5-step lane scan with the section's own 2x2 powers, correction by C_s A_s^i).
Not product code and not the oracle: it states the operation structure the kernels replay and
is checked against the sequential filter and scipy.signal.filtfilt (tests/test_proto_filtfilt.py).
"""
import numpy as np
from scipy import signal

LANES = 32


def _sec_step(c, z, x):
    """one DF2T section over the vector x (in place semantics of the kernel); returns y, z"""
    b0, b1, b2, a1, a2 = c
    y = np.empty(len(x))
    z0, z1 = z
    for i, xi in enumerate(x):
        yi = z0 + b0 * xi
        z0 = z1 + b1 * xi - a1 * yi
        z1 = b2 * xi - a2 * yi
        y[i] = yi
    return y, np.array([z0, z1])


def _cascade(secs, Z, x):
    """whole cascade over x from the state Z (nsec, 2); returns y, end state"""
    Z = Z.copy()
    for s, c in enumerate(secs):
        x, Z[s] = _sec_step(c, Z[s], x)
    return x, Z


def _odd_ext(x, n):
    return np.concatenate((2 * x[0] - x[n:0:-1], x, 2 * x[-1] - x[-2:-n - 2:-1]))


def filtfilt_sequential(sos, padlen, x):
    secs = [(r[0], r[1], r[2], r[4], r[5]) for r in sos]
    zi = signal.sosfilt_zi(sos)
    ext = _odd_ext(np.asarray(x, np.float64), padlen)
    y, _ = _cascade(secs, zi * ext[0], ext)
    y2, _ = _cascade(secs, zi * y[-1], y[::-1])
    return y2[::-1][padlen:len(ext) - padlen]


def filtfilt_tiled(sos, padlen, x, S=16):
    """The tiled two-sweep formulation; S samples per lane, T = 32 S per tile."""
    secs = [(r[0], r[1], r[2], r[4], r[5]) for r in sos]
    nsec = len(secs)
    T = LANES * S
    zi = signal.sosfilt_zi(sos)
    ext = _odd_ext(np.asarray(x, np.float64), padlen)
    nt = -(-len(ext) // T)
    pad = nt * T - len(ext)
    sig = np.concatenate((np.full(pad, ext[0]), ext))          # constant left padding
    tiles = sig.reshape(nt, T)
    zero = np.zeros((nsec, 2))

    def hom_state(Z, n):                                        # A^n Z (cascade, zero input)
        return _cascade(secs, Z, np.zeros(n))[1]

    # ---- sweep 1: zero-state summaries per tile
    F = np.empty((nt, nsec, 2)); Bz = np.empty((nt, nsec, 2)); ylast = np.empty(nt)
    for t in range(nt):
        y0, F[t] = _cascade(secs, zero, tiles[t])
        ylast[t] = y0[-1]
        _, Bz[t] = _cascade(secs, zero, y0[::-1])
    # ---- tile scan
    Zf = np.empty((nt, nsec, 2))
    Zf[0] = zi * sig[0]
    for t in range(1, nt):
        Zf[t] = hom_state(Zf[t - 1], T) + F[t - 1]
    y_last = ylast[-1] + _cascade(secs, Zf[-1], np.zeros(T))[0][-1]
    Zb = np.empty((nt, nsec, 2))
    Zb[-1] = zi * y_last
    for t in range(nt - 1, 0, -1):
        yh = _cascade(secs, Zf[t], np.zeros(T))[0]              # forward homogeneous output of tile t
        MZf = _cascade(secs, zero, yh[::-1])[1]
        Zb[t - 1] = hom_state(Zb[t], T) + Bz[t] + MZf
    # ---- per-section tables: C_s A_s^i (i < S) and A_s^(S 2^k)
    hrow = np.empty((nsec, S, 2)); P = np.empty((nsec, 5, 2, 2))
    for s, c in enumerate(secs):
        for j in range(2):
            e = np.zeros(2); e[j] = 1.0
            yh, _ = _sec_step(c, e, np.zeros(S))
            hrow[s, :, j] = yh
            for k in range(5):
                P[s, k, :, j] = _sec_step(c, e, np.zeros(S * (1 << k)))[1]

    def section_pass(c, h, Pm, v, ztile):
        """v: (32, S) lane-major tile in processing order; ztile: the section's tile state"""
        e = np.zeros((LANES, 2))
        for l in range(LANES):
            v[l], e[l] = _sec_step(c, ztile if l == 0 else np.zeros(2), v[l])
        for k in range(5):                                      # inclusive lane scan
            d = 1 << k
            o = np.vstack((np.zeros((d, 2)), e[:-d]))
            e = e + o @ Pm[k].T * (np.arange(LANES) >= d)[:, None]
        pre = np.vstack((np.zeros((1, 2)), e[:-1]))             # state at the lane's start
        v += pre @ h.T                                          # (32, 2) @ (2, S)
        return v

    out = np.empty_like(tiles)
    for t in range(nt):
        v = tiles[t].reshape(LANES, S).copy()
        for s, c in enumerate(secs):
            v = section_pass(c, hrow[s], P[s], v, Zf[t, s])
        v = v[::-1, ::-1].copy()                                # backward: lane 31 first, reversed samples
        for s, c in enumerate(secs):
            v = section_pass(c, hrow[s], P[s], v, Zb[t, s])
        out[t] = v[::-1, ::-1].reshape(T)
    return out.reshape(-1)[pad + padlen:pad + len(ext) - padlen]


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    b, a = signal.iirdesign(np.array([300, 6000]) * 2 / 30000, np.array([100, 8000]) * 2 / 30000, 1, 10,
                            ftype='butter')
    sos = signal.tf2sos(b, a)
    x = np.round(rng.standard_normal(3000) * 300)
    ref = signal.filtfilt(b, a, x)
    for f in (filtfilt_sequential, filtfilt_tiled):
        y = f(sos, 3 * len(b), x)
        print(f.__name__, np.abs(y - ref).max() / np.abs(ref).max())
