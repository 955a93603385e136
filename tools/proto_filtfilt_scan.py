"""Prototype (NumPy, CPU) of the tiled two-sweep scan used by the CUDA filtfilt
kernels (spikesort_b200/csrc/filtfilt.cu).  Not product code, not the oracle:
it exists to validate the algebra (constant left padding, tile summaries,
the M matrix, in-warp scan with seeded lane 0) against scipy.signal.filtfilt.

    python tools/proto_filtfilt_scan.py
"""
import numpy as np
from scipy import signal

LANES = 32


def tables(b, a, S):
    n = len(b) - 1
    A = np.zeros((n, n))
    A[:, 0] = -a[1:]
    for k in range(n - 1):
        A[k, k + 1] = 1.0
    T = LANES * S
    mp = np.linalg.matrix_power
    hrow = np.array([mp(A, i)[0] for i in range(S)])          # (S, n)
    P = [mp(A, S * (1 << k)) for k in range(5)]               # scan matrices
    AT = mp(A, T)
    hlast = mp(A, T - 1)[0]
    M = np.zeros((n, n))
    for k in range(n):
        e = np.zeros(n); e[k] = 1.0
        yh = np.array([(mp(A, i) @ e)[0] for i in range(T)])
        w = np.zeros(n)
        step(b, a, w, yh[::-1].copy())
        M[:, k] = w
    y_inf = b.sum() / a.sum()
    zi = np.flip(np.cumsum(np.flip(b - y_inf * a)))[1:]
    return dict(A=A, hrow=hrow, P=P, AT=AT, hlast=hlast, M=M, zi=zi, n=n, S=S, T=T)


def step(b, a, z, v):
    """DF2T in place over v, state z updated."""
    n = len(z)
    for t in range(len(v)):
        x = v[t]
        y = z[0] + b[0] * x
        for k in range(n - 1):
            z[k] = z[k + 1] + b[k + 1] * x - a[k + 1] * y
        z[n - 1] = b[n] * x - a[n] * y
        v[t] = y


def warp_tile_pass(b, a, tb, v, seed, reverse):
    """One warp-tile pass the way the kernel does it: lanes own S consecutive
    samples, lane 0 (or 31 when reversed) starts from `seed`, others from zero;
    inclusive warp scan with P_k; exclusive prefix corrects the local outputs.
    v: (LANES, S) array in TIME order, modified in place.  Returns end state."""
    n, S = tb['n'], tb['S']
    e = np.zeros((LANES, n))
    first = LANES - 1 if reverse else 0
    for l in range(LANES):
        z = seed.copy() if l == first else np.zeros(n)
        seg = v[l, ::-1].copy() if reverse else v[l].copy()
        step(b, a, z, seg)
        v[l] = seg[::-1] if reverse else seg
        e[l] = z
    # inclusive scan over lanes in processing order
    order = list(range(LANES - 1, -1, -1)) if reverse else list(range(LANES))
    inc = e[order].copy()
    for k in range(5):
        d = 1 << k
        new = inc.copy()
        for p in range(d, LANES):
            new[p] = inc[p] + tb['P'][k] @ inc[p - d]
        inc = new
    excl = np.zeros_like(inc)
    excl[1:] = inc[:-1]
    for p, l in enumerate(order):
        if p == 0:
            continue                      # seeded lane already exact
        for i in range(S):
            idx = S - 1 - i if reverse else i
            v[l, idx] += tb['hrow'][i] @ excl[p]
    return inc[-1]


def filtfilt_tiled(b, a, x, S=4):
    b = np.asarray(b, float); a = np.asarray(a, float)
    ntaps = len(b)
    tb = tables(b, a, S)
    n, T = tb['n'], tb['T']
    edge = 3 * ntaps
    L = len(x)
    ext = np.concatenate((2 * x[0] - x[edge:0:-1], x, 2 * x[-1] - x[-2:-(edge + 2):-1])).astype(float)
    Le = len(ext)
    nt = -(-Le // T)
    pad = nt * T - Le
    extp = np.concatenate((np.full(pad, ext[0]), ext))
    zero = np.zeros(n)
    F = np.zeros((nt, n)); Bz = np.zeros((nt, n)); ylast = np.zeros(nt)
    # kernel A
    for j in range(nt):
        v = extp[j * T:(j + 1) * T].reshape(LANES, S).copy()
        F[j] = warp_tile_pass(b, a, tb, v, zero, False)
        ylast[j] = v[-1, -1]
        Bz[j] = warp_tile_pass(b, a, tb, v, zero, True)
    # kernel B
    Zf = np.zeros((nt, n)); Zb = np.zeros((nt, n))
    z = tb['zi'] * ext[0]
    for j in range(nt):
        Zf[j] = z
        z = tb['AT'] @ z + F[j]
    y_last = ylast[nt - 1] + tb['hlast'] @ Zf[nt - 1]
    w = tb['zi'] * y_last
    for j in range(nt - 1, -1, -1):
        Zb[j] = w
        w = tb['AT'] @ w + Bz[j] + tb['M'] @ Zf[j]
    # kernel C
    out = np.zeros(nt * T)
    for j in range(nt):
        v = extp[j * T:(j + 1) * T].reshape(LANES, S).copy()
        warp_tile_pass(b, a, tb, v, Zf[j], False)
        warp_tile_pass(b, a, tb, v, Zb[j], True)
        out[j * T:(j + 1) * T] = v.reshape(-1)
    return out[pad + edge: pad + edge + L]


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    designs = [signal.iirdesign(800 * 2 / 25000, 100 * 2 / 25000, 1, 10, ftype='butter'),
               signal.iirdesign(np.array([300, 6000]) * 2 / 30000, np.array([100, 8000]) * 2 / 30000, 1, 10, ftype='butter')]
    for b, a in designs:
        x = rng.standard_normal(3000) * 100 + 50
        ref = signal.filtfilt(b, a, x)
        got = filtfilt_tiled(b, a, x)
        print(len(b) - 1, "max abs err", np.abs(ref - got).max(), "scale", np.abs(ref).max())
    for b, a in [signal.butter(1, [0.2, 0.5], 'bandpass'), signal.butter(2, [0.2, 0.5], 'bandpass'),
                 signal.butter(4, 0.3), signal.butter(2, [300 / 15000, 6000 / 15000], 'bandpass'),
                 signal.butter(3, [300 / 15000, 6000 / 15000], 'bandpass')]:
        x = rng.standard_normal(3000) * 100 + 50
        ref = signal.filtfilt(b, a, x)
        got = filtfilt_tiled(b, a, x)
        print(len(b) - 1, "max abs err", np.abs(ref - got).max(), "scale", np.abs(ref).max(), 'maxpole', np.abs(np.roots(a)).max())
