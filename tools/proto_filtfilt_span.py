"""Prototype (NumPy, CPU) of the single-pass span kernel of K1 (filt_span_kernel,
spikesort_b200/csrc/filtfilt.cu): fixed left padding with pad + edge = 0 (mod 16), right
padding to whole tiles, spans of W owned tiles, tile-local passes from zero seeds, start state of a tile = end state of
the tile before it (one tile of history; further back is below the truncation bound), the
span's edge states from dot products with the neighbour tiles' samples (F / Bz weights + M), correction by C A^j, the y[-1] override of the right padding.  Also the host-side
truncation bound that decides whether a filter may take this kernel (build_tables: span_err).
Not product code and not the oracle; checked against scipy.signal.filtfilt in
tests/test_proto_filtfilt.py."""
import numpy as np
from scipy import signal

T = 512


def _step(secs, z, x):
    """one sample through the cascade; z: flat state list; returns y"""
    p = 0
    for (b0, b1, b2, a1, a2, first_order) in secs:
        y = z[p] + b0 * x
        if first_order:
            z[p] = b1 * x - a1 * y
            p += 1
        else:
            z[p] = z[p + 1] + b1 * x - a1 * y
            z[p + 1] = b2 * x - a2 * y
            p += 2
        x = y
    return x


def _sections(sos):
    secs = []
    for k, r in enumerate(sos):
        fo = (r[2] == 0.0 and r[5] == 0.0 and k == len(sos) - 1)
        secs.append((r[0], r[1], r[2], r[4], r[5], fo))
    ns = sum(1 if s[5] else 2 for s in secs)
    return secs, ns


def _run(secs, z0, x):
    z = list(z0)
    y = np.empty(len(x))
    for i, xi in enumerate(x):
        y[i] = _step(secs, z, xi)
    return y, np.array(z)


def _state_space(secs, ns):
    A = np.zeros((ns, ns))
    C = np.zeros(ns)
    for k in range(ns):
        z = [1.0 if i == k else 0.0 for i in range(ns)]
        C[k] = _step(secs, z, 0.0)
        A[:, k] = z
    z = [0.0] * ns
    D = _step(secs, z, 1.0)
    B = np.array(z)
    return A, B, C, D


def span_error_bound(sos, K):
    """What build_tables computes as span_err[K-1]."""
    secs, ns = _sections(sos)
    A, B, C, D = _state_space(secs, ns)
    u = np.empty((T, ns))
    cur = B.copy()
    for m in range(T):
        u[m] = cur
        cur = A @ cur
    h = np.empty(T)
    h[0] = D
    h[1:] = u[:-1] @ C
    zmax = np.abs(u).sum(0)
    gain = max(np.abs(h).sum(), 1.0)
    AT = np.linalg.matrix_power(A, T)
    rows = np.empty((T, ns))
    r = C.copy()
    for i in range(T):
        rows[i] = r
        r = r @ A
    R = rows @ np.linalg.matrix_power(AT, K)
    return float((np.abs(R) * zmax).sum(1).max() * gain)


def span_bound(s, t_lo, t_end, nt, W):
    """Owned-tile boundaries of the kernel (span_bound in filtfilt.cu)."""
    b = min(t_lo + W * s, t_end)
    if b == nt - 1 and b > t_lo and b < t_end:
        b = nt - 2
    return b


def filtfilt_span(sos, padlen, x, W=14, K=1):
    """K is kept for the bound only: the kernel uses one tile of history."""
    secs, ns = _sections(sos)
    A, B, C, D = _state_space(secs, ns)
    zi = np.linalg.solve(np.eye(ns) - A, B)
    AT = np.linalg.matrix_power(A, T)
    rows = np.empty((T, ns))                                   # C A^j
    r = C.copy()
    for j in range(T):
        rows[j] = r
        r = r @ A
    # summary weights of one tile (build_tables): F = sum_i u_{T-1-i} x_i, Bz = sum_i x_i G_i
    u = np.empty((T, ns))
    cur = B.copy()
    for m in range(T):
        u[m] = cur
        cur = A @ cur
    h = np.empty(T)
    h[0] = D
    h[1:] = u[:-1] @ C
    Fw = u[::-1]                                               # (T, ns): weight of x_i
    Gw = np.empty((T, ns))
    for i in range(T):
        Gw[i] = (u[i:] * h[:T - i, None]).sum(0)
    M = np.empty((ns, ns))
    for k in range(ns):
        _, M[:, k] = _run(secs, np.zeros(ns), rows[::-1, k])

    x = np.asarray(x, np.float64)
    L = len(x)
    ext = np.concatenate((2 * x[0] - x[padlen:0:-1], x, 2 * x[-1] - x[-2:-padlen - 2:-1]))
    ext_len = L + 2 * padlen
    p = (16 - padlen % 16) % 16
    nt = -(-(p + ext_len) // T)
    q = nt * T - p - ext_len
    sig = np.concatenate((np.full(p, ext[0]), ext, np.zeros(q)))
    out = np.full(nt * T, np.nan)
    n_spans = -(-nt // W)
    for sp in range(n_spans):
        own_lo, own_hi = span_bound(sp, 0, nt, nt, W), span_bound(sp + 1, 0, nt, nt, W)
        if own_hi <= own_lo:
            continue
        tiles = list(range(own_lo, own_hi))
        v, ends = {}, {}
        for t in tiles:                                        # forward, tile-local
            xt = sig[t * T:(t + 1) * T]
            seed = np.zeros(ns)
            if t == 0:
                seed = zi * xt[0]
            elif t == own_lo:                                  # from the tile in front of the span
                xp = sig[(t - 1) * T:t * T]
                seed = xp @ Fw + AT @ (zi * xp[0])
            v[t], ends[t] = _run(secs, seed, xt)
        for t in tiles[1:]:
            v[t] = v[t] + rows @ ends[t - 1]
        endb = {}
        for t in tiles:                                        # backward, tile-local
            y = v[t]
            seed = np.zeros(ns)
            if t == nt - 1:
                o = ext_len - 1 - (t * T - p)
                yl = y[o]
                y = y.copy()
                y[o + 1:] = yl
                seed = zi * yl
            elif t == own_hi - 1:                              # from the tile behind the span
                xn = sig[(t + 1) * T:(t + 2) * T]
                seed = xn @ Gw + M @ ends[t]
            yb, endb[t] = _run(secs, seed, y[::-1])
            v[t] = yb[::-1]
        for t in tiles:
            res = v[t]
            if t < own_hi - 1:
                res = res + (rows @ endb[t + 1])[::-1]
            out[t * T:(t + 1) * T] = res
    return out[p + padlen:p + padlen + L]
