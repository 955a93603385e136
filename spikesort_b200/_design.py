"""(b, a) transfer function -> cascade of second-order sections for the CUDA scan.

The reference designs its filters with ``scipy.signal.iirdesign`` and applies them
in (b, a) form (``src/spike_sort/core/filters.py:86-101``).  A parallel scan in the
companion-matrix basis of (b, a) loses many digits for band-pass designs with poles
clustered near z = 1 (powers of the companion matrix are violently non-normal), so
the CUDA library runs the SAME filter as a cascade of biquads.  To stay the
reference's filter - the one defined by the ROUNDED float64 coefficients b, a -
the sections are derived from b and a themselves:

1. roots of both polynomials by Aberth-Ehrlich iteration in multi-precision
   arithmetic (mpmath, 250 digits), started from ``np.roots``; clustered roots
   converge fast, exactly repeated ones (e.g. Butterworth band-pass zeros at +-1)
   linearly, which the working precision leaves room for;
2. pole/zero pairing and section ordering by ``scipy.signal.zpk2sos``;
3. the overall gain b[0]/a[0] spread by zpk2sos (first section).

The cascade is then algebraically identical to (b, a) up to one float64 rounding
per section coefficient, which is below the reference's own rounding noise.
"""
import numpy as np

try:
    import mpmath as mp
except ImportError:                                   # pragma: no cover
    mp = None

_DPS = 250          # a 6-fold root is resolved to ~10^(-250/6); float64 needs 1e-17
_TOL_EXP = 28


def _strip(c):
    c = np.atleast_1d(np.asarray(c, dtype=np.float64))
    nz = np.flatnonzero(c)
    if len(nz) == 0:
        raise ValueError("polynomial is identically zero")
    return c[nz[0]:]


def _aberth(coeffs, guess):
    """Roots of sum coeffs[k] z^(n-k), refined from `guess` in multi-precision."""
    n = len(coeffs) - 1
    if n == 0:
        return []
    with mp.workdps(_DPS):
        c = [mp.mpf(float(v)) for v in coeffs]
        dc = [c[k] * (n - k) for k in range(n)]
        z = [mp.mpc(complex(g)) for g in guess]
        # separate coincident starting points
        for i in range(n):
            for j in range(i):
                if z[i] == z[j]:
                    z[i] += mp.mpc(1e-8 * (i + 1), 1e-8 * (j + 1))
        tol = mp.mpf(10) ** (-_TOL_EXP)
        for _ in range(3000):
            worst = mp.mpf(0)
            for i in range(n):
                p = mp.polyval(c, z[i])
                if p == 0:
                    continue
                d = mp.polyval(dc, z[i])
                s = mp.mpc(0)
                for j in range(n):
                    if j != i:
                        s += 1 / (z[i] - z[j])
                w = p / d
                step = w / (1 - w * s)
                z[i] -= step
                m = abs(step) / max(abs(z[i]), mp.mpf(1))
                if m > worst:
                    worst = m
            if worst < tol:
                break
        else:
            raise RuntimeError("filter design: root refinement did not converge")
        return [complex(v) for v in z]


def _clean_conjugates(r):
    """Make the root set exactly closed under conjugation (it is, to ~1e-45)."""
    r = np.asarray(r, dtype=np.complex128)
    scale = max(1.0, np.abs(r).max() if len(r) else 1.0)
    real = np.abs(r.imag) <= 1e-14 * scale
    out = list(r[real].real.astype(np.complex128))
    pos = sorted(r[~real & (r.imag > 0)], key=lambda v: (v.real, v.imag))
    neg = sorted(r[~real & (r.imag < 0)], key=lambda v: (v.real, -v.imag))
    if len(pos) != len(neg):
        raise RuntimeError("filter design: unpaired complex roots")
    for u, v in zip(pos, neg):
        m = 0.5 * (u + np.conj(v))
        out += [m, np.conj(m)]
    return np.array(out, dtype=np.complex128)


_FACTORED = {}      # (b bytes, a bytes) -> (sos, padlen): the factorisation costs milliseconds


def tf2sos_exact(b, a):
    """Return (sos, padlen): float64 array (n_sections, 6) in SciPy's layout with
    a0 == 1, a possible first-order section last; padlen = 3*max(len(a), len(b))
    as scipy.signal.filtfilt computes it for the ORIGINAL coefficient vectors.
    Results are memoised per coefficient vectors (fltLinearIIR builds a fresh Filter
    object on every call)."""
    key = (np.asarray(b, dtype=np.float64).tobytes(), np.asarray(a, dtype=np.float64).tobytes())
    hit = _FACTORED.get(key)
    if hit is None:
        if len(_FACTORED) > 256:
            _FACTORED.clear()
        hit = _FACTORED[key] = _tf2sos_exact(b, a)
    return hit[0].copy(), hit[1]


def _tf2sos_exact(b, a):
    from scipy import signal
    if mp is None:
        raise RuntimeError("mpmath is required to factor the filter")
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    if b.ndim != 1 or a.ndim != 1:
        raise ValueError("b and a must be 1-D")
    padlen = 3 * max(len(a), len(b))
    if a[0] == 0:
        raise ValueError("First coefficient of parameter `a` must be non-zero!")
    if len(a) == 1:
        raise NotImplementedError("FIR filters (len(a) == 1) are not on the CUDA path yet")
    # H(z) = (b0 + b1 z^-1 + ...)/(a0 + a1 z^-1 + ...): pad to equal length so both
    # are polynomials in z of the same degree (extra roots at z = 0)
    n = max(len(a), len(b))
    bz = np.concatenate((b, np.zeros(n - len(b))))
    az = np.concatenate((a, np.zeros(n - len(a))))
    bs, as_ = _strip(bz), _strip(az)
    k = bs[0] / as_[0]
    z = _clean_conjugates(_aberth(bs, np.roots(bs))) if len(bs) > 1 else np.zeros(0, complex)
    p = _clean_conjugates(_aberth(as_, np.roots(as_))) if len(as_) > 1 else np.zeros(0, complex)
    if len(p) and np.abs(p).max() >= 1.0:
        raise ValueError("unstable filter (pole magnitude %.6g >= 1): the reference's "
                         "filtfilt diverges too" % np.abs(p).max())
    # strip leading zeros of b <=> pure delay z^-(d); filtfilt is then NOT zero phase
    # but still well defined: keep it as zeros at z = 0 would need len(z) < len(p),
    # which zpk2sos handles (it pads with zeros at the origin).
    sos = signal.zpk2sos(z, p, k, pairing='nearest')
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    # the library wants the (at most one) first-order section last; a cascade of
    # LTI sections commutes, so reordering does not change the filter
    is_first = (sos[:, 2] == 0.0) & (sos[:, 5] == 0.0)
    fo = np.flatnonzero(is_first)
    if len(fo) > 0:
        keep_last = fo[-1]
        order = [i for i in range(len(sos)) if i != keep_last] + [keep_last]
        sos = sos[order]
    out = np.ascontiguousarray(sos, dtype=np.float64)
    if not np.all(out[:, 3] == 1.0):
        raise RuntimeError("filter design: section with a0 != 1")
    return out, padlen
