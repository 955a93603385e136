"""Prototype (NumPy, serial) of the top-k symmetric eigensolver planned for pca_eig_top_kernel:
Householder tridiagonalisation -> Sturm multisection -> inverse iteration (pivoted tridiagonal
LU) -> back-transformation.  Same operation order as the CUDA kernel; checked against
numpy.linalg.eigh on covariance-like matrices.  Development aid, not part of the product."""
import numpy as np

EPS = np.finfo(np.float64).eps


def tridiag(A):
    A = A.copy()
    n = A.shape[0]
    d = np.zeros(n); e = np.zeros(max(n - 1, 0)); tau = np.zeros(max(n - 2, 0))
    V = np.zeros((max(n - 2, 0), n))
    for k in range(n - 2):
        alpha = A[k + 1, k]
        xnorm2 = np.sum(A[k + 2:, k] ** 2)
        if xnorm2 == 0.0:
            tau[k] = 0.0
            e[k] = alpha
            d[k] = A[k, k]
            continue
        beta = -np.copysign(np.sqrt(alpha * alpha + xnorm2), alpha)
        tau[k] = (beta - alpha) / beta
        v = np.zeros(n)
        v[k + 1] = 1.0
        v[k + 2:] = A[k + 2:, k] / (alpha - beta)
        V[k] = v
        e[k] = beta
        d[k] = A[k, k]
        p = tau[k] * (A @ v)            # rows/cols <= k of A contribute nothing useful: mask
        p[:k + 1] = 0.0
        w = p - (0.5 * tau[k] * (p @ v)) * v
        A -= np.outer(v, w) + np.outer(w, v)
    if n >= 2:
        d[n - 2] = A[n - 2, n - 2]
        e[n - 2] = A[n - 1, n - 2]
    d[n - 1] = A[n - 1, n - 1]
    return d, e, V, tau


def sturm_count(d, e2, x, pivmin=None):
    """number of eigenvalues < x: sign changes of the leading-minor determinants
    p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2} (no division; T is scaled to unit size)"""
    cnt = 0
    p2, p1 = 0.0, 1.0
    for i in range(len(d)):
        t = (e2[i - 1] if i > 0 else 0.0) * p2
        p = (d[i] - x) * p1 - t
        if p == 0.0:
            p = -np.copysign(1e-300, p1)
        if (p < 0) != (p1 < 0):
            cnt += 1
        if abs(p) < 1e-120 and abs(p1) < 1e-120:
            p *= 1e120; p1 *= 1e120
        elif abs(p) > 1e120:
            p *= 1e-120; p1 *= 1e-120
        p2, p1 = p1, p
    return cnt


def top_eigenvalues(d, e, k, L=16):
    n = len(d)
    ea = np.abs(np.concatenate([[0.0], e, [0.0]]))
    lo0 = np.min(d - ea[:-1] - ea[1:]); hi0 = np.max(d + ea[:-1] + ea[1:])
    tn = max(abs(lo0), abs(hi0))
    lo0 -= 2 * EPS * tn * n + 2e-300; hi0 += 2 * EPS * tn * n + 2e-300
    e2 = e * e
    pivmin = max(np.max(e2) if n > 1 else 0.0, 1.0) * 2.2e-308 if False else max((np.max(e2) if n > 1 else 0.0) * 2.2e-308, 2.2e-308)
    passes = int(np.ceil(56 * np.log(2) / np.log(L + 1))) + 1
    lam = np.zeros(k)
    for m in range(k):
        j = n - 1 - m                      # ascending index of the m-th largest
        lo, hi = lo0, hi0                  # count(lo) <= j < count(hi)
        for _ in range(passes):
            xs = lo + (hi - lo) * (np.arange(L) + 1.0) / (L + 1.0)
            cs = np.array([sturm_count(d, e2, x, pivmin) for x in xs])
            above = np.nonzero(cs > j)[0]
            f = above[0] if len(above) else L
            nlo = xs[f - 1] if f > 0 else lo
            nhi = xs[f] if f < L else hi
            lo, hi = nlo, nhi
        lam[m] = 0.5 * (lo + hi)
    return lam


def lu_tridiag(d, e, lam):
    n = len(d)
    a = d - lam
    b = e.copy(); c = e.copy(); d2 = np.zeros(max(n - 2, 0)); piv = np.zeros(max(n - 1, 0), dtype=int)
    for k in range(n - 1):
        if abs(a[k]) >= abs(c[k]):
            piv[k] = 0
            if a[k] != 0.0:
                mult = c[k] / a[k]
                c[k] = mult
                a[k + 1] -= mult * b[k]
            else:
                c[k] = 0.0
            if k < n - 2: d2[k] = 0.0
        else:
            piv[k] = 1
            mult = a[k] / c[k]
            a[k] = c[k]
            temp = a[k + 1]
            a[k + 1] = b[k] - mult * temp
            if k < n - 2:
                d2[k] = b[k + 1]
                b[k + 1] = -mult * d2[k]
            b[k] = temp
            c[k] = mult
    return a, b, c, d2, piv


def lu_solve(a, b, c, d2, piv, y, tol):
    n = len(a)
    y = y.copy()
    for k in range(n - 1):
        if piv[k] == 0:
            y[k + 1] -= c[k] * y[k]
        else:
            t = y[k]; y[k] = y[k + 1]; y[k + 1] = t - c[k] * y[k]
    for k in range(n - 1, -1, -1):
        t = y[k]
        if k < n - 1: t -= b[k] * y[k + 1]
        if k < n - 2: t -= d2[k] * y[k + 2]
        ak = a[k]
        if abs(ak) < tol: ak = np.copysign(tol, ak) if ak != 0 else tol
        y[k] = t / ak
    return y


def eig_top(A, k, iters=3):
    n = A.shape[0]
    d, e, V, tau = tridiag(A)
    scale = max(np.max(np.abs(d)), np.max(np.abs(e)) if n > 1 else 0.0)
    if not scale > 0.0: scale = 1.0
    d = d / scale; e = e / scale
    lam = top_eigenvalues(d, e, k)
    tn = 1.0
    tol = EPS * tn
    Z = np.zeros((k, n))
    def start(m):
        h = ((np.arange(n, dtype=np.uint64) + 1) * 2654435761 & 0xffffffff) ^ (((m + 1) * 2246822519) & 0xffffffff)
        h ^= h >> 15; h = (h * 2246822519) & 0xffffffff; h ^= h >> 13
        return (h >> 8).astype(np.float64) / 16777216.0 + 0.25
    facts = []
    for m in range(k):
        # equal computed eigenvalues: separate them by a few ulps so the LUs differ (dstein)
        lm = lam[m]
        if m > 0 and abs(lam[m] - lam[m - 1]) <= 10 * EPS * tn:
            lm = lam_used[m - 1] - 10 * EPS * tn
        if m == 0: lam_used = np.zeros(k)
        lam_used[m] = lm
        facts.append(lu_tridiag(d, e, lm))
        Z[m] = start(m)
    for it in range(iters):
        for m in range(k):
            z = lu_solve(*facts[m], Z[m], tol)
            Z[m] = z / np.max(np.abs(z))
        for m in range(k):                 # Gram-Schmidt inside clusters, then normalise
            for j in range(m):
                if abs(lam[m] - lam[j]) < 1e-3 * tn:
                    Z[m] -= (Z[m] @ Z[j]) * Z[j]
            Z[m] /= np.sqrt(Z[m] @ Z[m])
    for kk in range(n - 3, -1, -1):        # back-transformation
        if tau[kk] == 0.0: continue
        for m in range(k):
            s = V[kk] @ Z[m]
            Z[m] -= tau[kk] * s * V[kk]
    return lam * scale, Z.T


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    worst = 0.0
    cases = []
    for n, nspk in [(30, 5000), (25, 3000), (42, 800), (112, 5000), (3, 50), (2, 10), (30, 40)]:
        t = np.linspace(0, 1, n)
        templ = np.stack([np.exp(-((t - 0.3) / 0.08) ** 2) - 0.4 * np.exp(-((t - 0.5) / 0.15) ** 2),
                          np.exp(-((t - 0.35) / 0.05) ** 2), np.sin(6 * t) * np.exp(-3 * t)])
        amp = rng.standard_normal((nspk, 3)) * np.array([120.0, 60.0, 20.0])
        X = (amp @ templ + rng.standard_normal((nspk, n)) * 8.0).astype(np.float32).astype(np.float64)
        cases.append(("spikes n=%d" % n, np.cov(X.T)))
    Q, _ = np.linalg.qr(rng.standard_normal((30, 30)))
    cases.append(("degenerate top pair", Q @ np.diag([5.0, 5.0] + list(np.linspace(1, 0.01, 28))) @ Q.T))
    cases.append(("near-degenerate", Q @ np.diag([5.0, 5.0 - 1e-9] + list(np.linspace(1, 0.01, 28))) @ Q.T))
    cases.append(("rank 1", np.outer(Q[:, 0], Q[:, 0]) * 7.0))
    cases.append(("diagonal", np.diag(np.arange(30, 0, -1.0))))
    cases.append(("zeros", np.zeros((8, 8))))
    for name, A in cases:
        A = 0.5 * (A + A.T)
        k = min(2, A.shape[0])
        lam, Z = eig_top(A, k)
        w, U = np.linalg.eigh(A)
        w = w[::-1]; U = U[:, ::-1]
        lerr = np.max(np.abs(lam - w[:k])) / max(np.max(np.abs(w)), 1e-300)
        res = max(np.linalg.norm(A @ Z[:, m] - lam[m] * Z[:, m]) for m in range(k)) / max(np.max(np.abs(w)), 1e-300)
        orth = np.max(np.abs(Z.T @ Z - np.eye(k)))
        # subspace / vector agreement (sign-normalised) where the gap allows it
        verr = max(min(np.linalg.norm(Z[:, m] - U[:, m]), np.linalg.norm(Z[:, m] + U[:, m])) for m in range(k))
        print("%-22s n=%3d  eval err %.2e  residual %.2e  orth %.2e  vec err %.2e" % (name, A.shape[0], lerr, res, orth, verr))
