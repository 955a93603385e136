"""Host-logic tier: the NumPy statement (tools/proto_eig_top.py) of the operation order that
pca_eig_top_kernel replays - Householder tridiagonalisation, division-free Sturm multisection,
pivoted tridiagonal LU + inverse iteration, back-transformation - against numpy.linalg.eigh.
The GPU tier (test_pca_eig_top_*) checks the kernel itself against the same oracle."""
import importlib.util
import os

import numpy as np
import pytest

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools",
                     "proto_eig_top.py")
_spec = importlib.util.spec_from_file_location("proto_eig_top", _PATH)
proto = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(proto)


def _spike_cov(rng, n, n_spk):
    t = np.linspace(0, 1, n)
    templ = np.stack([np.exp(-((t - 0.3) / 0.08) ** 2) - 0.4 * np.exp(-((t - 0.5) / 0.15) ** 2),
                      np.exp(-((t - 0.35) / 0.05) ** 2), np.sin(6 * t) * np.exp(-3 * t)])
    amp = rng.standard_normal((n_spk, 3)) * np.array([120.0, 60.0, 20.0])
    X = (amp @ templ + rng.standard_normal((n_spk, n)) * 8.0).astype(np.float32).astype(np.float64)
    return np.atleast_2d(np.cov(X.T))


@pytest.mark.parametrize("n,k", [(30, 2), (25, 1), (42, 3), (64, 8), (3, 2), (2, 2), (1, 1)])
def test_leading_pairs_of_spike_covariances(n, k):
    rng = np.random.default_rng(n * 10 + k)
    A = _spike_cov(rng, n, 2000)
    lam, Z = proto.eig_top(A, k)
    w, U = np.linalg.eigh(A)
    w, U = w[::-1], U[:, ::-1]
    assert np.allclose(lam, w[:k], rtol=1e-12, atol=1e-12 * w[0])
    assert np.abs(Z.T @ Z - np.eye(k)).max() < 1e-12
    for j in range(k):
        gap = min([abs(w[j] - w[i]) for i in range(n) if i != j] or [w[0]]) / w[0]
        err = min(np.linalg.norm(Z[:, j] - U[:, j]), np.linalg.norm(Z[:, j] + U[:, j]))
        assert err < 1e-12 / max(gap, 1e-6), (j, gap, err)


def test_degenerate_and_rank_deficient():
    rng = np.random.default_rng(4)
    n = 30
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    mats = [Q @ np.diag([5.0, 5.0] + list(np.linspace(1, 0.01, n - 2))) @ Q.T,
            Q @ np.diag([5.0, 5.0 - 1e-9] + list(np.linspace(1, 0.01, n - 2))) @ Q.T,
            7.0 * np.outer(Q[:, 0], Q[:, 0]), np.diag(np.arange(n, 0, -1.0)), np.zeros((n, n))]
    for A in mats:
        A = 0.5 * (A + A.T)
        for k in (1, 2, 4):
            lam, Z = proto.eig_top(A, k)
            w = np.linalg.eigvalsh(A)[::-1]
            scale = max(w[0], 1e-300)
            assert np.abs(lam - w[:k]).max() <= 1e-12 * scale + 1e-290
            assert np.abs(Z.T @ Z - np.eye(k)).max() < 1e-10
            assert np.abs(A @ Z - Z * lam).max() <= 1e-10 * scale + 1e-290


def test_sturm_count_matches_eigenvalue_positions():
    rng = np.random.default_rng(9)
    d = rng.standard_normal(40)
    e = rng.standard_normal(39) * 0.5
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    w = np.linalg.eigvalsh(T)
    for x in np.concatenate([[w[0] - 1, w[-1] + 1], 0.5 * (w[1:] + w[:-1])]):
        assert proto.sturm_count(d, e * e, x) == int(np.sum(w < x))
