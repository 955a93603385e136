"""Helpers shared by the CPU and GPU test tiers."""
import numpy as np


def sine_fixture():
    """Analytic signal of the reference's tests/test_core.py:45-53."""
    n_spikes, FS = 100, 25E3
    period = 1000.0 / FS * 100
    time = np.arange(0, period * n_spikes, 1000.0 / FS)
    spikes = np.sin(2 * np.pi / period * time)[np.newaxis, :]
    return {"data": spikes, "n_contacts": 1, "FS": FS}, period, time


def sign_normalise(scores, ref):
    """Flip every column of `scores` whose dot product with `ref` is negative
    (eigenvector signs are arbitrary in LAPACK and in Jacobi alike)."""
    s = np.sign(np.sum(scores * ref, axis=0))
    s[s == 0] = 1.0
    return scores * s[np.newaxis, :]


def rel_err(got, ref):
    """max |got - ref| relative to the largest reference magnitude."""
    scale = np.abs(ref).max()
    return np.abs(got - ref).max() / (scale if scale > 0 else 1.0)
