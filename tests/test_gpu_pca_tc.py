"""GPU tier: the tensor-core Gram kernel (tcgen05, 3xTF32, csrc/pca_tc.cu) against the
float64 CUDA-core kernel and the oracle.  Tolerances: covariance within 5e-6 of its largest
entry (3xTF32 drops the lo*lo term; the float32 TMEM accumulation truncates), scores within the north_star bar
of 1e-5 (measured ~1e-7)."""
import os

import numpy as np
import pytest

from util import sign_normalise, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_tc():
    old = os.environ.get("SS_PCA_TC")
    yield lambda v: os.environ.__setitem__("SS_PCA_TC", v)
    if old is None:
        os.environ.pop("SS_PCA_TC", None)
    else:
        os.environ["SS_PCA_TC"] = old


@pytest.mark.parametrize("n_win,n_spk,n_c", [(25, 100000, 4), (30, 40004, 1), (31, 20000, 2),
                                             (8, 3000, 3), (25, 36, 4), (25, 1000000, 4)])
def test_tc_gram_contact_mode(cuda, force_tc, n_win, n_spk, n_c):
    import torch
    from spikesort_b200 import _ops, synth
    w = synth.make_waveforms(n_win, n_spk, n_c, seed=11, device=cuda) + 2.5
    def timed(flag, **kw):
        force_tc(flag)
        out = _ops.pca_gram(w, **kw)                       # warm-up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = _ops.pca_gram(w, **kw)
        e1.record()
        torch.cuda.synchronize()
        return out, e0.elapsed_time(e1)

    (g0, s0, ref, counts), t0 = timed("0")
    (g1, s1, _, _), t1 = timed("1", ref=ref)
    scale = float(g0.abs().max())
    print("gram %dx%dx%d: fp64 cores %.3f ms, tensor cores %.3f ms, max rel diff %.3g" % (
        n_win, n_spk, n_c, t0, t1, float((g1 - g0).abs().max()) / scale))
    # the TMEM accumulation truncates: a systematic ~2e-6 shrink of the Gram (it rescales the
    # eigenvalues, i.e. the whitened scores, by ~1e-6 - far inside the 1e-5 bar)
    assert float((g1 - g0).abs().max()) <= 5e-6 * scale
    assert float((s1 - s0).abs().max()) <= 5e-6 * float(s0.abs().max() + 1.0)
    assert float((g1 - g1.transpose(1, 2)).abs().max()) <= 5e-6 * scale
    ev0, vec0 = _ops.pca_eig(g0, s0, counts)
    ev1, vec1 = _ops.pca_eig(g1, s1, counts)
    sc0 = _ops.pca_project(w, ev0, vec0, 2).cpu().numpy()
    sc1 = _ops.pca_project(w, ev1, vec1, 2).cpu().numpy()
    err = rel_err(sign_normalise(sc1, sc0), sc0)
    print("tc scores rel err %.3g" % err)
    assert err < 1e-5


def test_tc_gram_offsets_mode(cuda, force_tc):
    import torch
    from spikesort_b200 import _ops, synth
    n_win = 30
    counts = [5000, 0, 12345, 7, 40001, 1, 9999]
    n_spk = sum(counts)
    pad = (-n_spk) % 4
    counts[-1] += pad                                     # rows must stay 16-byte aligned
    n_spk += pad
    w = synth.make_waveforms(n_win, n_spk, 1, seed=12, device=cuda) - 1.0
    off = torch.tensor(np.concatenate(([0], np.cumsum(counts))), dtype=torch.int64, device=cuda)
    force_tc("0")
    g0, s0, ref, _ = _ops.pca_gram(w, offsets=off)
    force_tc("1")
    g1, s1, _, _ = _ops.pca_gram(w, offsets=off, ref=ref)
    torch.cuda.synchronize()
    for g in range(len(counts)):
        scale = float(g0[g].abs().max()) + 1e-30
        assert float((g1[g] - g0[g]).abs().max()) <= 5e-6 * scale, g
        assert float((s1[g] - s0[g]).abs().max()) <= 5e-6 * (float(s0[g].abs().max()) + 1.0), g


def test_tc_path_matches_oracle_fetpca(cuda, oracle, force_tc):
    from spikesort_b200 import features, synth
    w = (synth.make_waveforms(25, 80000, 4, seed=13, device=cuda) + 1.0).cpu().numpy()
    sp = {'data': w, 'time': np.arange(25.0), 'FS': 25000}
    ref = oracle.fetPCA(sp, ncomps=2)
    force_tc("1")
    got = features.fetPCA(sp, ncomps=2)
    err = rel_err(sign_normalise(got['data'], ref['data']), ref['data'])
    print("tc fetPCA vs oracle rel err %.3g" % err)
    assert err < 1e-5


@pytest.mark.parametrize("n_win,n_spk,n_c,ncomps", [(25, 300000, 4, 2), (30, 1100000, 1, 3), (31, 524288, 2, 1),
                                                    (25, 350001, 3, 2), (25, 262144, 4, 4)])
def test_vectorised_projection_equals_scalar_kernel(cuda, n_win, n_spk, n_c, ncomps):
    """Big float32 blocks in contact mode take pca_project_vec4_kernel (four elements per thread,
    128-bit loads / stores); the same block widened to float64 takes the scalar kernel.  Same
    products in the same order: the scores are bit-identical.  (350001 x 3 is odd: not a
    multiple of four elements, scalar kernel both times; ncomps = 4: scalar kernel too.)"""
    import torch
    from spikesort_b200 import _ops, synth
    w = synth.make_waveforms(n_win, n_spk, n_c, seed=5, device=cuda) + 1.5
    g, s, ref, counts = _ops.pca_gram(w, mode="f64")
    evals, evecs = _ops.pca_eig(g, s, counts, top=ncomps)
    a = _ops.pca_project(w, evals, evecs, ncomps)
    b = _ops.pca_project(w.double(), evals, evecs, ncomps)
    assert a.shape == (n_spk, ncomps * n_c)
    assert torch.equal(a, b)
    # a view whose rows start 4 bytes off a 16-byte boundary: scalar kernel, same scores
    if n_c == 4:
        big = torch.empty(w.numel() + 1, dtype=torch.float32, device=cuda)
        big[1:] = w.reshape(-1)
        c = _ops.pca_project(big[1:].view(n_win, n_spk, n_c), evals, evecs, ncomps)
        assert torch.equal(a, c)
