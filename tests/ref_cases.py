"""The reference's own hot-path unit tests (``/root/reference/tests/test_core.py``
TestFilter/TestExtract and ``tests/test_features.py`` PCA cases), restated so the
SAME checks run against any implementation of the spike_sort.core API: the CPU
oracle (tests/test_oracle.py) and the CUDA drop-in (tests/test_gpu_dropin.py).

`ex`, `flt`, `fet` are module-like objects exposing the reference's function names.
"""
import warnings

import numpy as np


class SineCase(object):
    """tests/test_core.py:45-53."""

    def __init__(self):
        self.n_spikes = 100
        self.FS = 25E3
        self.period = 1000.0 / self.FS * 100
        self.time = np.arange(0, self.period * self.n_spikes, 1000.0 / self.FS)
        self.spikes = np.sin(2 * np.pi / self.period * self.time)[np.newaxis, :]
        self.spk_data = {"data": self.spikes, "n_contacts": 1, "FS": self.FS}


def _quiet(fn, *a, **k):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


def check_detect(ex):                                   # test_core.py:55-64
    c = SineCase()
    threshold = 0.5
    crossings_real = c.period / 12.0 + np.arange(c.n_spikes) * c.period
    spt = ex.detect_spikes(c.spk_data, thresh=threshold)
    assert (np.abs(spt['data'] - crossings_real) <= 1000.0 / c.FS).all()


def check_align(ex):                                    # test_core.py:66-73
    c = SineCase()
    maxima_idx = c.period * (1 / 4.0 + np.arange(c.n_spikes))
    thr_crossings = c.period * (1 / 6.0 + np.arange(c.n_spikes))
    sp_win = [-c.period / 6.0, c.period / 3.0]
    spt = _quiet(ex.align_spikes, c.spk_data, {"data": thr_crossings}, sp_win)
    assert (np.abs(spt['data'] - maxima_idx) <= 1000.0 / c.FS).all()


def check_align_short_win(ex):                          # test_core.py:75-82
    c = SineCase()
    maxima_idx = c.period * (1 / 4.0 + np.arange(c.n_spikes))
    thr_crossings = c.period * (1 / 6.0 + np.arange(c.n_spikes))
    sp_win = [-c.period / 24.0, c.period / 12.0]
    spt = _quiet(ex.align_spikes, c.spk_data, {"data": thr_crossings}, sp_win)
    assert (np.abs(spt['data'] - maxima_idx) <= 1000.0 / c.FS).all()


def check_align_edge(ex):                               # test_core.py:84-94
    c = SineCase()
    spikes = np.sin(2 * np.pi / c.period * c.time + np.pi / 2.0)[np.newaxis, :]
    thr_crossings = c.period * (-1 / 6.0 + np.arange(1, c.n_spikes + 1))
    sp_win = [-c.period / 24.0, c.period / 12.0]
    spk_data = {"data": spikes, "n_contacts": 1, "FS": c.FS}
    spt = _quiet(ex.align_spikes, spk_data, {"data": thr_crossings}, sp_win)
    last = spt['data'][-1]
    assert (last >= (c.time[-1] - sp_win[1])) & (last <= c.time[-1])


def check_align_double_spikes(ex):                      # test_core.py:96-103
    c = SineCase()
    maxima_idx = c.period * (1 / 4.0 + np.arange(c.n_spikes))
    thr_crossings = c.period * (1 / 6.0 + np.arange(0, c.n_spikes, 0.5))
    sp_win = [-c.period / 24.0, c.period / 12.0]
    spt = _quiet(ex.align_spikes, c.spk_data, {"data": thr_crossings}, sp_win)
    assert (np.abs(spt['data'] - maxima_idx) <= 1000.0 / c.FS).all()


def check_remove_doubles_roundoff(ex):                  # test_core.py:105-116
    c = SineCase()
    tol = 1000.0 / c.FS
    data = [1.0, 1.0 + tol + tol * 0.01]
    clean = ex.remove_doubles({"data": np.array(data)}, tol)
    assert len(clean['data']) == 1


def check_extract(ex):                                  # test_core.py:118-125
    c = SineCase()
    zero_crossing = c.period * np.arange(c.n_spikes)
    zero_crossing += 1000.0 / c.FS / 2.0
    sp_win = [0, c.period]
    sp_waves = ex.extract_spikes(c.spk_data, {"data": zero_crossing}, sp_win)
    ref_sp = np.sin(2 * np.pi / c.period * sp_waves['time'])
    assert (np.abs(ref_sp[:, np.newaxis] - sp_waves['data'][:, :, 0]) < 1E-6).all()


def check_extract_and_resample(ex):                     # test_core.py:138-146
    c = SineCase()
    zero_crossing = c.period * np.arange(c.n_spikes)
    zero_crossing += 1000.0 / c.FS / 2.0
    sp_win = [0, c.period]
    sp_waves = ex.extract_spikes(c.spk_data, {"data": zero_crossing}, sp_win)
    sp_resamp = ex.resample_spikes(sp_waves, c.FS * 2)
    ref_sp = np.sin(2 * np.pi / c.period * sp_resamp['time'])
    assert (np.abs(ref_sp[:, np.newaxis] - sp_resamp['data'][:, :, 0]) < 1E-6).all()


def check_mask_of_truncated_spikes(ex):                 # test_core.py:148-174
    c = SineCase()
    sp_win = [0, c.period]
    ok = ex.extract_spikes(c.spk_data, {"data": np.array([c.period])}, sp_win)
    assert 'is_valid' not in ok
    spt = {"data": np.array([c.period, c.time[-1] - c.period / 2.0, -c.period / 2.0 + 0.02])}
    w = ex.extract_spikes(c.spk_data, spt, sp_win)
    assert np.array_equal(w['is_valid'], [True, False, False])
    n_half = int(round(len(w['time']) / 2.0))
    # end-truncated spike: second half is zero; begin-truncated: first half is zero
    assert (w['data'][n_half + 1:, 1, 0] == 0).all() and (w['data'][:n_half - 1, 1, 0] != 0).any()
    assert (w['data'][:n_half - 1, 2, 0] == 0).all() and (w['data'][n_half + 1:, 2, 0] != 0).any()


def check_filter_spt(ex):                               # test_core.py:176-198
    c = SineCase()
    spt = {"data": np.arange(c.n_spikes) * c.period + c.period / 4.}
    assert len(ex.filter_spt(c.spk_data, spt, [-c.period / 8., c.period / 8.])) == c.n_spikes
    assert len(ex.filter_spt(c.spk_data, spt, [-c.period / 2., c.period / 8.])) == c.n_spikes - 1
    assert len(ex.filter_spt(c.spk_data, spt, [-c.period / 8., c.period])) == c.n_spikes - 1


def check_filter_proxy_shape(flt):                      # test_core.py:19-23
    c = SineCase()
    sp_freq = 1000.0 / c.period
    filt = flt.Filter(sp_freq * 0.5, sp_freq * 0.4, 1, 10, 'ellip')
    out = flt.filter_proxy(c.spk_data, filt)
    assert c.spk_data['data'].shape == out['data'].shape


def check_linear_iir_detect(flt, ex):                   # test_core.py:33-43
    c = SineCase()
    sp_freq = 1000.0 / c.period
    c.spk_data['data'] = c.spk_data['data'] + 2
    out = flt.fltLinearIIR(c.spk_data, sp_freq * 0.5, sp_freq * 0.4, 1, 10, 'ellip')
    spt = ex.detect_spikes(out, thresh=0.5)
    assert len(spt['data']) == c.n_spikes


def check_pca_whitening(fet):                           # test_features.py:38-51
    rng = np.random.RandomState(1221)
    n_dim, n_obs = 2, 2000
    A = np.array([[2.0, 1.0], [1.0, 3.0]])
    data = np.dot(A, rng.randn(n_dim, n_obs)).astype(np.float32)
    evals, evecs, score = fet.PCA(data, n_dim)
    assert ((np.cov(score) - np.eye(n_dim)) ** 2).mean() < 0.01


def check_fetpca_separates(fet):                        # test_features.py:53-70
    rng = np.random.RandomState(3)
    n_pts, n_spk = 30, 200
    tt = np.linspace(0, 1, n_pts)
    t1, t2 = np.sin(2 * np.pi * tt), np.sin(4 * np.pi * tt)
    labels = rng.rand(n_spk) > 0.5
    waves = np.where(labels[None, :], t1[:, None], t2[:, None]) + 0.01 * rng.randn(n_pts, n_spk)
    sp = {'data': waves[:, :, None].astype(np.float32), 'time': tt, 'FS': 1}
    f = fet.fetPCA(sp, ncomps=1)
    pc = f['data'][:, 0]
    a, b = pc[labels], pc[~labels]
    assert abs(a.mean() - b.mean()) > 5 * max(a.std(), b.std())
    assert f['names'] == ['Ch0:PC0']


def check_fetpca_multichannel_labels(fet):              # test_features.py:72-100
    rng = np.random.RandomState(5)
    n_pts, n_spk, n_c = 20, 300, 3
    base = rng.randn(n_pts, n_spk).astype(np.float32)
    data = np.stack([base * (k + 1) for k in range(n_c)], axis=2)
    sp = {'data': data, 'time': np.arange(n_pts, dtype=float), 'FS': 1,
          'is_valid': rng.rand(n_spk) > 0.1}
    f = fet.fetPCA(sp, ncomps=2)
    assert f['names'] == ['Ch0:PC0', 'Ch0:PC1', 'Ch1:PC0', 'Ch1:PC1', 'Ch2:PC0', 'Ch2:PC1']
    assert f['data'].shape == (n_spk, 6)
    # scaling a channel does not change whitened scores (up to sign)
    for k in range(1, n_c):
        for j in range(2):
            r = f['data'][:, 2 * k + j]
            q = f['data'][:, j]
            assert min(np.abs(r - q).max(), np.abs(r + q).max()) < 1e-6 * np.abs(q).max()
    assert np.array_equal(f['is_valid'], sp['is_valid'])          # add_mask, :209-217
    sub = fet.fetPCA(sp, ncomps=1, contacts=[2])
    assert sub['data'].shape == (n_spk, 1)


EXTRACT_CASES = [check_detect, check_align, check_align_short_win, check_align_edge,
                 check_align_double_spikes, check_remove_doubles_roundoff, check_extract,
                 check_extract_and_resample, check_mask_of_truncated_spikes, check_filter_spt]
FEATURE_CASES = [check_pca_whitening, check_fetpca_separates, check_fetpca_multichannel_labels]
