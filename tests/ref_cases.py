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


class FeatureCase(object):
    """tests/test_features.py:9-23 (`setup`): two template "cells" of 100 points, integer
    valued, as a (n_pts, 2, 1) waveform block.  (`np.int` is `int` in today's NumPy.)"""

    def __init__(self):
        self.gain = 1000
        n_pts = 100
        FS = 5E3
        time = np.arange(n_pts, dtype=np.float32) / n_pts - 0.5
        cells = self.gain * np.vstack((np.sin(time * 2 * np.pi) / 2.0, np.abs(time) * 2 - 1))
        self.cells = cells.astype(int)
        self.spikes_dict = {"data": self.cells.T[:, :, np.newaxis], "time": time, "FS": FS}


def check_fetP2P(fet):                                  # test_features.py:25-36
    c = FeatureCase()
    spikes_dict = c.spikes_dict.copy()
    n_spikes = 200
    amps = np.random.RandomState(7).randint(1, 100, n_spikes)[:, np.newaxis]
    spikes = amps * c.cells[0, :]                       # int64 waveforms, as in the reference
    spikes_dict['data'] = spikes.T[:, :, np.newaxis]
    p2p = fet.fetP2P(spikes_dict)
    assert (p2p['data'] == amps * c.gain).all()


def check_pca_whitening(fet):                           # test_features.py:38-51 (float64 input)
    rng = np.random.RandomState(1221)
    n_dim, n_obs = 2, 100
    raw_data = rng.randn(n_dim, n_obs)
    mixing = np.array([[-1, 1], [1, 1]])
    mix_data = np.dot(mixing, raw_data)
    mix_data = mix_data - mix_data.mean(1)[:, np.newaxis]
    evals, evecs, score = fet.PCA(mix_data, n_dim)
    proj_cov = np.cov(score)
    assert np.mean((proj_cov - np.eye(n_dim)) ** 2) < 0.01


def check_fetpca_separates(fet):                        # test_features.py:53-70
    c = FeatureCase()
    rng = np.random.RandomState(3)
    spikes_dict = c.spikes_dict.copy()
    n_spikes = 200
    amp_var_fact = 0.4
    amp_var = 1 + amp_var_fact * rng.rand(n_spikes, 1)
    _amps = rng.rand(n_spikes) > 0.5
    amps = amp_var * np.vstack((_amps >= 0.5, _amps < 0.5)).T
    spikes = np.dot(amps, c.cells).T
    spikes = spikes.astype(np.float32)                  # the reference's own cast
    spikes_dict['data'] = spikes[:, :, np.newaxis]
    pcs = fet.fetPCA(spikes_dict, ncomps=1)['data']
    compare = ~np.logical_xor(pcs[:, 0].astype(int) + 1, _amps)
    # the sign of a principal axis is arbitrary (LAPACK's too): the reference's check holds
    # for one orientation, its mirror image for the other
    mirrored = ~np.logical_xor((-pcs[:, 0]).astype(int) + 1, _amps)
    assert n_spikes in (np.sum(compare), np.sum(mirrored))


def check_fetpca_multichannel_labels(fet):              # test_features.py:72-100 (float64 block)
    c = FeatureCase()
    spike1, spike2 = c.cells
    ch0_spikes = np.vstack((spike1, spike2, spike1 + spike2)).T
    ch1_spikes = np.vstack((spike1 * 2, spike2 / 2., spike1 * 2 - spike2 / 2.)).T
    spikes = np.empty((ch0_spikes.shape[0], ch0_spikes.shape[1], 2))
    spikes[:, :, 0] = ch0_spikes
    spikes[:, :, 1] = ch1_spikes
    features = fet.fetPCA({'data': spikes}, ncomps=3)
    names = features['names']
    ch0_feats = np.empty((3, 3))
    ch1_feats = np.empty((3, 3))
    for fidx in range(3):
        i0 = names.index('Ch%d:PC%d' % (0, fidx))
        i1 = names.index('Ch%d:PC%d' % (1, fidx))
        ch0_feats[fidx, :] = features['data'][:, i0]
        ch1_feats[fidx, :] = features['data'][:, i1]
    np.testing.assert_array_almost_equal(ch0_feats[:, 2], ch0_feats[:, 0] + ch0_feats[:, 1])
    np.testing.assert_array_almost_equal(ch1_feats[:, 2], ch1_feats[:, 0] - ch1_feats[:, 1])


def check_fetpca_mask_and_subset(fet):                  # features.py:172-181,298-307 via fetPCA
    rng = np.random.RandomState(5)
    n_pts, n_spk, n_c = 20, 300, 3
    base = rng.randn(n_pts, n_spk)
    data = np.stack([base * (k + 1) for k in range(n_c)], axis=2)       # float64
    sp = {'data': data, 'time': np.arange(n_pts, dtype=float), 'FS': 1,
          'is_valid': rng.rand(n_spk) > 0.1}
    f = fet.fetPCA(sp, ncomps=2)
    assert f['names'] == ['Ch0:PC0', 'Ch0:PC1', 'Ch1:PC0', 'Ch1:PC1', 'Ch2:PC0', 'Ch2:PC1']
    assert f['data'].shape == (n_spk, 6)
    for k in range(1, n_c):                             # scaling a channel: same whitened scores
        for j in range(2):
            r, q = f['data'][:, 2 * k + j], f['data'][:, j]
            assert min(np.abs(r - q).max(), np.abs(r + q).max()) < 1e-6 * np.abs(q).max()
    assert np.array_equal(f['is_valid'], sp['is_valid'])
    sub = fet.fetPCA(sp, ncomps=1, contacts=[2])
    assert sub['data'].shape == (n_spk, 1)
    try:
        fet.fetPCA(sp, contacts=[7])
    except IndexError:
        pass
    else:
        raise AssertionError("contacts=[7] on 3 contacts must raise IndexError")


EXTRACT_CASES = [check_detect, check_align, check_align_short_win, check_align_edge,
                 check_align_double_spikes, check_remove_doubles_roundoff, check_extract,
                 check_extract_and_resample, check_mask_of_truncated_spikes, check_filter_spt]
FEATURE_CASES = [check_fetP2P, check_pca_whitening, check_fetpca_separates,
                 check_fetpca_multichannel_labels, check_fetpca_mask_and_subset]
