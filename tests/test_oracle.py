"""CPU tier: pins the oracle (oracle/port.py + oracle/filtfilt_ref.c) against

* the golden vectors generated from the UNMODIFIED reference files
  (tests/golden/make_golden.py),
* the reference's doctest known answers (docs/source/datastructures.rst:98-109,
  163-199) and unit-test expectations (tests/test_core.py, tests/test_features.py),
* SciPy's filtfilt (bit-exact), the third-party routine the reference calls.
"""
import warnings

import numpy as np
import pytest

import ref_cases
from util import sine_fixture, sign_normalise


@pytest.mark.parametrize("case", ref_cases.EXTRACT_CASES, ids=lambda f: f.__name__)
def test_reference_extract_cases(oracle, case):
    case(oracle)


@pytest.mark.parametrize("case", ref_cases.FEATURE_CASES, ids=lambda f: f.__name__)
def test_reference_feature_cases(oracle, case):
    case(oracle)


def test_reference_filter_cases(oracle):
    ref_cases.check_filter_proxy_shape(oracle)
    ref_cases.check_linear_iir_detect(oracle, oracle)


def test_doc_detect_known_answer(oracle, golden):
    raw = {'data': np.array([[0, 1, 0, 0, 0, 1]]), 'FS': 10000, 'n_contacts': 1}
    spt = oracle.detect_spikes(raw, thresh=0.8)
    assert np.array_equal(spt['data'], [0.0, 0.4])                 # datastructures.rst:109
    assert np.array_equal(spt['data'], golden['doc_detect_spt'])
    assert set(spt.keys()) == {'thresh', 'contact', 'data'}


def test_doc_extract_known_answer(oracle, golden):
    raw = {'data': golden['doc_extract_in'], 'FS': 10000, 'n_contacts': 1}
    w = oracle.extract_spikes(raw, {'data': np.array([0.15, 0.65, 1])}, [0, 0.4])
    assert w['data'].shape == (4, 3, 1) and w['data'].dtype == np.float32
    assert np.array_equal(w['data'][:, :, 0].T, [[1, 1, 0, 0], [1, -1, 0, 0], [0, 0, 0, 0]])
    assert np.allclose(w['time'], [0., 0.1, 0.2, 0.3])
    assert np.array_equal(w['is_valid'], [True, True, False])
    assert np.array_equal(w['data'], golden['doc_extract_waves'])
    assert np.array_equal(w['time'], golden['doc_extract_time'])


def test_sine_detect_align_extract(oracle, golden):
    sp, period, time = sine_fixture()
    d = oracle.detect_spikes(sp, thresh=0.5)
    assert np.array_equal(d['data'], golden['sine_detect'])
    # tests/test_core.py:55-64: crossings within one sample of period/12 + k*period
    expected = period / 12. + np.arange(100) * period
    assert np.all(np.abs(d['data'] - expected) <= 1000.0 / sp['FS'])
    df = oracle.detect_spikes(sp, thresh=-0.5, edge='falling')
    assert np.array_equal(df['data'], golden['sine_detect_falling'])
    sp_win = list(golden['sine_sp_win'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        a = oracle.align_spikes(sp, d, sp_win, type='max')
        assert np.array_equal(a['data'], golden['sine_align_max'])
        a2 = oracle.align_spikes(sp, d, [-period / 24, period / 24], type='max')
        assert np.array_equal(a2['data'], golden['sine_align_short'])
        a3 = oracle.align_spikes(sp, {'data': golden['sine_align_doubles_in']}, sp_win, type='max')
        assert np.array_equal(a3['data'], golden['sine_align_doubles'])
        a4 = oracle.align_spikes(sp, d, sp_win, type='min')
        assert np.array_equal(a4['data'], golden['sine_align_min'])
    w = oracle.extract_spikes(sp, a, sp_win)
    assert np.array_equal(w['data'], golden['sine_extract'])
    assert np.array_equal(w['time'], golden['sine_extract_time'])
    assert not w['is_valid'][0] and w['is_valid'][1:].all()    # first spike is clipped
    we = oracle.extract_spikes(sp, {'data': golden['sine_edge_spt']}, [-0.5, 1.0])
    assert np.array_equal(we['data'], golden['sine_edge_waves'])
    assert np.array_equal(we['is_valid'], golden['sine_edge_valid'])


def test_filter_spt_counts(oracle):
    # tests/test_core.py:176-198: 100 / 99 / 99 spikes fit the window
    sp, period, _ = sine_fixture()
    spt = {'data': np.arange(100) * period + period / 4.}
    assert len(oracle.filter_spt(sp, spt, [-period / 8., period / 8.])) == 100
    assert len(oracle.filter_spt(sp, spt, [-period / 2., period / 8.])) == 99
    assert len(oracle.filter_spt(sp, spt, [-period / 8., period])) == 99


def test_remove_doubles(oracle, golden):
    out = oracle.remove_doubles({'data': golden['rd_in']}, 1000.0 / 25000.0)
    assert np.array_equal(out['data'], golden['rd_out'])
    assert len(oracle.remove_doubles({'data': np.zeros(0)}, 0.04)['data']) == 0


def test_detect_bad_edge(oracle):
    sp, _, _ = sine_fixture()
    with pytest.raises(TypeError):
        oracle.detect_spikes(sp, thresh=0.5, edge='sideways')


@pytest.mark.parametrize("dtype", [np.int16, np.float32, np.float64])
def test_filtfilt_bit_exact_vs_scipy(oracle, dtype):
    from scipy import signal
    rng = np.random.default_rng(3)
    designs = [signal.iirdesign(800 * 2 / 25000, 100 * 2 / 25000, 1, 10, ftype='butter'),
               signal.iirdesign(np.array([300, 6000]) * 2 / 30000, np.array([100, 8000]) * 2 / 30000,
                                1, 10, ftype='butter'),
               signal.iirdesign(0.02, 0.016, 1, 10, ftype='ellip')]
    for b, a in designs:
        x = (rng.standard_normal(5000) * 3000).astype(dtype)
        if dtype == np.int16:
            x[0], x[-1] = 30000, -30000          # odd extension wraps around in int16
        assert np.array_equal(signal.filtfilt(b, a, x), oracle.filtfilt(b, a, x))
        assert np.array_equal(signal.lfilter_zi(b, a), oracle.lfilter_zi(b, a))
    with pytest.raises(ValueError):
        oracle.filtfilt(designs[1][0], designs[1][1], np.zeros(27))     # len <= padlen


def test_filter_chain_golden(oracle, golden):
    FS = 25000
    raw = {'data': golden['chain_raw'], 'FS': FS, 'n_contacts': 2}
    f = oracle.filter_proxy(raw, oracle.Filter(800., 100.), chunksize=25000)
    assert f['data'].dtype == np.float64
    assert np.array_equal(f['data'], golden['chain_filtered'])
    # generic-callable path of filter_proxy gives the same numbers
    filt = oracle.Filter(800., 100.)
    f2 = oracle.filter_proxy(raw, lambda x, fs: filt(x, fs), chunksize=25000)
    assert np.array_equal(f2['data'], f['data'])
    det = oracle.detect_spikes(f, contact=1, thresh='auto')
    assert det['thresh'] == golden['chain_thresh'][0]
    assert np.array_equal(det['data'], golden['chain_detect'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        al = oracle.align_spikes(f, det, [-0.2, 0.8], type='max', contact=1)
    assert np.array_equal(al['data'], golden['chain_align'])
    wv = oracle.extract_spikes(f, al, [-0.2, 0.8])
    assert np.array_equal(wv['data'], golden['chain_waves'])
    pc = oracle.fetPCA(wv, ncomps=2)
    assert pc['names'] == ['Ch0:PC0', 'Ch0:PC1', 'Ch1:PC0', 'Ch1:PC1']
    assert np.allclose(sign_normalise(pc['data'], golden['chain_pca']), golden['chain_pca'],
                       rtol=0, atol=1e-9 * np.abs(golden['chain_pca']).max())
    ev, evec, sc = oracle.PCA(wv['data'][:, :, 0], 3)
    assert np.allclose(ev, golden['chain_pca0_evals'], rtol=1e-10, atol=1e-12 * ev.max())


def test_resample_golden(oracle, golden):
    """Spline upsampling (extract.py:179-205,245-258) - bit-exact: the port calls the same
    SciPy splrep/splev as the reference."""
    FS = 25000
    fd = {'data': golden['chain_filtered'], 'FS': FS, 'n_contacts': 2}
    det = {'data': golden['chain_detect']}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        al = oracle.align_spikes(fd, det, [-0.2, 0.8], type='max', contact=1, resample=10)
        al4 = oracle.align_spikes(fd, det, [-0.2, 0.8], type='min', contact=0, resample=4,
                                  remove=False)
        wv = oracle.extract_spikes(fd, al, [-0.2, 0.8], resample=3)
    assert np.array_equal(al['data'], golden['chain_align_rs10'])
    assert np.array_equal(al4['data'], golden['chain_align_rs4_min'])
    assert np.array_equal(wv['data'], golden['chain_waves_rs3'])
    assert np.array_equal(wv['time'], golden['chain_waves_rs3_time'])
    assert sorted(wv.keys()) == list(golden['chain_waves_rs3_keys'])
    sp, period, _ = sine_fixture()
    sw = oracle.extract_spikes(sp, {'data': golden['sine_resample_spt']}, [0, period])
    sr = oracle.resample_spikes(sw, sp['FS'] * 2)
    assert np.array_equal(sr['data'], golden['sine_resample'])
    assert np.array_equal(sr['time'], golden['sine_resample_time'])
    assert sr['FS'] == sp['FS']


def test_fir_and_zerophase_golden(oracle, golden):
    """FilterFir (filters.py:41-74) and ZeroPhaseFilter (filters.py:10-38) through
    filter_proxy, bit-exact against the reference's output."""
    raw = {'data': golden['fir_raw'], 'FS': 25000, 'n_contacts': 2}
    for key, flt in (('fir_hp_filtered', oracle.FilterFir(300., 100., 101)),
                     ('fir_lp_filtered', oracle.FilterFir(3000., 5000., 64)),
                     ('zpf_filtered', oracle.ZeroPhaseFilter('ellip', [300., 3000.]))):
        f = oracle.filter_proxy(raw, flt, chunksize=6000)
        assert np.array_equal(f['data'], golden[key]), key
    with pytest.raises(ValueError):
        oracle.FilterFir(300., 100., 101)(golden['fir_raw'][0, :303], 25000)   # len <= padlen


def test_bandpass_float32_golden(oracle, golden):
    raw = {'data': golden['bp_raw'], 'FS': 30000, 'n_contacts': 1}
    f = oracle.filter_proxy(raw, oracle.Filter(np.array([300., 6000.]), np.array([100., 8000.])))
    assert np.array_equal(f['data'], golden['bp_filtered'])


def test_sine_highpass_100_crossings(oracle, golden):
    # tests/test_core.py:33-43
    sp, period, _ = sine_fixture()
    sp['data'] = sp['data'] + 2
    f = oracle.fltLinearIIR(sp, 1000.0 / period * 0.5, 1000.0 / period * 0.4, 1, 10, 'ellip')
    assert np.array_equal(f['data'], golden['sine_hp_filtered'])
    spt = oracle.detect_spikes(f, thresh=0.5)
    assert len(spt['data']) == 100
    assert np.array_equal(spt['data'], golden['sine_hp_detect'])


def test_pca_whitening(oracle):
    # tests/test_features.py:38-51: projected covariance is the identity
    rng = np.random.default_rng(0)
    data = rng.standard_normal((5, 2000)) * np.array([5, 3, 1, .5, .1])[:, None]
    data = (rng.standard_normal((5, 5)) @ data)
    _, _, score = oracle.PCA(data, 2)
    assert np.mean((np.cov(score) - np.eye(2)) ** 2) < 0.01


def test_fetpca_mask_and_contacts(oracle):
    rng = np.random.default_rng(1)
    waves = {'data': rng.standard_normal((20, 50, 3)).astype(np.float32), 'time': np.arange(20.),
             'FS': 25000, 'is_valid': rng.random(50) > 0.2}
    f = oracle.fetPCA(waves, ncomps=2, contacts=[0, 2])
    assert f['data'].shape == (50, 4) and np.array_equal(f['is_valid'], waves['is_valid'])
    with pytest.raises(IndexError):
        oracle.fetPCA(waves, contacts=[5])


def test_cheap_features_and_combine_golden(oracle, golden):
    wv = {'data': golden['chain_waves'], 'time': np.arange(25.0), 'FS': 25000,
          'is_valid': np.arange(golden['chain_waves'].shape[1]) % 7 != 0}
    p2p = oracle.fetP2P(wv)
    assert p2p['data'].dtype == np.float32 and np.array_equal(p2p['data'], golden['chain_p2p'])
    assert p2p['names'] == ['Ch0:P2P', 'Ch1:P2P'] and np.array_equal(p2p['is_valid'], wv['is_valid'])
    assert np.array_equal(oracle.fetMin(wv)['data'], golden['chain_min'])
    assert oracle.fetP2P(wv, contacts=[1])['data'].shape == (wv['data'].shape[1], 1)
    comb = oracle.combine((oracle.fetP2P(wv), oracle.fetPCA(wv, ncomps=2)), norm=True,
                          feat_method_names=['P2P', 'PCA'])
    assert comb['names'] == list(golden['chain_combined_names'])
    assert comb['names'][0] == 'P2P:Ch0:P2P' and comb['names'][-1] == 'PCA:Ch1:PC1'
    # the P2P columns are bit-exact; the PCA columns up to eigenvector sign (x -> 1 - x)
    g = golden['chain_combined']
    assert np.array_equal(comb['data'][:, :2], g[:, :2])
    for j in range(2, 6):
        d = min(np.abs(comb['data'][:, j] - g[:, j]).max(), np.abs(1 - comb['data'][:, j] - g[:, j]).max())
        assert d < 1e-9
    assert comb['data'].min() == 0.0 and comb['data'].max() == 1.0
    with pytest.raises(ValueError):
        oracle.combine(({'data': np.zeros((3, 1)), 'names': ['a']},
                        {'data': np.zeros((4, 1)), 'names': ['b']}))
