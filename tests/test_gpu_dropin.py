"""GPU tier: the CUDA drop-in (spikesort_b200.{filters,extract,features}) against

* the reference's own unit tests restated in tests/ref_cases.py,
* the golden vectors produced by the unmodified reference (tests/golden/),
through the public spike_sort.core-compatible API, i.e. through the C ABI.

Bars: bit-exact for spike times, indices and waveforms; rel 1e-5 (north_star) for the
filtered signal and PCA scores - the asserts below use the much tighter bounds the
kernels actually meet, and say so.
"""
import warnings

import numpy as np
import pytest

import ref_cases
from util import sine_fixture, sign_normalise, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ss(cuda):
    import spikesort_b200
    from spikesort_b200 import _lib
    _lib.load()                       # must be the native library, no fallback
    return spikesort_b200


def test_native_library_is_loaded(ss):
    from spikesort_b200 import _lib
    assert _lib._lib is not None
    maps = open("/proc/self/maps").read()
    assert "libspikesort_b200.so" in maps


@pytest.mark.parametrize("case", ref_cases.EXTRACT_CASES, ids=lambda f: f.__name__)
def test_reference_extract_cases(ss, case):
    case(ss.extract)


@pytest.mark.parametrize("case", ref_cases.FEATURE_CASES, ids=lambda f: f.__name__)
def test_reference_feature_cases(ss, case):
    case(ss.features)


def test_reference_filter_cases(ss):
    ref_cases.check_filter_proxy_shape(ss.filters)
    ref_cases.check_linear_iir_detect(ss.filters, ss.extract)


def test_doc_known_answers(ss, golden):
    raw = {'data': np.array([[0, 1, 0, 0, 0, 1]]), 'FS': 10000, 'n_contacts': 1}
    spt = ss.extract.detect_spikes(raw, thresh=0.8)
    assert np.array_equal(spt['data'], [0.0, 0.4]) and spt['data'].dtype == np.float64
    assert set(spt.keys()) == {'thresh', 'contact', 'data'}
    raw = {'data': golden['doc_extract_in'], 'FS': 10000, 'n_contacts': 1}
    w = ss.extract.extract_spikes(raw, {'data': np.array([0.15, 0.65, 1])}, [0, 0.4])
    assert sorted(w.keys()) == ['FS', 'data', 'is_valid', 'time']
    assert w['data'].dtype == np.float32 and w['data'].shape == (4, 3, 1)
    assert np.array_equal(w['data'], golden['doc_extract_waves'])
    assert np.array_equal(w['time'], golden['doc_extract_time'])
    assert np.array_equal(w['is_valid'], [True, True, False])


def test_sine_golden_bit_exact(ss, golden):
    ex = ss.extract
    sp, period, time = sine_fixture()
    d = ex.detect_spikes(sp, thresh=0.5)
    assert np.array_equal(d['data'], golden['sine_detect'])
    df = ex.detect_spikes(sp, thresh=-0.5, edge='falling')
    assert np.array_equal(df['data'], golden['sine_detect_falling'])
    sp_win = list(golden['sine_sp_win'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        a = ex.align_spikes(sp, d, sp_win, type='max')
        assert np.array_equal(a['data'], golden['sine_align_max'])
        a2 = ex.align_spikes(sp, d, [-period / 24, period / 24], type='max')
        assert np.array_equal(a2['data'], golden['sine_align_short'])
        a3 = ex.align_spikes(sp, {'data': golden['sine_align_doubles_in']}, sp_win, type='max')
        assert np.array_equal(a3['data'], golden['sine_align_doubles'])
        a4 = ex.align_spikes(sp, d, sp_win, type='min')
        assert np.array_equal(a4['data'], golden['sine_align_min'])
    w = ex.extract_spikes(sp, a, sp_win)
    assert np.array_equal(w['data'], golden['sine_extract'])
    assert np.array_equal(w['time'], golden['sine_extract_time'])
    we = ex.extract_spikes(sp, {'data': golden['sine_edge_spt']}, [-0.5, 1.0])
    assert np.array_equal(we['data'], golden['sine_edge_waves'])
    assert np.array_equal(we['is_valid'], golden['sine_edge_valid'])
    out = ex.remove_doubles({'data': golden['rd_in']}, 1000.0 / 25000.0)
    assert np.array_equal(out['data'], golden['rd_out'])


def test_chain_golden(ss, golden):
    FS = 25000
    raw = {'data': golden['chain_raw'], 'FS': FS, 'n_contacts': 2}
    f = ss.filters.filter_proxy(raw, ss.filters.Filter(800., 100.), chunksize=25000)
    assert f['data'].shape == (2, 60000) and f['data'].dtype == np.float64
    got = np.asarray(f['data'])
    assert rel_err(got, golden['chain_filtered']) < 1e-11          # bar: 1e-5
    # spike indices / waveforms: bit-exact GIVEN the reference's filtered signal
    fd = {'data': golden['chain_filtered'], 'FS': FS, 'n_contacts': 2}
    det = ss.extract.detect_spikes(fd, contact=1, thresh='auto')
    assert abs(det['thresh'] - golden['chain_thresh'][0]) <= 1e-12 * golden['chain_thresh'][0]
    det = ss.extract.detect_spikes(fd, contact=1, thresh=float(golden['chain_thresh'][0]))
    assert np.array_equal(det['data'], golden['chain_detect'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        al = ss.extract.align_spikes(fd, det, [-0.2, 0.8], type='max', contact=1)
    assert np.array_equal(al['data'], golden['chain_align'])
    wv = ss.extract.extract_spikes(fd, al, [-0.2, 0.8])
    assert np.array_equal(wv['data'], golden['chain_waves'])
    pc = ss.features.fetPCA(wv, ncomps=2)
    assert pc['names'] == ['Ch0:PC0', 'Ch0:PC1', 'Ch1:PC0', 'Ch1:PC1']
    assert rel_err(sign_normalise(pc['data'], golden['chain_pca']), golden['chain_pca']) < 1e-9
    ev, evec, sc = ss.features.PCA(wv['data'][:, :, 0], 3)
    assert np.allclose(ev[:10], golden['chain_pca0_evals'][:10], rtol=1e-9)
    assert rel_err(sign_normalise(sc.T, golden['chain_pca0_score'].T), golden['chain_pca0_score'].T) < 1e-9
    # end to end on the GPU-filtered signal: count spike-time mismatches (north_star)
    det_g = ss.extract.detect_spikes(f, contact=1, thresh='auto')
    mism = len(np.setxor1d(det_g['data'], golden['chain_detect']))
    print("end-to-end spike-time mismatches: %d of %d" % (mism, len(golden['chain_detect'])))
    assert mism == 0


def test_resample_golden(ss, golden):
    """8f rank 1: spline upsampling.  The kernels replay FITPACK's operations in order and
    unfused: upsampled waveforms and the times aligned on them are BIT-exact."""
    FS = 25000
    fd = {'data': golden['chain_filtered'], 'FS': FS, 'n_contacts': 2}
    det = {'data': golden['chain_detect']}
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        al = ss.extract.align_spikes(fd, det, [-0.2, 0.8], type='max', contact=1, resample=10)
        al4 = ss.extract.align_spikes(fd, det, [-0.2, 0.8], type='min', contact=0, resample=4,
                                      remove=False)
        wv = ss.extract.extract_spikes(fd, {'data': golden['chain_align_rs10']}, [-0.2, 0.8],
                                       resample=3)
        assert any(issubclass(x.category, DeprecationWarning) for x in w)
    for got, key in ((al, 'chain_align_rs10'), (al4, 'chain_align_rs4_min')):
        assert got['data'].shape == golden[key].shape
        mism = int(np.sum(got['data'] != golden[key]))
        print("%s: %d of %d aligned times differ" % (key, mism, len(golden[key])))
        assert mism == 0
    assert wv['data'].dtype == np.float64 and wv['data'].shape == golden['chain_waves_rs3'].shape
    assert np.array_equal(wv['data'], golden['chain_waves_rs3'])
    assert np.array_equal(wv['time'], golden['chain_waves_rs3_time'])
    assert sorted(wv.keys()) == list(golden['chain_waves_rs3_keys'])      # no 'is_valid'
    sp, period, _ = sine_fixture()
    sw = ss.extract.extract_spikes(sp, {'data': golden['sine_resample_spt']}, [0, period])
    sr = ss.extract.resample_spikes(sw, sp['FS'] * 2)
    assert np.array_equal(sr['data'], golden['sine_resample'])
    assert np.array_equal(sr['time'], golden['sine_resample_time'])
    assert sr['FS'] == sp['FS']
    # float64 waveforms in (a user-built dict) take the same route
    sr64 = ss.extract.resample_spikes({'data': sw['data'].astype(np.float64), 'time': sw['time'],
                                       'FS': sw['FS']}, sp['FS'] * 2)
    assert np.array_equal(sr64['data'], sr['data'])
    empty = ss.extract.resample_spikes({'data': np.zeros((len(sw['time']), 0, 1), np.float32),
                                        'time': sw['time'], 'FS': sw['FS']}, sp['FS'] * 2)
    assert empty['data'].shape == (len(sr['time']), 0, 1)


def test_fir_and_zerophase_golden(ss, golden):
    """8f rank 4: FilterFir (two-stage FIR stencil) and ZeroPhaseFilter (the IIR scan) through
    filter_proxy against the reference's output; bar rel 1e-5, asserted far tighter (the
    elliptic band-pass is limited by the round-off of the REFERENCE's (b, a)-form recursion,
    as for the high orders of test_filtfilt_vs_oracle)."""
    raw = {'data': golden['fir_raw'], 'FS': 25000, 'n_contacts': 2}
    for key, flt, tol in (('fir_hp_filtered', ss.filters.FilterFir(300., 100., 101), 1e-13),
                          ('fir_lp_filtered', ss.filters.FilterFir(3000., 5000., 64), 1e-13),
                          ('zpf_filtered', ss.filters.ZeroPhaseFilter('ellip', [300., 3000.]), 1e-6)):
        f = ss.filters.filter_proxy(raw, flt, chunksize=6000)
        got = np.asarray(f['data'])
        assert got.shape == golden[key].shape and got.dtype == np.float64
        err = rel_err(got, golden[key])
        print("%s: rel err %.2e" % (key, err))
        assert err < tol, key
    fir = ss.filters.FilterFir(300., 100., 101)
    one = fir(golden['fir_raw'][0, :6000], 25000)                   # the callable form
    assert rel_err(one, golden['fir_hp_filtered'][0, :6000]) < 1e-13
    with pytest.raises(ValueError):
        fir(golden['fir_raw'][0, :303], 25000)                      # len <= padlen, as SciPy
    # float32 / float64 recordings and the chain on an FIR-filtered signal
    for dt in (np.float32, np.float64):
        rawd = {'data': golden['fir_raw'].astype(dt) * 0.5, 'FS': 25000, 'n_contacts': 2}
        import oracle.port as port
        ref = port.filter_proxy(rawd, port.FilterFir(300., 100., 101), chunksize=6000)['data']
        got = np.asarray(ss.filters.filter_proxy(rawd, fir, chunksize=6000)['data'])
        assert rel_err(got, ref) < 1e-13
    out = ss.sort_frontend(raw, filt=fir, sp_win=[-0.2, 0.8], edge='falling', align_type='min',
                           thresh='4', ncomps=2, chunksize=6000)
    f = ss.filters.filter_proxy(raw, fir, chunksize=6000)
    for c in range(2):
        spt = ss.extract.detect_spikes(f, thresh='4', edge='falling', contact=c)
        assert np.array_equal(out['spt'][c]['data'], spt['data'])
        assert len(spt['data']) > 5


def test_bandpass_float32_golden(ss, golden):
    raw = {'data': golden['bp_raw'], 'FS': 30000, 'n_contacts': 1}
    f = ss.filters.filter_proxy(raw, ss.filters.Filter(np.array([300., 6000.]),
                                                        np.array([100., 8000.])))
    assert rel_err(np.asarray(f['data']), golden['bp_filtered']) < 1e-8   # bar: 1e-5


def test_sine_highpass_golden(ss, golden):
    sp, period, _ = sine_fixture()
    sp['data'] = sp['data'] + 2
    f = ss.filters.fltLinearIIR(sp, 1000.0 / period * 0.5, 1000.0 / period * 0.4, 1, 10, 'ellip')
    assert rel_err(np.asarray(f['data']), golden['sine_hp_filtered']) < 1e-9
    spt = ss.extract.detect_spikes(f, thresh=0.5)
    assert np.array_equal(spt['data'], golden['sine_hp_detect'])


def test_errors_and_warnings(ss):
    sp, period, _ = sine_fixture()
    with pytest.raises(TypeError):
        ss.extract.detect_spikes(sp, thresh=0.5, edge='sideways')
    with pytest.raises(TypeError):
        ss.filters.filter_proxy(sp, lambda x, fs: x)          # no CPU fallback
    assert ss.filters.filter_proxy(sp, None) is sp
    with pytest.raises(ValueError):                           # chunk not longer than padlen
        ss.filters.filter_proxy({'data': np.zeros((1, 5)), 'FS': 25000, 'n_contacts': 1},
                                ss.filters.Filter(800., 100.))
    with pytest.warns(UserWarning):
        ss.extract.align_spikes(sp, {'data': np.array([10.0])}, [-0.05, 0.05])
    with pytest.raises(IndexError):
        ss.features.fetPCA({'data': np.zeros((5, 10, 2), np.float32)}, contacts=[4])
    out = ss.extract.detect_spikes(sp, thresh=5.0)            # nothing crosses
    assert out['data'].shape == (0,) and out['data'].dtype == np.float64
    al = ss.extract.align_spikes(sp, out, [-0.5, 0.5])
    assert al['data'].shape == (0,)
    w = ss.extract.extract_spikes(sp, al, [-0.5, 0.5])
    assert w['data'].shape[1:] == (0, 1)


def test_filter_object_call_and_device_signal(ss):
    rng = np.random.default_rng(2)
    x = rng.standard_normal(3000)
    from scipy import signal
    flt = ss.filters.Filter(800., 100.)
    y = flt(x, 25000)
    b, a = flt._design_filter(25000)
    assert rel_err(y, signal.filtfilt(b, a, x)) < 1e-11
    zp = ss.filters.ZeroPhaseFilter('ellip', [300., 3000.])
    b, a = zp._design_filter(25000)
    assert rel_err(zp(x, 25000), signal.filtfilt(b, a, x)) < 1e-7
    raw = {'data': rng.standard_normal((3, 5000)), 'FS': 25000, 'n_contacts': 3}
    f = ss.filters.filter_proxy(raw, flt)
    d = f['data']
    assert d.shape == (3, 5000) and d[1, :].shape == (5000,) and d[[0, 2], 10:20].shape == (2, 10)
    assert np.array_equal(d[:], np.asarray(d))
    assert f['FS'] == 25000 and raw['data'] is not f['data']


def test_install_rebinds_reference_namespace(ss):
    import types
    fake = types.SimpleNamespace(filters=types.SimpleNamespace(fltLinearIIR=1),
                                 extract=types.SimpleNamespace(detect_spikes=2),
                                 features=types.SimpleNamespace())
    orig = ss.install(fake)
    assert fake.extract.detect_spikes is ss.extract.detect_spikes
    assert fake.filters.fltLinearIIR is ss.filters.fltLinearIIR
    assert fake.features.fetPCA is ss.features.fetPCA
    assert orig[("extract", "detect_spikes")] == 2


def test_cheap_features_and_combine(ss, oracle, golden):
    """fetP2P / fetMin / combine(norm=True), bit-exact."""
    rng = np.random.default_rng(9)
    wv = {'data': golden['chain_waves'], 'time': np.arange(25.0), 'FS': 25000,
          'is_valid': np.arange(golden['chain_waves'].shape[1]) % 7 != 0}
    p2p = ss.features.fetP2P(wv)
    assert p2p['data'].dtype == np.float32 and np.array_equal(p2p['data'], golden['chain_p2p'])
    assert p2p['names'] == ['Ch0:P2P', 'Ch1:P2P'] and np.array_equal(p2p['is_valid'], wv['is_valid'])
    assert np.array_equal(ss.features.fetMin(wv)['data'], golden['chain_min'])
    big = {'data': rng.standard_normal((30, 20011, 3)).astype(np.float32)}
    assert np.array_equal(ss.features.fetP2P(big)['data'], oracle.fetP2P(big)['data'])
    assert np.array_equal(ss.features.fetMin(big, contacts=[2, 0])['data'],
                          oracle.fetMin(big, contacts=[2, 0])['data'])
    # normalize: bit-exact against NumPy on arbitrary float64 columns
    feats = {'data': rng.standard_normal((50021, 7)) * np.array([1, 1e3, 1e-3, 5, 7, 9, 11.0]) + 3,
             'names': list("abcdefg")}
    got = ss.features.normalize(feats)
    ref = oracle.normalize(feats)
    assert np.array_equal(got['data'], ref['data']) and feats['data'].max() > 1.5   # copy=True
    comb = ss.features.combine((ss.features.fetP2P(wv), oracle.fetPCA(wv, ncomps=2)), norm=True,
                               feat_method_names=['P2P', 'PCA'])
    refc = oracle.combine((oracle.fetP2P(wv), oracle.fetPCA(wv, ncomps=2)), norm=True,
                          feat_method_names=['P2P', 'PCA'])
    assert comb['names'] == refc['names'] and np.array_equal(comb['data'], refc['data'])
    assert np.array_equal(comb['is_valid'], refc['is_valid'])
    assert ss.features.fetSpIdx(wv)['data'].shape == (wv['data'].shape[1], 1)
    assert ss.features.fetSpTime({'data': np.array([1.0, 2.0])})['data'].shape == (2, 1)


def test_sort_frontend_equals_per_contact_calls(ss, golden):
    """The one-call front end returns, per contact, what the reference-style calls return."""
    import torch
    FS = 25000
    x = golden['chain_raw']
    raw = {'data': torch.from_numpy(x).pin_memory(), 'FS': FS, 'n_contacts': 2}
    flt = ss.filters.Filter(800., 100.)
    out = ss.sort_frontend(raw, filt=flt, sp_win=[-0.2, 0.8], edge='rising', align_type='max',
                           thresh='auto', ncomps=2, chunksize=25000)
    f = ss.filters.filter_proxy({'data': x, 'FS': FS, 'n_contacts': 2}, flt, chunksize=25000)
    for c in range(2):
        spt = ss.extract.detect_spikes(f, thresh='auto', edge='rising', contact=c)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            spa = ss.extract.align_spikes(f, spt, [-0.2, 0.8], type='max', contact=c)
        wv = ss.extract.extract_spikes(f, spa, [-0.2, 0.8], contacts=c)
        ft = ss.features.fetPCA(wv, ncomps=2)
        assert np.array_equal(out['spt'][c]['data'], spt['data'])
        assert out['spt'][c]['thresh'] == spt['thresh'] and out['spt'][c]['contact'] == c
        assert np.array_equal(out['spt_aligned'][c]['data'], spa['data'])
        assert np.array_equal(out['sp_waves'][c]['data'], wv['data'])
        assert np.array_equal(out['sp_waves'][c]['time'], wv['time'])
        assert ('is_valid' in out['sp_waves'][c]) == ('is_valid' in wv)
        assert out['features'][c]['names'] == ft['names']
        assert rel_err(sign_normalise(out['features'][c]['data'], ft['data']), ft['data']) < 1e-9
    assert np.array_equal(out['spt_aligned'][1]['data'], golden['chain_align'])


@pytest.mark.parametrize("pinned", [True, False])
def test_sort_frontend_slab_pipeline_equals_single_upload(ss, pinned):
    """The upload-overlapped path (time slabs = shards on one GPU over a LocalFabric) returns
    what the single-upload chain returns: spike times and waveforms bit for bit, PCA scores
    to round-off (the Gram is summed slab by slab)."""
    import torch
    from spikesort_b200 import synth, pipeline
    FS, n_c, n, cs = 3000, 5, 170500, 10000          # 17 full chunks + a short one
    x = synth.make_recording(n_c, n, FS, seed=77, rate_hz=30.0, noise=6.0, amp=(70.0, 160.0),
                             burst=True, negative=True)
    assert pipeline._slab_bounds(n, cs, 10 * FS, 6) is not None
    data = x.pin_memory() if pinned else x.numpy()
    raw = {'data': data, 'FS': FS, 'n_contacts': n_c}
    flt = ss.filters.Filter(300., 40.)
    kw = dict(filt=flt, sp_win=[-1.0, 2.0], edge='falling', align_type='min', thresh='4', ncomps=2,
              chunksize=cs)
    one = ss.sort_frontend(raw, slabs=1, **kw)
    for slabs in ('auto', 3):
        got = ss.sort_frontend(raw, slabs=slabs, **kw)
        tot = 0
        for c in range(n_c):
            assert np.array_equal(got['spt'][c]['data'], one['spt'][c]['data']), c
            assert got['spt'][c]['thresh'] == one['spt'][c]['thresh']
            assert np.array_equal(got['spt_aligned'][c]['data'], one['spt_aligned'][c]['data']), c
            assert np.array_equal(got['sp_waves'][c]['data'], one['sp_waves'][c]['data']), c
            assert ('is_valid' in got['sp_waves'][c]) == ('is_valid' in one['sp_waves'][c])
            a, b = got['features'][c]['data'], one['features'][c]['data']
            assert rel_err(sign_normalise(a, b), b) < 1e-9, c
            tot += len(one['spt_aligned'][c]['data'])
        assert tot > 300


def test_chain_with_resampled_alignment(ss, golden):
    """The batched chain with align_spikes(resample=r) (the tutorial aligns with resample=10):
    per contact identical to the per-contact drop-in calls, which are bit-exact against the
    reference (test_resample_golden); monolithic and slab-pipelined front ends agree."""
    FS = 25000
    x = golden['chain_raw']
    flt = ss.filters.Filter(800., 100.)
    f = ss.filters.filter_proxy({'data': x, 'FS': FS, 'n_contacts': 2}, flt, chunksize=25000)
    kw = dict(filt=flt, sp_win=[-0.2, 0.8], edge='rising', align_type='max', thresh='4', ncomps=2,
              chunksize=10000, resample=10)
    outs = [ss.sort_frontend({'data': x, 'FS': FS, 'n_contacts': 2}, slabs=s, **kw) for s in (1, 'auto')]
    f10 = ss.filters.filter_proxy({'data': x, 'FS': FS, 'n_contacts': 2}, flt, chunksize=10000)
    for c in range(2):
        spt = ss.extract.detect_spikes(f10, thresh='4', edge='rising', contact=c)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            spa = ss.extract.align_spikes(f10, spt, [-0.2, 0.8], type='max', contact=c, resample=10)
            plain = ss.extract.align_spikes(f10, spt, [-0.2, 0.8], type='max', contact=c)
        assert not np.array_equal(spa['data'], plain['data'])          # sub-sample shifts happened
        for out in outs:
            assert np.array_equal(out['spt_aligned'][c]['data'], spa['data']), c
        wv = ss.extract.extract_spikes(f10, spa, [-0.2, 0.8], contacts=c)
        assert np.array_equal(outs[0]['sp_waves'][c]['data'], wv['data'])
        assert np.array_equal(outs[1]['sp_waves'][c]['data'], wv['data'])


def test_sort_frontend_pageable_numpy_recording(cuda, monkeypatch):
    """A big pageable NumPy recording goes through the slab feeder's pinned double buffer
    (threaded copies one slab ahead of the DMA); the result is the one of the plain path."""
    import spikesort_b200 as ss
    from spikesort_b200 import filters, synth
    FS, C, N = 30000, 8, 4200000                              # 67 MB of int16: above the 64 MB bar
    arr = synth.make_recording(C, N, FS, seed=21, device=cuda, rate_hz=15.0, noise=10.0).cpu().numpy().copy()
    kw = dict(filt=filters.Filter(800., 100.), sp_win=[-0.2, 0.8], edge='falling', align_type='min',
              thresh='auto', ncomps=2, chunksize=1000000)
    monkeypatch.setenv("SS_PAGEABLE_STAGING", "0")
    plain = ss.sort_frontend({'data': arr, 'FS': FS, 'n_contacts': C}, **kw)
    monkeypatch.setenv("SS_PAGEABLE_STAGING", "1")
    monkeypatch.setenv("SS_STAGING_THREADS", "3")
    staged = ss.sort_frontend({'data': arr, 'FS': FS, 'n_contacts': C}, **kw)
    assert sum(len(plain['spt_aligned'][c]['data']) for c in range(C)) > 100
    for c in range(C):
        for key in ('spt', 'spt_aligned', 'sp_waves', 'features'):
            assert np.array_equal(plain[key][c]['data'], staged[key][c]['data']), (c, key)


@pytest.mark.parametrize("shape,dtype", [((9, 4000001), np.int16), ((25, 700001, 1), np.float32),
                                         ((3, 3000000), np.float64)])
def test_staged_upload_of_pageable_arrays(cuda, monkeypatch, shape, dtype):
    """_lib.upload: big pageable arrays go through two pinned staging buffers (threaded chunk
    copies under the DMA of the chunk before); byte-identical to torch's own copy, odd sizes and
    a chunk-sized tail included."""
    import torch
    from spikesort_b200 import _lib
    rng = np.random.default_rng(1)
    host = (rng.standard_normal(shape) * 1000).astype(dtype)
    assert host.nbytes >= 64 << 20
    monkeypatch.setenv("SS_STAGING_THREADS", "3")
    got = _lib.upload(torch.from_numpy(host))
    assert got.is_cuda and got.shape == torch.Size(shape)
    assert torch.equal(got.cpu(), torch.from_numpy(host))
    monkeypatch.setenv("SS_PAGEABLE_STAGING", "0")
    assert torch.equal(_lib.upload(torch.from_numpy(host)), got)
