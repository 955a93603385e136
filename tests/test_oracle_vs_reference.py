"""CPU tier, build container only: re-pins the oracle against the UNMODIFIED reference
files executed through oracle/ref_shim.py on fresh seeded inputs (skipped on machines
without /root/reference, e.g. the GPU box, where the committed golden vectors stand in)."""
import warnings

import numpy as np
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.filters(), ref_shim.extract(), ref_shim.features()


@pytest.mark.parametrize("seed,dtype,FS", [(1, np.int16, 25000), (2, np.float64, 30000),
                                           (3, np.float32, 25000)])
def test_chain_matches_reference(oracle, ref, seed, dtype, FS):
    rf, re_, rfe = ref
    from spikesort_b200 import synth
    x = synth.make_recording(3, 80000, FS, seed=seed, rate_hz=50.0, noise=9.0, burst=(seed == 2),
                             dtype={np.int16: "int16", np.float32: "float32",
                                    np.float64: "float64"}[dtype]).numpy()
    raw = {'data': x, 'FS': FS, 'n_contacts': 3}
    f_ref = rf.filter_proxy(raw, rf.Filter(800., 100.), chunksize=30000)
    f_orc = oracle.filter_proxy(raw, oracle.Filter(800., 100.), chunksize=30000)
    assert np.array_equal(np.asarray(f_ref['data']), f_orc['data'])
    fd = {'data': np.asarray(f_ref['data']), 'FS': FS, 'n_contacts': 3}
    for c, edge, typ in ((0, 'falling', 'min'), (2, 'rising', 'max')):
        d0 = re_.detect_spikes(fd, contact=c, thresh='4', edge=edge)
        d1 = oracle.detect_spikes(fd, contact=c, thresh='4', edge=edge)
        assert d0['thresh'] == d1['thresh'] and np.array_equal(d0['data'], d1['data'])
        assert len(d0['data']) > 20
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            a0 = re_.align_spikes(fd, d0, [-0.6, 0.8], type=typ, contact=c)
            a1 = oracle.align_spikes(fd, d1, [-0.6, 0.8], type=typ, contact=c)
        assert np.array_equal(a0['data'], a1['data'])
        for contacts in ('all', c, [2, 1]):
            w0 = re_.extract_spikes(fd, a0, [-0.6, 0.8], contacts=contacts)
            w1 = oracle.extract_spikes(fd, a1, [-0.6, 0.8], contacts=contacts)
            assert np.array_equal(w0['data'], w1['data']) and np.array_equal(w0['time'], w1['time'])
            assert ('is_valid' in w0) == ('is_valid' in w1)
        p0, p1 = rfe.fetPCA(w0, ncomps=2), oracle.fetPCA(w1, ncomps=2)
        assert p0['names'] == p1['names'] and np.array_equal(p0['data'], p1['data'])


def test_bandpass_and_ellip_match_reference(oracle, ref):
    rf = ref[0]
    rng = np.random.default_rng(0)
    raw = {'data': rng.standard_normal((2, 9000)) * 40, 'FS': 30000, 'n_contacts': 2}
    for args in ((np.array([300., 6000.]), np.array([100., 8000.]), 1, 10, 'butter'),
                 (600., 450., 1, 20, 'ellip')):
        f0 = rf.filter_proxy(raw, rf.Filter(*args), chunksize=4000)
        f1 = oracle.filter_proxy(raw, oracle.Filter(*args), chunksize=4000)
        assert np.array_equal(np.asarray(f0['data']), f1['data'])
