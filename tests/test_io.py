"""Raw-recording ingest (spikesort_b200/io.py): the reference's own
BakerlabFilter raw-waveform tests (``/root/reference/tests/test_io.py:94-200``) restated, on
the host path (CPU tier) and on the streamed-to-HBM path (GPU tier)."""
import filecmp
import glob
import json
import os

import numpy as np
import pytest


@pytest.fixture
def bakerlab(tmp_path):
    """test_io.py:95-120: one contact, 100 int16 samples, FS 5 kHz."""
    rng = np.random.default_rng(12)
    descr = {"fspike": "{ses_id}{el_id}.sp", "cell": "{ses_id}{el_id}{cell_id}.spt",
             "dirname": str(tmp_path), "FS": 5.E3, "n_contacts": 1}
    conf = tmp_path / "test.conf"
    conf.write_text(json.dumps(descr))
    data = rng.integers(-1000, 1000, (100,))
    data.astype(np.int16).tofile(str(tmp_path / "32test011.sp"))
    return dict(conf=str(conf), dir=str(tmp_path), data=data, el_node='/Test/s32test01/el1',
                descr=descr)


def _reader(conf, **kw):
    from spikesort_b200.io import BakerlabFilter
    return BakerlabFilter(conf, **kw)


def test_read_sp_host(bakerlab):                                     # test_io.py:184-189
    sp = _reader(bakerlab['conf']).read_sp(bakerlab['el_node'], host=True)
    assert sp['FS'] == 5.E3 and sp['n_contacts'] == 1
    assert np.array_equal(sp['data'][0, :], bakerlab['data'])
    assert sp['data'].dtype == np.int16


def test_sp_shape_host(bakerlab):                                    # test_io.py:191-200
    d = dict(bakerlab['descr'], n_contacts=4)
    open(bakerlab['conf'], 'w').write(json.dumps(d))
    sp = _reader(bakerlab['conf']).read_sp(bakerlab['el_node'], host=True)
    assert sp['data'].shape == (4, len(bakerlab['data']))


def test_write_sp_and_multichan(bakerlab):                           # test_io.py:161-182
    flt = _reader(bakerlab['conf'])
    flt.write_sp({'data': bakerlab['data'][np.newaxis, :]}, '/Test/s32test01/el2')
    assert filecmp.cmp(os.path.join(bakerlab['dir'], "32test011.sp"),
                       os.path.join(bakerlab['dir'], "32test012.sp"), shallow=False)
    d = dict(bakerlab['descr'], n_contacts=4, fspike="test{contact_id}.sp")
    open(bakerlab['conf'], 'w').write(json.dumps(d))
    _reader(bakerlab['conf']).write_sp({'data': np.repeat(bakerlab['data'][np.newaxis, :], 4, 0)},
                                       bakerlab['el_node'])
    assert len(glob.glob(os.path.join(bakerlab['dir'], "test?.sp"))) == 4


def test_bad_dataset(bakerlab):
    with pytest.raises(ValueError):
        _reader(bakerlab['conf']).read_sp('not-a-dataset', host=True)


@pytest.mark.gpu
def test_read_sp_streams_to_device(cuda, tmp_path):
    """Four contact files, staged in chunks that do not divide the length: the recording in
    HBM equals the files; the chain runs on it without another upload; write_sp round-trips."""
    import spikesort_b200 as ss
    from spikesort_b200 import synth
    FS, n = 25000, 70001
    x = synth.make_recording(4, n, FS, seed=5, rate_hz=40.0, noise=12.0, amp=(120.0, 300.0),
                             negative=True).numpy()
    descr = {"fspike": "{subject}_{ses_id}_el{el_id}_c{contact_id}.sp", "dirname": str(tmp_path),
             "FS": FS, "n_contacts": 4}
    conf = tmp_path / "rec.inf"
    conf.write_text(json.dumps(descr))
    for c in range(4):
        x[c].tofile(str(tmp_path / ("Gollum_5gollum01_el3_c%d.sp" % (c + 1))))
    flt = _reader(str(conf), chunksize=16384)
    sp = flt.read_sp("/Gollum/s5gollum01/el3")
    assert isinstance(sp['data'], ss.DeviceSignal) and sp['data'].shape == (4, n)
    assert sp['data'].dtype == np.int16 and sp['n_contacts'] == 4 and sp['FS'] == FS
    assert np.array_equal(np.asarray(sp['data']), x)
    assert np.array_equal(flt.read_sp("/Gollum/s5gollum01/el3", host=True)['data'], x)
    hp = ss.filters.Filter(800., 100.)
    a = ss.sort_frontend(sp, filt=hp, thresh='4', chunksize=30000)
    b = ss.sort_frontend({'data': x, 'FS': FS, 'n_contacts': 4}, filt=hp, thresh='4', chunksize=30000)
    for c in range(4):
        assert np.array_equal(a['spt_aligned'][c]['data'], b['spt_aligned'][c]['data'])
        assert len(a['spt_aligned'][c]['data']) > 10
    # lazy source: the slabs are read from the files while earlier slabs are on the GPU
    lazy = flt.read_sp("/Gollum/s5gollum01/el3", lazy=True)
    assert lazy['data'].shape == (4, n) and np.array_equal(np.asarray(lazy['data']), x)
    from spikesort_b200 import pipeline
    assert pipeline._slab_bounds(n, 10007, 10 * FS, 6) is None      # 'auto'-style threshold needs 250000 samples
    assert len(pipeline._slab_bounds(n, 10007, 0, 6)) == 6          # numeric threshold: 6 slabs
    for kw in (dict(thresh=-150.0, chunksize=10007), dict(thresh='4', chunksize=30000)):
        c_l = ss.sort_frontend(lazy, filt=hp, **kw)
        c_h = ss.sort_frontend({'data': x, 'FS': FS, 'n_contacts': 4}, filt=hp, slabs=1, **kw)
        for c in range(4):
            assert np.array_equal(c_l['spt'][c]['data'], c_h['spt'][c]['data'])
            assert np.array_equal(c_l['spt_aligned'][c]['data'], c_h['spt_aligned'][c]['data'])
            assert np.array_equal(c_l['sp_waves'][c]['data'], c_h['sp_waves'][c]['data'])
    flt.write_sp(sp, "/Gollum/s5gollum01/el4")
    for c in range(4):
        assert filecmp.cmp(str(tmp_path / ("Gollum_5gollum01_el3_c%d.sp" % (c + 1))),
                           str(tmp_path / ("Gollum_5gollum01_el4_c%d.sp" % (c + 1))), shallow=False)


def test_pageable_source_copies_slabs_with_threads():
    """pipeline._PageableSource (an in-memory NumPy recording as a slab source of the feeder):
    any slab, any thread count, ragged last slab - the bytes of arr[:, a:b], nothing else touched."""
    from spikesort_b200 import pipeline
    rng = np.random.default_rng(0)
    arr = rng.integers(-30000, 30000, size=(7, 5003)).astype(np.int16)
    src = pipeline._PageableSource(arr)
    assert src.shape == (7, 5003) and src.dtype == np.int16
    default_threads = pipeline._PageableSource.THREADS
    for threads in (1, 3, 16):
        pipeline._PageableSource.THREADS = threads
        for a, b in [(0, 1000), (1000, 5003), (4999, 5003), (17, 18)]:
            out = np.full((7, 4100), 7, dtype=np.int16)
            src.read_slab(a, b, out[:, :b - a] if b - a < 4100 else out)
            assert np.array_equal(out[:, :b - a], arr[:, a:b])
            assert (out[:, b - a:] == 7).all()
    pipeline._PageableSource.THREADS = default_threads
    single = pipeline._PageableSource(arr[:1])
    out = np.zeros((1, 10), np.int16)
    single.read_slab(5, 15, out)
    assert np.array_equal(out, arr[:1, 5:15])
