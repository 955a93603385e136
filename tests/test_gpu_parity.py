"""GPU tier: CUDA kernels (through the C ABI) against the CPU oracle on the same seeded
inputs, at sizes the oracle finishes in seconds; edge cases; and size-independent
properties at the BASELINE.json configuration size.

Bars (north_star): spike times / indices / waveforms bit-exact given the same input
signal; filtered signal and PCA scores within rel 1e-5 (sign-normalised components).
"""
import warnings

import numpy as np
import pytest
from scipy import signal

from util import sign_normalise, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(cuda):
    import torch
    import spikesort_b200
    from spikesort_b200 import _ops, _lib, pipeline, synth
    _lib.load()
    return dict(torch=torch, ss=spikesort_b200, ops=_ops, pipeline=pipeline, synth=synth,
                dev=cuda)


def _dev(env, arr):
    return env['torch'].from_numpy(np.ascontiguousarray(arr)).to(env['dev'])


DESIGNS = {
    'hp1_butter': lambda: signal.iirdesign(800 * 2 / 25000, 100 * 2 / 25000, 1, 10, ftype='butter'),
    'bp2': lambda: signal.butter(1, [0.02, 0.4], 'bandpass'),
    'lp3': lambda: signal.butter(3, 0.2),
    'ellip_hp3': lambda: signal.iirdesign(0.02, 0.016, 1, 10, ftype='ellip'),
    'bp4': lambda: signal.butter(2, [300 / 15000, 6000 / 15000], 'bandpass'),
    'bp8': lambda: signal.iirdesign(np.array([300, 6000]) * 2 / 30000,
                                    np.array([100, 8000]) * 2 / 30000, 1, 10, ftype='butter'),
    'ellip10': lambda: signal.iirdesign(np.array([300, 6000]) * 2 / 30000,
                                        np.array([150, 7000]) * 2 / 30000, 1, 40, ftype='ellip'),
    'butter12': lambda: signal.butter(6, [300 / 15000, 6000 / 15000], 'bandpass'),
}
# measured agreement is far inside the 1e-5 bar; these bounds catch regressions.  The
# high orders are limited by the REFERENCE's (b, a)-form round-off, not by the scan.
TIGHT = {'hp1_butter': 1e-12, 'bp2': 1e-12, 'lp3': 1e-12, 'ellip_hp3': 1e-11, 'bp4': 1e-10,
         'bp8': 1e-8, 'ellip10': 1e-6, 'butter12': 1e-6}


@pytest.mark.parametrize("name", sorted(DESIGNS))
@pytest.mark.parametrize("dtype", ["int16", "float32", "float64"])
def test_filtfilt_vs_oracle(env, oracle, name, dtype):
    from spikesort_b200._design import tf2sos_exact
    b, a = DESIGNS[name]()
    sos, padlen = tf2sos_exact(b, a)
    rng = np.random.default_rng(11)
    # 3 contacts, chunks of 7000 (several tiles + ragged), last chunk short (1234)
    n, cs = 3 * 7000 + 1234, 7000
    x = rng.standard_normal((3, n)) * 500 + 300
    x = np.round(x).astype(np.int16) if dtype == "int16" else x.astype(dtype)
    ref = np.empty((3, n))
    for c in range(3):
        for lo in range(0, n, cs):
            ref[c, lo:lo + cs] = oracle.filtfilt(b, a, x[c, lo:lo + cs])
    got = env['ops'].filtfilt_sos(_dev(env, x), sos, padlen, cs).cpu().numpy()
    err = rel_err(got, ref)
    print("%s %s rel err %.3g" % (name, dtype, err))
    assert err < 1e-5                      # the north_star bar
    assert err < TIGHT[name]


def test_filtfilt_chunk_shapes_and_int16_wrap(env, oracle):
    from spikesort_b200._design import tf2sos_exact
    b, a = DESIGNS['hp1_butter']()
    sos, padlen = tf2sos_exact(b, a)
    rng = np.random.default_rng(5)
    for n, cs in [(7, 7), (500, 500), (512, 512), (1000, 100), (1024 - 12, 1024 - 12),
                  (4096, 1000), (30000, 1000000), (2999, 1000)]:
        x = np.round(rng.standard_normal((2, n)) * 9000).astype(np.int16)
        x[:, 0] = 30000                      # 2*x[0] - x[k] wraps around in int16
        x[:, -1] = -30000
        ref = np.empty((2, n))
        for c in range(2):
            for lo in range(0, n, cs):
                ref[c, lo:lo + cs] = oracle.filtfilt(b, a, x[c, lo:lo + cs])
        got = env['ops'].filtfilt_sos(_dev(env, x), sos, padlen, cs).cpu().numpy()
        assert rel_err(got, ref) < 1e-12, (n, cs)
    with pytest.raises(ValueError):
        env['ops'].filtfilt_sos(_dev(env, np.zeros((1, 1006), np.int16)), sos, padlen, 1000)
    # strided rows (ld > n_samples)
    big = np.round(rng.standard_normal((4, 6000)) * 100).astype(np.int16)
    view = _dev(env, big)[:, 500:5500]
    got = env['ops'].filtfilt_sos(view, sos, padlen, 2000).cpu().numpy()
    ref = np.stack([np.concatenate([oracle.filtfilt(b, a, big[c, 500:5500][lo:lo + 2000])
                                    for lo in range(0, 5000, 2000)]) for c in range(4)])
    assert rel_err(got, ref) < 1e-12


def test_auto_threshold(env, oracle):
    rng = np.random.default_rng(8)
    for dtype in (np.int16, np.float32, np.float64):
        x = (rng.standard_normal((5, 40000)) * 37 + 11).astype(dtype)
        for n0 in (40000, 25000, 10 * 30000):
            ref = np.array([8.0 * np.sqrt(x[r, :n0].var(dtype=np.float64)) for r in range(5)])
            got = env['ops'].var_threshold(_dev(env, x), n0, -8.0).cpu().numpy()
            assert np.all(np.abs(got + ref) <= 1e-12 * ref)


@pytest.mark.parametrize("dtype", ["int16", "float32", "float64"])
def test_detect_bit_exact(env, oracle, dtype):
    rng = np.random.default_rng(21)
    n = 3 * 8192 + 77
    x = rng.standard_normal((4, n)) * 50
    x = np.round(x).astype(np.int16) if dtype == "int16" else x.astype(dtype)
    x[2, :] = 0                                         # a silent contact: no spikes
    for FS in (25000, 30000.0):
        for edge, thr in (("rising", 60.0), ("falling", -60.0), ("max", 0.0), ("min", 49.5)):
            thr_dev = _dev(env, np.full(4, thr))
            spt, off = env['ops'].detect(_dev(env, x), thr_dev, edge in ("falling", "min"), FS)
            spt, off = spt.cpu().numpy(), off.cpu().numpy()
            assert off[0] == 0 and off[-1] == len(spt)
            for c in range(4):
                ref = oracle.detect_spikes({'data': x, 'FS': FS, 'n_contacts': 4},
                                           thresh=np.float64(thr), edge=edge, contact=c)['data']
                assert np.array_equal(spt[off[c]:off[c + 1]], ref), (FS, edge, c)


def test_detect_python_float_threshold_on_float32(env, oracle):
    # NumPy compares a float32 array with a PYTHON float in float32 (weak promotion)
    x = np.array([[0.0, 0.1, 0.30000001192092896, 0.5, 0.2, 0.30000001192092896, 0.31]],
                 dtype=np.float32)
    sp = {'data': x, 'FS': 1000, 'n_contacts': 1}
    for thr in (0.3, np.float64(0.3), np.float32(0.3)):
        ref = oracle.detect_spikes(sp, thresh=thr)['data']
        got = env['ss'].extract.detect_spikes(sp, thresh=thr)['data']
        assert np.array_equal(got, ref), thr


def _burst_recording(env, n_c=3, n=120000, FS=30000, seed=4):
    x = env['synth'].make_recording(n_c, n, FS, seed=seed, rate_hz=200.0, noise=6.0,
                                    amp=(60.0, 150.0), burst=True, negative=True).numpy()
    return x, FS


@pytest.mark.parametrize("align_type", ["min", "max"])
def test_align_extract_bit_exact_burst(env, oracle, align_type):
    x, FS = _burst_recording(env)
    raw = {'data': x.astype(np.float64) * 0.37, 'FS': FS, 'n_contacts': x.shape[0]}
    sp_win = [-0.6, 0.8]                                  # 42 samples: more than one warp pass
    for c in range(x.shape[0]):
        edge = 'falling' if align_type == 'min' else 'rising'
        thr = -30.0 if align_type == 'min' else 12.0
        det = oracle.detect_spikes(raw, thresh=thr, edge=edge, contact=c)
        assert len(det['data']) > 200
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ref = oracle.align_spikes(raw, det, sp_win, type=align_type, contact=c)
            got = env['ss'].extract.align_spikes(raw, det, sp_win, type=align_type, contact=c)
            ref_keep = oracle.align_spikes(raw, det, sp_win, type=align_type, contact=c, remove=False)
            got_keep = env['ss'].extract.align_spikes(raw, det, sp_win, type=align_type, contact=c,
                                                      remove=False)
        assert np.array_equal(got_keep['data'], ref_keep['data'])
        assert np.array_equal(got['data'], ref['data'])
        assert len(ref['data']) < len(det['data'])        # doubles were actually removed
        wr = oracle.extract_spikes(raw, ref, sp_win)
        wg = env['ss'].extract.extract_spikes(raw, ref, sp_win)
        assert np.array_equal(wg['data'], wr['data']) and np.array_equal(wg['time'], wr['time'])
        assert ('is_valid' in wg) == ('is_valid' in wr)
        for contacts in (1, [2, 0], np.array([1])):
            wr = oracle.extract_spikes(raw, ref, sp_win, contacts=contacts)
            wg = env['ss'].extract.extract_spikes(raw, ref, sp_win, contacts=contacts)
            assert np.array_equal(wg['data'], wr['data'])


@pytest.mark.parametrize("align_type,resample", [("min", 10), ("max", 3)])
def test_align_resampled_burst(env, oracle, align_type, resample):
    """Alignment on the spline-upsampled window (extract.py:245-258, resample != 1) on a
    burst recording with edge re-examinations and many exact ties (quantised samples):
    aligned times and upsampled waveforms bit-exact."""
    x, FS = _burst_recording(env, n_c=2, n=60000)
    raw = {'data': x.astype(np.float64) * 0.37, 'FS': FS, 'n_contacts': x.shape[0]}
    sp_win = [-0.6, 0.8]
    total = mism = 0
    for c in range(x.shape[0]):
        edge = 'falling' if align_type == 'min' else 'rising'
        thr = -30.0 if align_type == 'min' else 12.0
        det = oracle.detect_spikes(raw, thresh=thr, edge=edge, contact=c)
        assert len(det['data']) > 100
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ref = oracle.align_spikes(raw, det, sp_win, type=align_type, contact=c,
                                      resample=resample, remove=False)
            got = env['ss'].extract.align_spikes(raw, det, sp_win, type=align_type, contact=c,
                                                 resample=resample, remove=False)
            plain = oracle.align_spikes(raw, det, sp_win, type=align_type, contact=c, remove=False)
        assert got['data'].shape == ref['data'].shape
        total += len(ref['data'])
        mism += int(np.sum(got['data'] != ref['data']))
        assert not np.array_equal(ref['data'], plain['data'])     # sub-sample shifts happened
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            wr = oracle.extract_spikes(raw, ref, sp_win, resample=resample, contacts=[c])
            wg = env['ss'].extract.extract_spikes(raw, ref, sp_win, resample=resample, contacts=[c])
        assert np.array_equal(wg['data'], wr['data'])
        assert np.array_equal(wg['time'], wr['time'])
    print("resample=%d %s: %d of %d aligned times differ" % (resample, align_type, mism, total))
    assert mism == 0


def test_align_ties_and_plateaus(env, oracle):
    # flat tops: argmax must return the FIRST maximum, also after float32 rounding
    FS = 10000
    x = np.zeros((1, 4000))
    for k in range(20, 3900, 200):
        x[0, k:k + 7] = [1, 3, 5, 5, 5.0000000001, 3, 1]   # equal in float32
    x[0, 3000:3030] = 7.0                                   # long plateau: walks left edge by edge
    raw = {'data': x, 'FS': FS, 'n_contacts': 1}
    det = oracle.detect_spikes(raw, thresh=2.0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for sp_win in ([-0.5, 0.5], [-0.3, 1.0], [-0.2, 0.2]):
            ref = oracle.align_spikes(raw, det, sp_win, type='max', remove=False)
            got = env['ss'].extract.align_spikes(raw, det, sp_win, type='max', remove=False)
            assert np.array_equal(got['data'], ref['data']), sp_win


def test_extract_truncated_and_outside(env, oracle):
    rng = np.random.default_rng(3)
    FS = 25000
    x = np.round(rng.standard_normal((4, 5000)) * 100).astype(np.int16)
    raw = {'data': x, 'FS': FS, 'n_contacts': 4}
    tmax = 5000 * 1000.0 / FS
    spt = {'data': np.array([-3.0, -0.1, 0.0, 0.04, 0.19, 0.2, 0.21, 100.0, tmax - 0.81, tmax - 0.8,
                             tmax - 0.79, tmax - 0.01, tmax, tmax + 0.1, tmax + 50.0])}
    for sp_win in ([-0.2, 0.8], [-0.6, 0.8], [0, 0.4]):
        wr = oracle.extract_spikes(raw, spt, sp_win)
        wg = env['ss'].extract.extract_spikes(raw, spt, sp_win)
        assert np.array_equal(wg['data'], wr['data']), sp_win
        assert np.array_equal(wg['is_valid'], wr['is_valid'])


@pytest.mark.parametrize("n_spk,n_c,n_win", [(500, 4, 25), (3000, 1, 30), (77, 2, 42), (5000, 3, 112)])
def test_pca_vs_oracle(env, oracle, n_spk, n_c, n_win):
    w = env['synth'].make_waveforms(n_win, n_spk, n_c, seed=5).numpy()
    w += 3.0                                              # large mean: cancellation stress
    sp = {'data': w, 'time': np.arange(n_win, dtype=float), 'FS': 25000}
    ref = oracle.fetPCA(sp, ncomps=3)
    got = env['ss'].features.fetPCA(sp, ncomps=3)
    assert got['names'] == ref['names'] and got['data'].shape == ref['data'].shape
    err = rel_err(sign_normalise(got['data'], ref['data']), ref['data'])
    print("pca rel err %.3g" % err)
    assert err < 1e-5 and err < 1e-9
    ev_r, evec_r, sc_r = oracle.PCA(w[:, :, 0], 2)
    ev_g, evec_g, sc_g = env['ss'].features.PCA(w[:, :, 0], 2)
    k = min(n_win, 8)
    assert np.allclose(ev_g[:k], ev_r[:k], rtol=1e-9)
    assert rel_err(sign_normalise(evec_g[:, :3], evec_r[:, :3]), evec_r[:, :3]) < 1e-7
    assert np.all(np.diff(ev_g) <= 1e-12 * ev_g[0])       # sorted descending


def test_chain_vs_oracle_per_contact(env, oracle):
    """The batched device-resident chain equals the reference-style per-contact loop."""
    FS = 30000
    x = env['synth'].make_recording(6, 150000, FS, seed=2, rate_hz=40.0, noise=8.0,
                                    amp=(70.0, 180.0), negative=True).numpy()
    raw = {'data': x, 'FS': FS, 'n_contacts': 6}
    sp_win = [-0.2, 0.8]
    flt_g = env['ss'].filters.Filter(800., 100.)
    res = env['pipeline'].run_chain(_dev(env, x), FS, filt=flt_g, sp_win=sp_win, edge='falling',
                                    align_type='min', ncomps=2, chunksize=50000)
    filtered, ref = oracle.chain_per_contact(raw, oracle.Filter(800., 100.), sp_win,
                                             edge='falling', align_type='min', ncomps=2,
                                             chunksize=50000)
    assert rel_err(res.filtered.cpu().numpy(), filtered['data']) < 1e-11
    total = mism = 0
    for c in range(6):
        spt_d, spt_a, waves, feats = ref[c]
        got = res.per_contact(c)
        assert abs(got['thresh'] - spt_d['thresh']) <= 1e-12 * abs(spt_d['thresh'])
        total += len(spt_a['data'])
        mism += len(np.setxor1d(got['spt'], spt_a['data']))
    print("end-to-end spike-time mismatches (GPU filter + auto threshold): %d of %d" % (mism, total))
    assert total > 100
    assert mism <= max(1, total // 1000)        # a sample within 1 ulp of the threshold may flip
    # bit-exact GIVEN the reference's filtered signal and thresholds
    thr = _dev(env, np.array([ref[c][0]['thresh'] for c in range(6)]))
    res2 = env['pipeline'].run_chain(_dev(env, filtered['data']), FS, filt=None, sp_win=sp_win,
                                     edge='falling', align_type='min', ncomps=2, thr_override=thr)
    for c in range(6):
        spt_d, spt_a, waves, feats = ref[c]
        got = res2.per_contact(c)
        assert np.array_equal(got['spt_detected'], spt_d['data'])
        assert np.array_equal(got['spt'], spt_a['data'])
        assert np.array_equal(got['waves'], waves['data'])
        assert np.array_equal(got['is_valid'].all(), 'is_valid' not in waves)
        err = rel_err(sign_normalise(got['scores'], feats['data']), feats['data'])
        assert err < 1e-9, (c, err)


@pytest.mark.parametrize("FS,n,cs", [(1000, 50000, 3000), (1000, 50000, 50000), (30000, 200000, 60000),
                                     (500, 9000, 2000), (1000, 8192 * 3, 8192), (1000, 20001, 512)])
def test_fused_filter_detect_equals_separate_passes(env, FS, n, cs):
    """ss_filtfilt_detect_sos == ss_filtfilt_sos + ss_var_threshold + ss_detect_mark/emit,
    bit for bit (output, thresholds, spike times), for auto and given thresholds."""
    ops, torch = env['ops'], env['torch']
    x = env['synth'].make_recording(5, n, FS, seed=9, rate_hz=max(40.0, FS / 100.0), noise=7.0,
                                    amp=(60.0, 160.0)).to(env['dev'])
    x[3] = 0                                                # silent contact
    flt = env['ss'].filters.Filter(FS * 0.03, FS * 0.004)
    sos, padlen = flt.sections(FS)
    for falling in (True, False):
        y = ops.filtfilt_sos(x, sos, padlen, cs)
        thr = ops.var_threshold(y, int(10 * FS), -1.5 if falling else 1.5)
        spt, off = ops.detect(y, thr, falling, FS)
        assert spt.numel() > 3
        y2, thr2, spt2, off2 = ops.filtfilt_detect(x, sos, padlen, cs, FS, falling, thr=None,
                                                   n0=int(10 * FS), factor=-1.5 if falling else 1.5)
        assert torch.equal(y, y2) and torch.equal(thr, thr2)
        assert torch.equal(off, off2) and torch.equal(spt, spt2)
        y3, thr3, spt3, off3 = ops.filtfilt_detect(x, sos, padlen, cs, FS, falling, thr=thr)
        assert torch.equal(y, y3) and torch.equal(off, off3) and torch.equal(spt, spt3)
        # a threshold inside the noise: crossings everywhere, incl. tile and chunk edges
        low = torch.full((5,), -2.0 if falling else 2.0, dtype=torch.float64, device=env['dev'])
        spt4, off4 = ops.detect(y, low, falling, FS)
        assert spt4.numel() > n // 10
        y5, thr5, spt5, off5 = ops.filtfilt_detect(x, sos, padlen, cs, FS, falling, thr=low)
        assert torch.equal(off4, off5) and torch.equal(spt4, spt5)
    # the chain gives the same result either way
    win = (-5000.0 / FS, 10000.0 / FS)
    a = env['pipeline'].run_chain(x, FS, filt=flt, sp_win=win, chunksize=cs, thresh='1.5', fused=True)
    b = env['pipeline'].run_chain(x, FS, filt=flt, sp_win=win, chunksize=cs, thresh='1.5', fused=False)
    assert torch.equal(a.spt, b.spt) and torch.equal(a.waves, b.waves)
    assert torch.equal(a.offsets, b.offsets) and torch.equal(a.thresh, b.thresh)


def test_full_size_properties(env):
    """BASELINE.json configs[1] (32 contacts x 18 000 000 samples @ 30 kHz, int16):
    size-independent properties of the chain."""
    torch, ops = env['torch'], env['ops']
    FS, n_c, n = 30000, 32, 18000000
    x = env['synth'].make_recording(n_c, n, FS, seed=2, device=env['dev'], rate_hz=10.0,
                                    noise=10.0)
    flt = env['ss'].filters.Filter(800., 100.)
    sos, padlen = flt.sections(FS)
    y = ops.filtfilt_sos(x, sos, padlen, 1000000)
    assert bool(torch.isfinite(y).all())
    # (1) chunk independence: filtering one (contact, chunk) alone gives the same values
    for c, j in ((0, 0), (17, 9), (31, 17)):
        alone = ops.filtfilt_sos(x[c:c + 1, j * 1000000:(j + 1) * 1000000], sos, padlen, 1000000)
        assert torch.equal(alone[0], y[c, j * 1000000:(j + 1) * 1000000])
    # (2) linearity: filter(2x) == 2 filter(x) exactly (power-of-two scaling is exact;
    #     int16 inputs here stay far from the wrap-around range)
    y2 = ops.filtfilt_sos((x[:2] * 2).contiguous(), sos, padlen, 1000000)
    assert torch.equal(y2, 2 * y[:2])
    # (3) a high-pass removes the mean: per-contact mean of the output ~ 0
    assert float(y.mean(dim=1).abs().max()) < 0.05
    # (4) detection: sorted per contact, all within the recording, crossings verified
    thr = ops.var_threshold(y, 10 * FS, -8.0)
    spt, off = ops.detect(y, thr, True, FS)
    offh = off.cpu().numpy()
    assert offh[-1] == spt.numel() and offh[-1] > 1000
    d = spt[1:] - spt[:-1]
    starts = torch.zeros(spt.numel() - 1, dtype=torch.bool, device=spt.device)
    starts[torch.from_numpy(offh[1:-1][(offh[1:-1] > 0) & (offh[1:-1] < spt.numel())] - 1).to(spt.device)] = True
    assert bool((d[~starts] > 0).all())
    idx = torch.round(spt / 1000.0 * FS).long()
    rows = torch.repeat_interleave(torch.arange(n_c, device=spt.device), (off[1:] - off[:-1]))
    a, b = y[rows, idx], y[rows, idx + 1]
    assert bool(((a > thr[rows]) & (b < thr[rows])).all())
    # checksum of checksums: the number of crossings equals a torch recount
    recount = ((y[:, :-1] > thr[:, None]) & (y[:, 1:] < thr[:, None])).sum()
    assert int(recount) == spt.numel()
    # (5) alignment never moves to a shallower sample and lands on the window minimum
    sp_win = [-0.2, 0.8]
    al = spt.clone()
    ops.align(y, al, off, None, FS, sp_win, True)
    i_det = torch.round(spt / 1000.0 * FS).long()
    # (the reference's truncating ms->index round trip reports ~12 % of the times one
    #  sample late, SURVEY.md 0.4, so the extremum is at the reported sample or the one before;
    #  the arg-min runs on float32-rounded values)
    i_al = torch.round(al / 1000.0 * FS).long().clamp(1, n - 1)
    deepest = torch.minimum(y[rows, i_al], y[rows, i_al - 1])
    assert bool((deepest <= y[rows, i_det] + 1e-6 * y[rows, i_det].abs()).all())
    kept, off2 = ops.remove_doubles(al, off, 1000.0 / FS)
    assert kept.numel() <= al.numel() and int(off2[-1]) == kept.numel()
    waves, valid, n_inv = ops.extract(y, kept, FS, sp_win, contacts=None, offsets=off2)
    win0, win1, _ = ops.window_params(sp_win, FS)
    inner = valid.bool()
    amin = waves[:, inner, 0].argmin(dim=0)
    assert bool(((amin + win0).abs() <= 1).float().mean() > 0.95)   # minimum at t = 0 (+-1 sample)
    # (6) PCA: whitened scores have unit variance per contact
    gram, ssum, _, counts = ops.pca_gram(waves, offsets=off2)
    evals, evecs = ops.pca_eig(gram, ssum, counts)
    scores = ops.pca_project(waves, evals, evecs, 2, offsets=off2)
    o2 = off2.cpu().numpy()
    for c in (0, 13, 31):
        s = scores[o2[c]:o2[c + 1]]
        assert abs(float(s[:, 0].var()) - 1.0) < 1e-6
        vt = evecs[c].t() @ evecs[c]
        assert float((vt - torch.eye(vt.shape[0], device=vt.device, dtype=vt.dtype)).abs().max()) < 1e-12


def test_burst_stress_384_channels(env, oracle):
    """BASELINE.json configs[3]: >= 200 Hz/channel bursts on 384 channels - compaction,
    alignment re-queue, duplicate removal and the Gram at high spike density.  The whole
    recording runs through the batched chain; a sample of contacts is checked against the
    oracle (bit-exact given the same filtered signal), the rest through invariants."""
    torch, ops = env['torch'], env['ops']
    FS, n_c, n = 30000, 384, 300000                       # 10 s
    x = env['synth'].make_recording(n_c, n, FS, seed=4, device=env['dev'], rate_hz=220.0,
                                    noise=6.0, amp=(60.0, 150.0), burst=True, negative=True)
    flt = env['ss'].filters.Filter(800., 100.)
    sp_win = [-0.2, 0.8]
    res = env['pipeline'].run_chain(x, FS, filt=flt, sp_win=sp_win, edge='falling',
                                    align_type='min', thresh='4', ncomps=2, chunksize=100000)
    off = res.offsets.cpu().numpy()
    offd = res.offsets_detected.cpu().numpy()
    counts = np.diff(off)
    assert counts.min() > 1000 and off[-1] == res.spt.numel() == res.waves.shape[1]
    assert (np.diff(offd) >= counts).all() and (np.diff(offd) > counts).any()   # doubles removed
    assert res.scores.shape == (off[-1], 2) and bool(torch.isfinite(res.scores).all())
    filtered = res.filtered
    for c in (0, 191, 383):
        raw = {'data': filtered[c:c + 1].cpu().numpy(), 'FS': FS, 'n_contacts': 1}
        thr = float(res.thresh[c])
        det = oracle.detect_spikes(raw, thresh=np.float64(thr), edge='falling', contact=0)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            al = oracle.align_spikes(raw, det, sp_win, type='min', contact=0)
        wv = oracle.extract_spikes(raw, al, sp_win, contacts=0)
        got = res.per_contact(c)
        assert np.array_equal(got['spt_detected'], det['data'])
        assert np.array_equal(got['spt'], al['data'])
        assert np.array_equal(got['waves'], wv['data'])
        ft = oracle.fetPCA(wv, ncomps=2)
        assert rel_err(sign_normalise(got['scores'], ft['data']), ft['data']) < 1e-9


def test_fetpca_ten_million_waveforms_properties(env):
    """BASELINE.json configs[4]: fetPCA only on 10 M pre-extracted 4-contact waveforms
    (25 x 10 000 000 x 4 float32 = 4 GB).  Size-independent properties: whitened scores
    have unit variance and are uncorrelated; eigenvectors orthonormal; the Gram of a
    subsample agrees with torch's float64 covariance."""
    torch, ops = env['torch'], env['ops']
    n_win, n_spk, n_c = 25, 10000000, 4
    w = env['synth'].make_waveforms(n_win, n_spk, n_c, seed=5, device=env['dev'])
    gram, ssum, ref, counts = ops.pca_gram(w)
    evals, evecs = ops.pca_eig(gram, ssum, counts)
    scores = ops.pca_project(w, evals, evecs, 2)
    assert scores.shape == (n_spk, 8)
    cov = torch.cov(scores[:2000000].t())
    for c in range(n_c):                                   # PCs of ONE contact are uncorrelated
        blk = cov[2 * c:2 * c + 2, 2 * c:2 * c + 2]
        assert float((blk - torch.eye(2, device=cov.device, dtype=cov.dtype)).abs().max()) < 0.01
    v = scores.var(dim=0)
    assert float((v - 1).abs().max()) < 2e-5
    for c in range(n_c):
        vt = evecs[c].t() @ evecs[c]
        assert float((vt - torch.eye(n_win, device=vt.device, dtype=vt.dtype)).abs().max()) < 1e-12
        full = torch.cov(w[:, :, c].to(torch.float64))
        mine = (gram[c] - torch.outer(ssum[c], ssum[c]) / n_spk) / (n_spk - 1)
        # 10 M spikes take the tensor-core Gram (3xTF32, truncating float32 accumulation)
        assert float((mine - full).abs().max()) < 5e-6 * float(full.abs().max())


def test_speculative_chain_equals_exact_chain(env):
    """run_chain sized by the previous call's spike count (+25 %), one host sync at the end,
    returns what the exact-size chain returns - bit for bit - and falls back to exact sizes
    when the capacity is too small."""
    torch, pipeline = env['torch'], env['pipeline']
    FS = 30000
    x = env['synth'].make_recording(6, 400000, FS, seed=21, rate_hz=60.0, noise=8.0,
                                    amp=(80.0, 200.0), negative=True, burst=True).to(env['dev'])
    flt = env['ss'].filters.Filter(800., 100.)
    kw = dict(filt=flt, sp_win=[-0.2, 0.8], edge='falling', align_type='min', thresh='auto',
              ncomps=2, chunksize=100000)
    pipeline._CAPACITY_HINTS.clear()
    exact = pipeline.run_chain(x, FS, speculate=False, **kw)
    assert not pipeline._CAPACITY_HINTS
    first = pipeline.run_chain(x, FS, **kw)                   # learns the count (exact sizes)
    assert len(pipeline._CAPACITY_HINTS) == 1
    spec = pipeline.run_chain(x, FS, **kw)                    # capacity-sized
    key = next(iter(pipeline._CAPACITY_HINTS))
    pipeline._CAPACITY_HINTS[key] = 10                        # far too small: must fall back
    small = pipeline.run_chain(x, FS, **kw)
    assert pipeline._CAPACITY_HINTS[key] == exact.spt_detected.numel()
    assert exact.spt.numel() > 500
    for r in (first, spec, small):
        for name in ('thresh', 'spt_detected', 'offsets_detected', 'spt', 'offsets', 'valid',
                     'evals', 'evecs', 'scores'):
            assert torch.equal(getattr(r, name), getattr(exact, name)), name
        assert torch.equal(r.waves, exact.waves)
        assert int(r.n_invalid) == int(exact.n_invalid)
    for c in (0, 5):
        a, b = spec.per_contact(c), exact.per_contact(c)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    # lazy results: two chains queued back to back, finalised afterwards (in any order), and a
    # lazy result whose capacity was too small
    l1 = pipeline.run_chain(x, FS, lazy=True, **kw)
    l2 = pipeline.run_chain(x, FS, lazy=True, **kw)
    assert l1._pending is not None and l2._pending is not None
    assert l1.spt.numel() > exact.spt.numel()                 # still capacity-sized
    l2.finalize()
    l1.finalize()
    pipeline._CAPACITY_HINTS[key] = 10
    l3 = pipeline.run_chain(x, FS, lazy=True, **kw)
    assert l3.finalize() is l3 and l3.finalize() is l3        # idempotent
    assert pipeline._CAPACITY_HINTS[key] == exact.spt_detected.numel()
    for r in (l1, l2, l3):
        for name in ('thresh', 'spt_detected', 'offsets_detected', 'spt', 'offsets', 'valid',
                     'evals', 'evecs', 'scores', 'waves'):
            assert torch.equal(getattr(r, name), getattr(exact, name)), name
    # unfused path too
    pipeline._CAPACITY_HINTS.clear()
    u1 = pipeline.run_chain(x, FS, fused=False, **kw)
    u2 = pipeline.run_chain(x, FS, fused=False, **kw)
    assert torch.equal(u1.spt, u2.spt) and torch.equal(u1.waves, u2.waves)
    assert torch.equal(u1.spt, exact.spt)


@pytest.mark.parametrize("seed", range(12))
def test_random_small_cases_bit_exact(env, oracle, seed):
    """Randomised sweep of shapes, dtypes, rates, windows and edges at sizes the oracle does
    in milliseconds: detect -> align -> extract per contact must be bit-identical given the
    same signal (includes lengths that are no multiple of 32 / 8192, windows wider than a
    warp, spikes at both ends of the recording, thresholds nothing crosses)."""
    rng = np.random.default_rng(1000 + seed)
    FS = int(rng.choice([1000, 8000, 25000, 30000, 44100]))
    n_c = int(rng.integers(1, 5))
    n = int(rng.integers(40, 30000))
    dtype = [np.int16, np.float32, np.float64][seed % 3]
    x = rng.standard_normal((n_c, n)) * 40.0
    for c in range(n_c):                                   # spikes, some at the very edges
        for pos in rng.integers(0, n, size=max(n // 400, 2)):
            w = int(rng.integers(2, 12))
            a, b = max(pos - w, 0), min(pos + w, n)
            x[c, a:b] -= rng.uniform(100, 400) * np.hanning(b - a + 2)[1:-1]
    x = np.round(x).astype(dtype) if dtype is np.int16 else x.astype(dtype)
    raw = {'data': x, 'FS': FS, 'n_contacts': n_c}
    lo = float(rng.choice([-0.2, -0.6, -1.0, -2.5]))
    hi = float(rng.choice([0.4, 0.8, 2.0]))
    sp_win = [lo, hi]
    edge, typ = [('falling', 'min'), ('rising', 'max'), ('min', 'min'), ('max', 'max')][seed % 4]
    thr = float(rng.choice([-90.0, -150.0, -1e9])) if edge in ('falling', 'min') else \
        float(rng.choice([60.0, 90.0, 1e9]))
    ex = env['ss'].extract
    total = 0
    for c in range(n_c):
        ref = oracle.detect_spikes(raw, thresh=thr, edge=edge, contact=c)
        got = ex.detect_spikes(raw, thresh=thr, edge=edge, contact=c)
        assert np.array_equal(got['data'], ref['data']), (seed, c)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ra = oracle.align_spikes(raw, ref, sp_win, type=typ, contact=c)
            ga = ex.align_spikes(raw, ref, sp_win, type=typ, contact=c)
        assert np.array_equal(ga['data'], ra['data']), (seed, c)
        rw = oracle.extract_spikes(raw, ra, sp_win)
        gw = ex.extract_spikes(raw, ra, sp_win)
        assert gw['data'].shape == rw['data'].shape and np.array_equal(gw['data'], rw['data'])
        assert ('is_valid' in gw) == ('is_valid' in rw)
        if 'is_valid' in rw:
            assert np.array_equal(gw['is_valid'], rw['is_valid'])
        total += len(ra['data'])
    # the batched chain on the same signal (no filter) agrees with the per-contact calls
    res = env['pipeline'].run_chain(_dev(env, x), FS, filt=None, sp_win=sp_win, edge=edge,
                                    align_type=typ, thresh=thr, ncomps=2, pca=False)
    for c in range(n_c):
        ref = oracle.detect_spikes(raw, thresh=thr, edge=edge, contact=c)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ra = oracle.align_spikes(raw, ref, sp_win, type=typ, contact=c)
        assert np.array_equal(res.per_contact(c)['spt'], ra['data']), (seed, c)


def test_chain_without_spikes(env):
    """Nothing crosses the threshold: every path returns empty, well-shaped results - exact
    sizes, capacity-sized (second call) and the host-in/host-out front ends."""
    torch, pipeline, ss = env['torch'], env['pipeline'], env['ss']
    FS, n_c, n = 30000, 3, 90000
    x = torch.zeros((n_c, n), dtype=torch.int16)
    x[:, ::7] = 1
    flt = ss.filters.Filter(800., 100.)
    kw = dict(filt=flt, sp_win=[-0.2, 0.8], edge='falling', align_type='min', thresh=-1e6,
              ncomps=2, chunksize=30000)
    pipeline._CAPACITY_HINTS.clear()
    for _ in range(2):
        res = pipeline.run_chain(x.to(env['dev']), FS, **kw)
        assert res.spt.numel() == 0 and res.spt_detected.numel() == 0
        assert tuple(res.waves.shape) == (30, 0, 1) and res.scores is None
        assert res.offsets.tolist() == [0] * (n_c + 1)
    for slabs in (1, 'auto'):
        out = ss.sort_frontend({'data': x.numpy(), 'FS': FS, 'n_contacts': n_c}, slabs=slabs, **kw)
        for c in range(n_c):
            assert out['spt'][c]['data'].shape == (0,) and out['spt_aligned'][c]['data'].shape == (0,)
            assert out['sp_waves'][c]['data'].shape == (30, 0, 1)
            assert out['features'][c]['data'] is None


def test_filtfilt_on_odd_offset_int16_view(env, oracle):
    """A recording that is a VIEW starting at an odd int16 sample of its parent (2-byte aligned
    address, e.g. a time slab cut at an odd chunk size): the paired 32-bit loads of the
    summary sweep must fall back to scalar loads; results as for a fresh array."""
    rng = np.random.default_rng(8)
    parent = np.round(rng.standard_normal((3, 40001)) * 200).astype(np.int16)
    b, a = signal.iirdesign(800 * 2 / 25000, 100 * 2 / 25000, 1, 10, ftype='butter')
    from spikesort_b200._design import tf2sos_exact
    sos, padlen = tf2sos_exact(b, a)
    dev_parent = _dev(env, parent)
    for start in (1, 3, 2):
        view = dev_parent[:, start:]
        assert view.data_ptr() % 4 == (2 if start % 2 else 0)
        got = env['ops'].filtfilt_sos(view, sos, padlen, 10007).cpu().numpy()
        fresh = env['ops'].filtfilt_sos(_dev(env, parent[:, start:]), sos, padlen, 10007).cpu().numpy()
        assert np.array_equal(got, fresh) or rel_err(got, fresh) < 1e-13     # other summation order
        ref = np.stack([np.concatenate([oracle.filtfilt(b, a, parent[c, start:][lo:lo + 10007])
                                        for lo in range(0, parent.shape[1] - start, 10007)])
                        for c in range(3)])
        assert rel_err(got, ref) < 1e-12


@pytest.mark.parametrize("seed", range(10))
def test_random_filter_geometry(env, oracle, seed):
    """Randomised chunk geometry for the filter: odd and even chunk sizes, short last chunks,
    chunks barely longer than the padding, row-strided views with odd offsets, all dtypes,
    orders 1-8 - against the oracle's per-(contact, chunk) filtfilt."""
    rng = np.random.default_rng(500 + seed)
    name = ['hp1_butter', 'bp2', 'lp3', 'ellip_hp3', 'bp4', 'bp8'][seed % 6]
    b, a = DESIGNS[name]()
    from spikesort_b200._design import tf2sos_exact
    sos, padlen = tf2sos_exact(b, a)
    dtype = [np.int16, np.float32, np.float64][seed % 3]
    n_c = int(rng.integers(1, 5))
    cs = int(rng.integers(padlen + 2, 6000))
    n_chunks = int(rng.integers(1, 6))
    last = int(rng.integers(padlen + 1, cs + 1))
    n = (n_chunks - 1) * cs + last
    off = int(rng.integers(0, 4))
    parent = rng.standard_normal((n_c, n + off + 5)) * 300.0
    parent = np.round(parent).astype(dtype) if dtype is np.int16 else parent.astype(dtype)
    view = _dev(env, parent)[:, off:off + n]                      # strided rows, odd/even start
    got = env['ops'].filtfilt_sos(view, sos, padlen, cs).cpu().numpy()
    x = parent[:, off:off + n]
    ref = np.stack([np.concatenate([oracle.filtfilt(b, a, x[c, lo:min(lo + cs, n)])
                                    for lo in range(0, n, cs)]) for c in range(n_c)])
    assert got.shape == ref.shape
    assert rel_err(got, ref) < TIGHT[name] * 10, (name, dtype, n_c, cs, n_chunks, last, off)
    # a chunk that is not longer than the padding raises what SciPy raises
    if n > padlen + 1:
        with pytest.raises(ValueError):
            env['ops'].filtfilt_sos(view, sos, padlen, padlen)


@pytest.mark.parametrize("name", ['lp3', 'bp4', 'bp8', 'butter12'])
def test_filtfilt_high_order_large_chunks(env, oracle, name):
    """Orders >= 3 on int16 take the tile-per-lane summary sweep (filt_summary_wide_kernel) and
    the section-by-section second sweep: chunks long enough for several warps and more than
    one CTA per unit (> 1024 tiles), a short last chunk, even and odd sample offsets of the
    view (tile starts on either side of a 32-bit word), against the oracle's filtfilt."""
    from spikesort_b200._design import tf2sos_exact
    b, a = DESIGNS[name]()
    sos, padlen = tf2sos_exact(b, a)
    rng = np.random.default_rng(21)
    cs, n = 600000, 600000 + 100001
    parent = np.round(rng.standard_normal((2, n + 3)) * 400 + 50).astype(np.int16)
    dev_parent = _dev(env, parent)
    for start in (0, 1):
        x = parent[:, start:start + n]
        got = env['ops'].filtfilt_sos(dev_parent[:, start:start + n], sos, padlen, cs).cpu().numpy()
        ref = np.stack([np.concatenate([oracle.filtfilt(b, a, x[c, lo:lo + cs])
                                        for lo in range(0, n, cs)]) for c in range(2)])
        err = rel_err(got, ref)
        print("%s start %d rel err %.3g" % (name, start, err))
        assert err < TIGHT[name], (name, start)


@pytest.mark.parametrize("n_win,k,n_c", [(30, 2, 4), (25, 1, 2), (42, 3, 3), (112, 8, 2), (3, 2, 2), (2, 2, 1),
                                         (30, 5, 33)])
def test_pca_eig_top_matches_full_eigensolve(env, n_win, k, n_c):
    """ss_pca_eig_top (tridiagonalisation + multisection + inverse iteration, what the chain and
    fetPCA use) against ss_pca_eig (Jacobi) and numpy.linalg.eigh on covariances of spike-like
    waveforms: the k leading eigenvalues, eigenvectors up to sign, zeros elsewhere; identical
    scores within round-off."""
    ops, torch = env['ops'], env['torch']
    rng = np.random.default_rng(100 + n_win + k)
    n_spk = 4000
    tt = np.linspace(0, 1, n_win)
    templ = np.stack([np.exp(-((tt - 0.3) / 0.08) ** 2) - 0.4 * np.exp(-((tt - 0.5) / 0.15) ** 2),
                      np.exp(-((tt - 0.35) / 0.05) ** 2), np.sin(6 * tt) * np.exp(-3 * tt)])
    waves = np.empty((n_win, n_spk, n_c), np.float32)
    for c in range(n_c):
        amp = rng.standard_normal((n_spk, 3)) * np.array([120.0, 60.0, 20.0]) * (1 + c % 3)
        waves[:, :, c] = (amp @ templ + rng.standard_normal((n_spk, n_win)) * 8.0 + 15.0).T
    w = _dev(env, waves)
    gram, ssum, _, counts = ops.pca_gram(w)
    ev_f, vec_f = ops.pca_eig(gram, ssum, counts)
    ev_t, vec_t = ops.pca_eig(gram, ssum, counts, top=k)
    ev_f, vec_f, ev_t, vec_t = (a.cpu().numpy() for a in (ev_f, vec_f, ev_t, vec_t))
    assert np.all(ev_t[:, k:] == 0) and np.all(vec_t[:, :, k:] == 0)
    for c in range(n_c):
        cov = np.cov(waves[:, :, c].astype(np.float64))
        wv, U = np.linalg.eigh(np.atleast_2d(cov))
        wv, U = wv[::-1], U[:, ::-1]
        assert np.allclose(ev_t[c, :k], wv[:k], rtol=1e-11, atol=1e-11 * wv[0])
        assert np.allclose(ev_t[c, :k], ev_f[c, :k], rtol=1e-11, atol=1e-11 * wv[0])
        for j in range(k):
            gap = min(abs(wv[j] - wv[i]) for i in range(n_win) if i != j) / wv[0] if n_win > 1 else 1.0
            tol = 1e-12 / max(gap, 1e-6)
            z = vec_t[c, :, j]
            assert abs(np.linalg.norm(z) - 1.0) < 1e-12
            assert min(np.linalg.norm(z - U[:, j]), np.linalg.norm(z + U[:, j])) < tol, (c, j, gap)
            assert min(np.linalg.norm(z - vec_f[c, :, j]), np.linalg.norm(z + vec_f[c, :, j])) < tol
    s_f = ops.pca_project(w, _dev(env, ev_f), _dev(env, vec_f), min(k, 2)).cpu().numpy()
    s_t = ops.pca_project(w, _dev(env, ev_t), _dev(env, vec_t), min(k, 2)).cpu().numpy()
    sgn = np.sign(np.sum(s_f * s_t, axis=0))
    assert np.abs(s_f - s_t * sgn).max() <= 1e-9 * np.abs(s_f).max()


def test_pca_eig_top_special_matrices(env):
    """Degenerate and rank-deficient covariances: the leading pairs are any orthonormal basis of
    the leading invariant subspace - checked through residuals and orthonormality."""
    ops, torch = env['ops'], env['torch']
    rng = np.random.default_rng(3)
    n = 30
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    mats = [Q @ np.diag([5.0, 5.0] + list(np.linspace(1, 0.01, n - 2))) @ Q.T,
            Q @ np.diag([5.0, 5.0 - 1e-9] + list(np.linspace(1, 0.01, n - 2))) @ Q.T,
            7.0 * np.outer(Q[:, 0], Q[:, 0]),
            np.diag(np.arange(n, 0, -1.0)),
            np.zeros((n, n)),
            Q @ np.diag(np.r_[3.0, np.full(n - 1, 1e-14)]) @ Q.T]
    A = np.stack([0.5 * (m + m.T) for m in mats])
    cnt = 11
    gram = _dev(env, A * (cnt - 1.0))                       # sum = 0: covariance = A exactly
    ssum = _dev(env, np.zeros((len(mats), n)))
    counts = _dev(env, np.full(len(mats), cnt, np.int64))
    for k in (1, 2, 4):
        ev, vec = (a.cpu().numpy() for a in ops.pca_eig(gram, ssum, counts, top=k))
        for g in range(len(mats)):
            wv = np.linalg.eigvalsh(A[g])[::-1]
            scale = max(wv[0], 1e-300)
            assert np.abs(ev[g, :k] - np.abs(wv[:k])).max() <= 1e-12 * scale + 1e-290, (g, k)
            Z = vec[g][:, :k]
            assert np.abs(Z.T @ Z - np.eye(k)).max() < 1e-10, (g, k)
            res = np.abs(A[g] @ Z - Z * wv[:k]).max()
            assert res <= 1e-10 * scale + 1e-290, (g, k, res)


@pytest.mark.parametrize("seed", range(10))
def test_random_filter_geometry_int16_high_order(env, oracle, seed):
    """The int16 path of orders >= 3 (tile-per-lane summary sweep with cp.async staging and
    funnel-shift realignment, section-by-section second sweep) on random chunk geometry: chunks
    barely longer than the padding (every tile touches the odd extension), chunks of a few
    tiles, short last chunks, row-strided views starting on odd and even samples."""
    rng = np.random.default_rng(900 + seed)
    name = ['lp3', 'bp4', 'bp8', 'ellip10', 'butter12'][seed % 5]
    b, a = DESIGNS[name]()
    from spikesort_b200._design import tf2sos_exact
    sos, padlen = tf2sos_exact(b, a)
    n_c = int(rng.integers(1, 4))
    cs = int(rng.integers(padlen + 2, 400 if seed % 2 else 5000))
    n_chunks = int(rng.integers(1, 5))
    last = int(rng.integers(padlen + 1, cs + 1))
    n = (n_chunks - 1) * cs + last
    off = int(rng.integers(0, 4))
    parent = np.round(rng.standard_normal((n_c, n + off + 3)) * 250.0 + 20.0).astype(np.int16)
    view = _dev(env, parent)[:, off:off + n]
    got = env['ops'].filtfilt_sos(view, sos, padlen, cs).cpu().numpy()
    x = parent[:, off:off + n]
    ref = np.stack([np.concatenate([oracle.filtfilt(b, a, x[c, lo:min(lo + cs, n)])
                                    for lo in range(0, n, cs)]) for c in range(n_c)])
    assert got.shape == ref.shape
    assert rel_err(got, ref) < TIGHT[name] * 10, (name, n_c, cs, n_chunks, last, off)
