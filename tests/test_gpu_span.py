"""GPU tier: the single-pass span kernel of K1 (filt_span_kernel, opt-in with SS_FILT_SPAN=1)
against the oracle and against the two-sweep scan: same filter to float64 round-off for the
fast-decaying designs it accepts (the host bound span_err), any dtype / chunk geometry / view,
results independent of the tile range a launch stores (fused filter + detect equals the
separate passes bit for bit) and of the sharding; slow designs silently stay on the scan."""
import numpy as np
import pytest
from scipy import signal

from util import rel_err

pytestmark = pytest.mark.gpu

FAST = {
    'hp1_30k': lambda: signal.iirdesign(800 * 2 / 30000, 100 * 2 / 30000, 1, 10, ftype='butter'),
    'hp1_25k': lambda: signal.iirdesign(800 * 2 / 25000, 100 * 2 / 25000, 1, 10, ftype='butter'),
    'bp2': lambda: signal.butter(1, [0.1, 0.4], 'bandpass'),
    'lp3': lambda: signal.butter(3, 0.2),
    'bp4': lambda: signal.butter(2, [0.1, 0.4], 'bandpass'),
}


@pytest.fixture
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import spikesort_b200 as ss
    from spikesort_b200 import _ops, synth
    return {'torch': torch, 'dev': torch.device("cuda", 0), 'ss': ss, 'ops': _ops, 'synth': synth}


def _dev(env, a):
    return env['torch'].from_numpy(np.ascontiguousarray(a)).to(env['dev'])


@pytest.mark.parametrize("name", sorted(FAST))
@pytest.mark.parametrize("dtype", ["int16", "float32", "float64"])
def test_span_kernel_vs_oracle_and_scan(env, oracle, monkeypatch, name, dtype):
    from spikesort_b200._design import tf2sos_exact
    b, a = FAST[name]()
    sos, padlen = tf2sos_exact(b, a)
    rng = np.random.default_rng(3)
    # 3 contacts; chunks of 9000 (17-18 tiles: more than one span, ragged), last chunk 1234;
    # a DC offset far above the signal (what the one-tile truncation would show first)
    n, cs = 3 * 9000 + 1234, 9000
    x = rng.standard_normal((3, n)) * 300 + 5000
    x = np.round(x).astype(np.int16) if dtype == "int16" else x.astype(dtype)
    ref = np.empty((3, n))
    for c in range(3):
        for lo in range(0, n, cs):
            ref[c, lo:lo + cs] = oracle.filtfilt(b, a, x[c, lo:lo + cs])
    monkeypatch.setenv("SS_FILT_SPAN", "0")
    scan = env['ops'].filtfilt_sos(_dev(env, x), sos, padlen, cs).cpu().numpy()
    monkeypatch.setenv("SS_FILT_SPAN", "1")
    span = env['ops'].filtfilt_sos(_dev(env, x), sos, padlen, cs).cpu().numpy()
    scale = max(np.abs(ref).max(), 1.0)
    print("%s %s: span vs oracle %.3g, scan vs oracle %.3g, span vs scan %.3g" % (
        name, dtype, np.abs(span - ref).max() / scale, np.abs(scan - ref).max() / scale,
        np.abs(span - scan).max() / scale))
    assert not np.array_equal(span, scan) or name == 'never'      # it really is another kernel
    assert np.abs(span - ref).max() <= 2e-12 * max(scale, np.abs(x).max())
    assert np.abs(span - scan).max() <= 2e-12 * max(scale, np.abs(x).max())


def test_span_kernel_geometries_views_and_ranges(env, oracle, monkeypatch):
    """Chunk lengths around the tile / span sizes (one tile, a last span of one tile, ...),
    int16 wrap of the odd extension, a strided view at an odd sample, and the fused call whose
    two launches store different tile ranges of the same chunk."""
    from spikesort_b200._design import tf2sos_exact
    torch, ops = env['torch'], env['ops']
    monkeypatch.setenv("SS_FILT_SPAN", "1")
    b, a = FAST['hp1_30k']()
    sos, padlen = tf2sos_exact(b, a)
    rng = np.random.default_rng(5)
    T, W = 512, 14
    for n, cs in [(7, 7), (500, 500), (512, 512), (1000, 100), (T * W - 12 - 4, T * W - 12 - 4),
                  (T * (W + 1) - 12 - 500, T * (W + 1) - 12 - 500), (T * 2 * W + 77, T * 2 * W + 77),
                  (40000, 1000000), (2999, 1000)]:
        x = np.round(rng.standard_normal((2, n)) * 9000).astype(np.int16)
        x[:, 0] = 30000                      # 2*x[0] - x[k] wraps around in int16
        x[:, -1] = -30000
        ref = np.empty((2, n))
        for c in range(2):
            for lo in range(0, n, cs):
                ref[c, lo:lo + cs] = oracle.filtfilt(b, a, x[c, lo:lo + cs])
        got = ops.filtfilt_sos(_dev(env, x), sos, padlen, cs).cpu().numpy()
        assert rel_err(got, ref) < 1e-12, (n, cs)
    big = np.round(rng.standard_normal((4, 60001)) * 100).astype(np.int16)
    view = _dev(env, big)[:, 501:50501]                      # odd start: element loads
    got = ops.filtfilt_sos(view, sos, padlen, 25000).cpu().numpy()
    ref = np.stack([np.concatenate([oracle.filtfilt(b, a, big[c, 501 + lo:501 + lo + 25000])
                                    for lo in (0, 25000)]) for c in range(4)])
    assert rel_err(got, ref) < 1e-12
    # fused filter + auto threshold + marks == separate passes, bit for bit
    FS = 1000
    for n, cs in [(50000, 3000), (50000, 50000), (8192 * 3, 8192)]:
        xs = env['synth'].make_recording(4, n, FS, seed=9, rate_hz=40.0, noise=7.0,
                                         amp=(60.0, 160.0)).to(env['dev'])
        flt = env['ss'].filters.Filter(FS * 0.3, FS * 0.04)
        s2, p2 = flt.sections(FS)
        y = ops.filtfilt_sos(xs, s2, p2, cs)
        thr = ops.var_threshold(y, int(10 * FS), -1.5)
        spt, off = ops.detect(y, thr, True, FS)
        y2, thr2, spt2, off2 = ops.filtfilt_detect(xs, s2, p2, cs, FS, True, thr=None, n0=int(10 * FS),
                                                   factor=-1.5)
        assert torch.equal(y, y2) and torch.equal(thr, thr2) and torch.equal(spt, spt2) and \
            torch.equal(off, off2), (n, cs)


def test_span_kernel_declines_slow_filters(env, oracle, monkeypatch):
    """max |pole| 0.98 (the order-8 band-pass) does not decay within a tile: SS_FILT_SPAN=1
    must leave it on the exact scan - identical output with and without the switch."""
    from spikesort_b200._design import tf2sos_exact
    b, a = signal.iirdesign(np.array([300, 6000]) * 2 / 30000, np.array([100, 8000]) * 2 / 30000,
                            1, 10, ftype='butter')
    sos, padlen = tf2sos_exact(b, a)
    x = np.round(np.random.default_rng(1).standard_normal((2, 30000)) * 200).astype(np.int16)
    monkeypatch.setenv("SS_FILT_SPAN", "0")
    y0 = env['ops'].filtfilt_sos(_dev(env, x), sos, padlen, 10000)
    monkeypatch.setenv("SS_FILT_SPAN", "1")
    y1 = env['ops'].filtfilt_sos(_dev(env, x), sos, padlen, 10000)
    assert env['torch'].equal(y0, y1)


@pytest.mark.parametrize("S", [2, 3])
def test_span_kernel_sharded_equals_whole(cuda, monkeypatch, S):
    """The sharded chain with the span kernel: shards are whole chunks, a unit's arithmetic does
    not depend on where it is processed -> bit-identical to the unsharded chain."""
    monkeypatch.setenv("SS_FILT_SPAN", "1")
    import test_gpu_sharded as tgs
    from spikesort_b200 import filters
    tgs._check_sharded(cuda, S, 10000, "fabric-capacity", filters.Filter(300., 40.))
