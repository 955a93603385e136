"""CPU tier: host logic above the C ABI - filter factorisation and the exact float64
expressions the host evaluates for the kernels."""
import numpy as np
import pytest
from scipy import signal

from spikesort_b200 import _ops
from spikesort_b200._design import tf2sos_exact


DESIGNS = {
    'hp1': lambda: signal.iirdesign(800 * 2 / 25000, 100 * 2 / 25000, 1, 10, ftype='butter'),
    'bp8': lambda: signal.iirdesign(np.array([300, 6000]) * 2 / 30000,
                                    np.array([100, 8000]) * 2 / 30000, 1, 10, ftype='butter'),
    'ellip10': lambda: signal.iirdesign(np.array([300, 6000]) * 2 / 30000,
                                        np.array([150, 7000]) * 2 / 30000, 1, 40, ftype='ellip'),
    'butter12': lambda: signal.butter(6, [300 / 15000, 6000 / 15000], 'bandpass'),
    'lp3': lambda: signal.butter(3, 0.2),
    'ellip_hp3': lambda: signal.iirdesign(0.02, 0.016, 1, 10, ftype='ellip'),
    'cheby2_lp': lambda: signal.iirdesign(0.2, 0.3, 1, 40, ftype='cheby2'),
}


@pytest.mark.parametrize("name", sorted(DESIGNS))
def test_sections_are_the_same_filter(name):
    b, a = DESIGNS[name]()
    sos, padlen = tf2sos_exact(b, a)
    assert padlen == 3 * max(len(a), len(b))
    assert np.all(sos[:, 3] == 1.0)
    order = max(len(a), len(b)) - 1
    n_states = sum(1 if (s[2] == 0 and s[5] == 0) else 2 for s in sos)
    assert n_states in (order, order + 1)     # odd orders may carry one idle delay
    first_order = [i for i, s in enumerate(sos) if s[2] == 0 and s[5] == 0]
    assert first_order in ([], [len(sos) - 1])
    # same frequency response as (b, a)
    w, h0 = signal.freqz(b, a, worN=2048)
    _, h1 = signal.sosfreqz(sos, worN=2048)
    # (freqz of a 12th-order (b, a) polynomial pair is itself only good to ~1e-7)
    assert np.abs(h0 - h1).max() <= 1e-6 * max(1.0, np.abs(h0).max())
    # and the same zero-phase output (the reference's (b, a) arithmetic is the noisier one)
    x = np.random.default_rng(0).standard_normal(4000)
    y0 = signal.filtfilt(b, a, x)
    y1 = signal.sosfiltfilt(sos, x, padlen=padlen)
    assert np.abs(y0 - y1).max() <= 1e-6 * np.abs(y0).max()


def test_unstable_and_fir_rejected():
    with pytest.raises(ValueError):
        tf2sos_exact([1.0, 0.5], [1.0, -1.5])
    with pytest.raises(NotImplementedError):
        tf2sos_exact([0.25, 0.5, 0.25], [1.0])


@pytest.mark.parametrize("FS", [25000, 25E3, 30000, 10000, 44100.0])
@pytest.mark.parametrize("sp_win", [[-0.2, 0.8], [-0.6, 0.8], [0, 0.4], [-1.0, 2.0]])
def test_window_and_bounds_match_reference_expressions(FS, sp_win):
    win0, win1, time = _ops.window_params(sp_win, FS)
    win = (np.asarray(sp_win) / 1000.0 * FS).astype(np.int32)           # extract.py:151
    assert (win0, win1) == (int(win[0]), int(win[1]))
    assert np.array_equal(time, np.arange(win[1] - win[0]) * 1000.0 / FS + sp_win[0])
    n = 123457
    t_min, t_max = _ops.time_bounds(n, FS, sp_win)
    max_time = n * 1000.0 / FS                                          # extract.py:106-109
    assert t_min == np.max((-sp_win[0], 0)) and t_max == np.min((max_time, max_time - sp_win[1]))


def test_known_truncation_quirks():
    # sp_win=[-0.6, 0.8] at 25 kHz gives win=[-14, 20], not [-15, 20]
    assert _ops.window_params([-0.6, 0.8], 25000)[:2] == (-14, 20)
    assert _ops.window_params([-0.2, 0.8], 25000)[:2] == (-5, 20)


@pytest.mark.parametrize("FS,sp_win,resample", [(25000, [-0.2, 0.8], 10), (30000, [-0.6, 0.8], 3),
                                                (25000, [0, 4.0], 2), (10000, [-0.2, 0.2], 7)])
def test_spline_plan_bit_exact_vs_scipy(FS, sp_win, resample):
    """Host half of the spline upsampling (spikesort_b200/_spline.py): the plan, replayed in
    plain Python in the order the kernels use, reproduces scipy splrep(s=0)/splev - what
    resample_spikes (extract.py:199-203) calls - bit for bit."""
    from scipy import interpolate
    from spikesort_b200._spline import spline_plan
    win = (np.asarray(sp_win) / 1000.0 * FS).astype(np.int32)
    time = np.arange(win[1] - win[0]) * 1000.0 / FS + sp_win[0]
    plan = spline_plan(time, FS * resample)
    assert np.array_equal(plan.new_time, np.arange(time[0], time[-1], 1000.0 / (FS * resample)))
    rng = np.random.default_rng(11)
    for trial in range(20):
        y = (rng.standard_normal(len(time)) * 10.0 ** rng.integers(-3, 4)).astype(np.float32)
        if trial == 0:
            y[:] = 0
        if trial == 1:
            y = np.round(y)                                  # many exact ties
        tck = interpolate.splrep(time, y, s=0)
        assert np.array_equal(tck[0], plan.knots)
        assert np.array_equal(plan.apply(y), interpolate.splev(plan.new_time, tck, der=0))
    with pytest.raises(TypeError):
        spline_plan(time[:3], FS * resample)                # splrep: m > k must hold


def test_slab_bounds_host_logic():
    """Time slabs of the upload-overlapped front end (pipeline._slab_bounds): whole chunks, in
    order, covering the recording once; the first slab holds the auto-threshold samples; a
    short trailing chunk joins the last slab; nothing to slice -> None."""
    from spikesort_b200.pipeline import _slab_bounds
    b = _slab_bounds(18000000, 1e6, 300000, 6)
    assert b == [(k * 3000000, (k + 1) * 3000000) for k in range(6)]
    for n, cs, n0, want in [(170500, 10000, 30000, 6), (70001, 30000, 0, 6), (3000000, 1e6, 300000, 6),
                            (999999, 1000, 5000, 9), (10 ** 7 + 1, 10 ** 6, 0, 3)]:
        b = _slab_bounds(n, cs, n0, want)
        assert b[0][0] == 0 and b[-1][1] == n and len(b) <= want + 1
        assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert all(a % int(cs) == 0 for a, _ in b)
        assert b[0][1] >= min(n0, n)
    assert _slab_bounds(60000, 25000, 250000, 6) is None       # threshold needs everything
    assert _slab_bounds(1500000, 1e6, 300000, 6) is None       # one full chunk only
    assert _slab_bounds(500, 1000, 0, 6) is None
