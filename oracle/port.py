"""TEST INFRASTRUCTURE - CPU oracle for the SpikeSort front end.

A NumPy (+ plain C for the filter recurrence, ``oracle/filtfilt_ref.c``)
restatement of the reference algorithms on the hot path.  It exists to CHECK the
CUDA library and to be timed as the CPU baseline; the product package
``spikesort_b200`` never imports it.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it.

Parity pin: ``tests/test_oracle.py`` checks every function here against the
golden vectors in ``tests/golden/`` (generated from the UNMODIFIED reference
files by ``tests/golden/make_golden.py`` through ``oracle/ref_shim.py``), against
the reference's own doctest known answers (``docs/source/datastructures.rst:
98-109,163-199``) and its unit-test expectations (``tests/test_core.py``,
``tests/test_features.py``); ``tests/test_oracle_vs_reference.py`` re-runs the
comparison live wherever ``/root/reference`` is present.

Every function cites the reference lines it restates
(paths relative to ``/root/reference/src/spike_sort/core/``).
"""
import ctypes
import os
import subprocess
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_filtfilt.so")
_lib = None

_DT = {np.dtype(np.int16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}


def build(force=False):
    """Compile oracle/filtfilt_ref.c with gcc (a few hundred ms)."""
    src = os.path.join(_HERE, "filtfilt_ref.c")
    if (not force and os.path.isfile(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


def _clib():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.orc_filtfilt.argtypes = [dp, dp, dp, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_int, ctypes.c_long, dp]
        lib.orc_filtfilt.restype = ctypes.c_int
        lib.orc_filter_proxy.argtypes = [dp, dp, dp, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_int, ctypes.c_long, ctypes.c_long,
                                         ctypes.c_long, ctypes.c_long, dp, ctypes.c_long]
        lib.orc_filter_proxy.restype = ctypes.c_int
        _lib = lib
    return _lib


def _as_dp(arr):
    return arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _pad_ba(b, a):
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    ntaps = max(len(a), len(b))
    bb = np.zeros(ntaps)
    aa = np.zeros(ntaps)
    bb[:len(b)] = b
    aa[:len(a)] = a
    return bb, aa, ntaps


def lfilter_zi(b, a):
    """scipy.signal.lfilter_zi (SciPy 1.18 formulation): steady state of the
    DF2T delays for a unit step, zi[k] = sum_{j>k} (b[j] - y_inf*a[j]) with
    y_inf = sum(b)/sum(a), after normalising a[0] to 1."""
    bb, aa, ntaps = _pad_ba(b, a)
    if aa[0] != 1:
        bb, aa = bb / aa[0], aa / aa[0]
    y_inf = np.sum(bb) / np.sum(aa)
    return np.flip(np.cumsum(np.flip(bb - y_inf * aa)))[1:]


def _supported(x):
    x = np.asarray(x)
    if x.dtype not in _DT:
        # NumPy promotes every other real dtype to float64 inside lfilter; the
        # odd extension of other integer widths is not modelled here.
        x = x.astype(np.float64)
    return np.ascontiguousarray(x)


# ---------------------------------------------------------------------------
# filters.py
# ---------------------------------------------------------------------------

def filtfilt(b, a, x):
    """scipy.signal.filtfilt(b, a, x) for one 1-D chunk, as called at
    filters.py:38,74,101 (defaults: odd padding, padlen 3*ntaps)."""
    bb, aa, ntaps = _pad_ba(b, a)
    x = _supported(x)
    if np.atleast_1d(a).size == 1:
        return _filtfilt_fir(np.atleast_1d(np.asarray(b, dtype=np.float64)) / np.atleast_1d(a)[0], x)
    out = np.empty(x.shape[0], dtype=np.float64)
    zi = np.ascontiguousarray(lfilter_zi(bb, aa))
    rc = _clib().orc_filtfilt(_as_dp(bb), _as_dp(aa), _as_dp(zi), ntaps, x.ctypes.data,
                              _DT[x.dtype], x.shape[0], _as_dp(out))
    if rc == -3:
        raise ValueError("The length of the input vector x must be greater "
                         "than padlen, which is %d." % (3 * ntaps))
    if rc:
        raise RuntimeError("oracle filtfilt failed (%d)" % rc)
    return out


def _odd_ext(x, n):
    """scipy.signal._arraytools.odd_ext, evaluated in the dtype of x as NumPy does."""
    left = 2 * x[0] - x[n:0:-1]
    right = 2 * x[-1] - x[-2:-(n + 2):-1]
    return np.concatenate((left, x, right))


def _lfilter_fir(b, x, zi):
    """scipy.signal.lfilter with a == [1]: SciPy special-cases FIR filters and evaluates
    them as np.convolve(b, x), adding zi to the first len(b)-1 outputs."""
    full = np.convolve(b, x)
    full[:len(zi)] += zi
    return full[:len(x)]


def _filtfilt_fir(b, x):
    """filtfilt(b, [1], x) as called at filters.py:74: odd padding by 3*len(b), forward
    pass from zi*ext[0], backward pass over the reversed output from zi*y[-1]."""
    ntaps = len(b)
    edge = 3 * ntaps
    if x.shape[0] <= edge:
        raise ValueError("The length of the input vector x must be greater "
                         "than padlen, which is %d." % edge)
    ext = _odd_ext(x, edge)
    zi = lfilter_zi(b, np.ones(1))
    y = _lfilter_fir(b, ext, zi * ext[0])
    y = _lfilter_fir(b, y[::-1], zi * y[-1])[::-1]
    return np.ascontiguousarray(y[edge:-edge], dtype=np.float64)


class FilterFir(object):
    """filters.py:41-74 - remez FIR of `order` taps + filtfilt.  The reference passes the
    sampling rate as remez(..., Hz=FS); SciPy renamed that keyword to fs (same meaning)."""

    def __init__(self, f_pass, f_stop, order):
        self.fp, self.fs, self.order = f_pass, f_stop, order
        self._coefs = {}

    def coefficients(self, FS):
        if FS not in self._coefs:
            from scipy import signal
            bands = [0, min(self.fs, self.fp), max(self.fs, self.fp), FS / 2]
            gains = [int(self.fp < self.fs), int(self.fp > self.fs)]
            self._coefs[FS] = (signal.remez(self.order, bands, gains, fs=FS), [1])
        return self._coefs[FS]

    def __call__(self, x, FS):
        b, a = self.coefficients(FS)
        return filtfilt(b, a, x)


class ZeroPhaseFilter(object):
    """filters.py:10-38 - iirdesign from a band and a transition width, then filtfilt."""

    def __init__(self, ftype, fband, tw=200., stop=20):
        self.ftype, self.fband, self.tw, self.gstop, self.gpass = ftype, fband, tw, stop, 1
        self._coefs = {}

    def coefficients(self, FS):
        if FS not in self._coefs:
            from scipy import signal
            wp = np.array(self.fband)
            ws = wp + np.array([-self.tw, self.tw])
            self._coefs[FS] = signal.iirdesign(wp=wp * 2.0 / FS, ws=ws * 2.0 / FS,
                                               gstop=self.gstop, gpass=self.gpass,
                                               ftype=self.ftype)
        return self._coefs[FS]

    def __call__(self, x, FS):
        b, a = self.coefficients(FS)
        return filtfilt(b, a, x)


class Filter(object):
    """filters.py:77-101 - iirdesign once per FS, then filtfilt per call."""

    def __init__(self, fpass, fstop, gpass=1, gstop=10, ftype='butter'):
        self.fp = np.asarray(fpass)
        self.fs = np.asarray(fstop)
        self.gpass, self.gstop, self.ftype = gpass, gstop, ftype
        self._coefs = {}

    def coefficients(self, FS):
        if FS not in self._coefs:
            from scipy import signal
            self._coefs[FS] = signal.iirdesign(wp=self.fp * 2 / FS, ws=self.fs * 2 / FS,
                                               gpass=self.gpass, gstop=self.gstop,
                                               ftype=self.ftype)
        return self._coefs[FS]

    def __call__(self, x, FS):
        b, a = self.coefficients(FS)
        return filtfilt(b, a, x)


def filter_proxy(spikes, filter_obj, chunksize=1E6):
    """filters.py:104-143: every contact, every chunk of `chunksize` samples is
    filtered on its own (own padding, own initial state); output float64."""
    if filter_obj is None:
        return spikes
    data = spikes['data']
    out_dict = dict(spikes)
    n_c, n = data.shape
    cs = int(chunksize)
    out = np.zeros((n_c, n), dtype=np.float64)
    if isinstance(filter_obj, Filter):
        # same arithmetic, loop in C (used when timing the CPU baseline)
        b, a = filter_obj.coefficients(spikes['FS'])
        bb, aa, ntaps = _pad_ba(b, a)
        x = _supported(data)
        zi = np.ascontiguousarray(lfilter_zi(bb, aa))
        rc = _clib().orc_filter_proxy(_as_dp(bb), _as_dp(aa), _as_dp(zi), ntaps, x.ctypes.data,
                                      _DT[x.dtype], n_c, n, x.strides[0] // x.itemsize,
                                      cs, _as_dp(out), n)
        if rc == -3:
            raise ValueError("The length of the input vector x must be greater "
                             "than padlen, which is %d." % (3 * ntaps))
        if rc:
            raise RuntimeError("oracle filter_proxy failed (%d)" % rc)
    else:
        for c in range(n_c):
            for lo in range(0, n, cs):
                hi = min(lo + cs, n)
                out[c, lo:hi] = filter_obj(data[c, lo:hi], spikes['FS'])
    out_dict['data'] = out
    return out_dict


def fltLinearIIR(signal, fpass, fstop, gpass=1, gstop=10, ftype='butter'):
    """filters.py:146-171."""
    return filter_proxy(signal, Filter(fpass, fstop, gpass, gstop, ftype))


# ---------------------------------------------------------------------------
# extract.py
# ---------------------------------------------------------------------------

_RISING = ('rising', 'max')
_FALLING = ('falling', 'min')


def auto_threshold(row, FS, frac=8.0):
    """extract.py:79 - frac * population std of the first int(10*FS) samples."""
    return frac * np.sqrt(row[:int(10 * FS)].var(dtype=np.float64))


def detect_spikes(spike_data, thresh='auto', edge="rising", contact=0):
    """extract.py:43-94.  String thresholds are a multiple of the std of the
    first 10 s (negated for falling edges); numeric ones are used as given.  A
    crossing at i needs x[i] strictly on one side and x[i+1] strictly on the
    other; the time reported is that of sample i, in ms."""
    row = spike_data['data'][contact, :]
    FS = spike_data['FS']
    if isinstance(thresh, str):
        frac = 8.0 if thresh == 'auto' else float(thresh)
        thresh = auto_threshold(row, FS, frac)
        if edge in _FALLING:
            thresh = -thresh
    if edge not in _RISING + _FALLING:
        raise TypeError("'edge' parameter must be 'rising' or 'falling'")
    before, after = row[:-1], row[1:]
    if edge in _RISING:
        hit = (before < thresh) & (after > thresh)
    else:
        hit = (before > thresh) & (after < thresh)
    idx = np.flatnonzero(hit)
    return {'data': idx * 1000.0 / FS, 'thresh': thresh, 'contact': contact}


def _time_bounds(spike_data, sp_win):
    """extract.py:102-109: inclusive [t_min, t_max] (ms) of fully covered windows."""
    sp_data = spike_data['data']
    n_pts = sp_data.shape[1] if np.ndim(sp_data) > 1 else len(sp_data)
    max_time = n_pts * 1000.0 / spike_data['FS']
    t_min = np.max((-sp_win[0], 0))
    t_max = np.min((max_time, max_time - sp_win[1]))
    return t_min, t_max


def filter_spt(spike_data, spt_dict, sp_win):
    """extract.py:97-111."""
    t_min, t_max = _time_bounds(spike_data, sp_win)
    spt = spt_dict['data']
    return np.flatnonzero((spt >= t_min) & (spt <= t_max))


def _window(sp_win, FS):
    """extract.py:151-152: truncating int32 window and its float64 time axis."""
    win = (np.asarray(sp_win) / 1000.0 * FS).astype(np.int32)
    time = np.arange(win[1] - win[0]) * 1000.0 / FS + sp_win[0]
    return win, time


def extract_spikes(spike_data, spt_dict, sp_win, resample=1, contacts='all'):
    """extract.py:114-184.  float32 (n_win, n_spk, n_c); windows that leave the
    recording are clipped, zero-filled and flagged through 'is_valid' (present only
    when at least one spike is clipped).  resample != 1 hands the result to
    resample_spikes (:179-183), which drops 'is_valid'."""
    sp_data = spike_data['data']
    if isinstance(contacts, str) and contacts == 'all':
        contacts = np.arange(spike_data['n_contacts'])
    elif isinstance(contacts, (int, np.integer)):
        contacts = np.array([contacts])
    else:
        contacts = np.asarray(contacts)
    FS = spike_data['FS']
    spt = spt_dict['data']
    n_spk = len(spt)
    inner = np.zeros(n_spk, dtype=bool)
    inner[filter_spt(spike_data, spt_dict, sp_win)] = True
    centre = (spt / 1000.0 * FS).astype(np.int32)
    win, time = _window(sp_win, FS)
    n_pts = sp_data.shape[1]
    waves = np.zeros((len(time), n_spk, len(contacts)), dtype=np.float32)
    for s in range(n_spk):
        first = int(centre[s]) + int(win[0])
        last = int(centre[s]) + int(win[1])
        if inner[s]:
            lo, hi = first, last
        else:
            lo = max(min(n_pts, first), 0)
            hi = max(min(n_pts, last), 0)
            if lo == hi:
                continue
        waves[lo - first:hi - first, s, :] = np.atleast_2d(sp_data[contacts, lo:hi]).T
    out = {'data': waves, 'time': time, 'FS': FS}
    if not inner.all():
        out['is_valid'] = inner
    if resample != 1:
        warnings.warn("resample argument is deprecated."
                      "Please update your code to use function"
                      "resample_spikes", DeprecationWarning)
        out = resample_spikes(out, FS * resample)
    return out


def resample_spikes(spikes_dict, FS_new):
    """extract.py:187-205: every (spike, contact) waveform is interpolated by the cubic
    B-spline splrep(time, wave, s=0) (SciPy/FITPACK, not-a-knot ends) and evaluated on
    arange(time[0], time[-1], 1000/FS_new); float64 out, 'FS' stays the ORIGINAL rate."""
    from scipy import interpolate
    waves, time = spikes_dict['data'], spikes_dict['time']
    fine = np.arange(time[0], time[-1], 1000.0 / FS_new)
    n_spk, n_c = waves.shape[1], waves.shape[2]
    up = np.empty((len(fine), n_spk, n_c))
    for s in range(n_spk):
        for c in range(n_c):
            up[:, s, c] = interpolate.splev(fine, interpolate.splrep(time, waves[:, s, c], s=0), der=0)
    return {'data': up, 'time': fine, 'FS': spikes_dict['FS']}


def remove_doubles(spt_dict, tol):
    """extract.py:277-288: keep the first time and every time whose distance to
    its predecessor, in units of `tol`, truncates to more than 1."""
    out = dict(spt_dict)
    spt = spt_dict['data']
    if len(spt) > 0:
        gap = (np.diff(spt) / tol).astype('int')
        keep = np.ones(len(spt), dtype=bool)
        keep[1:] = gap > 1
        spt = spt[keep]
    out['data'] = spt
    return out


def align_spikes(spike_data, spt_dict, sp_win, type="max", resample=1,
                 contact=0, remove=True):
    """extract.py:208-274.  Each in-bounds spike is moved to the first extremum of its
    float32 window on `contact` (of the spline-upsampled float64 window for resample != 1); spikes whose extremum sat
    within 0.1 ms of a window edge are re-examined from the new position until
    none is left.  Out-of-bounds spikes are kept unmoved."""
    tol = 0.1
    if (sp_win[0] > -tol) or (sp_win[1] < tol):
        warnings.warn('You are using very short sp_win. '
                      'This may lead to alignment problems.')
    spt = spt_dict['data'].copy()
    todo = np.arange(len(spt))
    while len(todo) > 0:
        cand = {'data': spt[todo]}
        ok = filter_spt(spike_data, cand, sp_win)
        todo = todo[ok]
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', DeprecationWarning)
            waves = extract_spikes(spike_data, cand, sp_win, resample=resample, contacts=contact)
        w = waves['data'][:, ok, 0]
        where = w.argmax(0) if type == "max" else w.argmin(0)
        shift = waves['time'][where]
        spt[todo] += shift
        todo = todo[(shift < (sp_win[0] + tol)) | (shift > (sp_win[1] - tol))]
    out = {'data': spt}
    if remove:
        out = remove_doubles(out, 1000.0 / spike_data['FS'])
    return out


# ---------------------------------------------------------------------------
# features.py
# ---------------------------------------------------------------------------

def PCA(data, ncomps=2):
    """features.py:200-232: eigen-decomposition of np.cov (ddof=1) of the
    (n_vars, n_obs) matrix; scores are projections of the UN-centred data
    divided by sqrt(|eigenvalue|)."""
    data = data.astype(np.float64)
    K = np.cov(data)
    evals, evecs = np.linalg.eig(K)
    order = np.argsort(evals)[::-1]
    evecs = np.real(evecs[:, order])
    evals = np.abs(evals[order])
    score = np.dot(evecs[:, :ncomps].T, data) / np.sqrt(evals[:ncomps, np.newaxis])
    return evals, evecs, score


def fetPCA(spikes_data, ncomps=2, contacts='all'):
    """features.py:298-346 incl. the add_mask decorator (172-181): per-contact
    PCA over the time axis, columns contact-major, 'is_valid' passed through."""
    spikes = spikes_data['data']
    if not (isinstance(contacts, str) and contacts == 'all'):
        try:
            spikes = spikes[:, :, np.asarray(contacts)]
        except IndexError:
            raise IndexError("contacts must be either 'all' or a "
                             "list of valid contact indices")
    if spikes.ndim == 3:
        n_ch = spikes.shape[2]
        sc = np.vstack([PCA(spikes[:, :, c], ncomps)[2] for c in range(n_ch)])
    else:
        n_ch = 1
        sc = PCA(spikes, ncomps)[2]
    out = {'data': np.asarray(sc).T,
           'names': ["Ch%d:PC%d" % (c, j) for c in range(n_ch) for j in range(ncomps)]}
    if 'is_valid' in spikes_data:
        out['is_valid'] = spikes_data['is_valid']
    return out


def _select(spikes_data, contacts):
    spikes = spikes_data['data']
    if not (isinstance(contacts, str) and contacts == 'all'):
        try:
            spikes = spikes[:, :, np.asarray(contacts)]
        except IndexError:
            raise IndexError("contacts must be either 'all' or a "
                             "list of valid contact indices")
    return spikes


def _masked(out, spikes_data):
    if 'is_valid' in spikes_data:
        out['is_valid'] = spikes_data['is_valid']
    return out


def fetP2P(spikes_data, contacts='all'):
    """features.py:445-479: peak-to-peak amplitude per (spike, contact)."""
    p2p = np.ptp(_select(spikes_data, contacts), axis=0)
    if p2p.ndim < 2:
        p2p = p2p[:, np.newaxis]
    return _masked({'data': p2p, 'names': ["Ch%d:P2P" % i for i in range(p2p.shape[1])]},
                   spikes_data)


def fetMin(spikes_data, contacts='all'):
    """features.py:481-489."""
    mn = np.min(_select(spikes_data, contacts), axis=0)
    return _masked({'data': mn, 'names': ["Ch%d:Min" % i for i in range(mn.shape[1])]},
                   spikes_data)


def normalize(features, copy=True):
    """features.py:184-197: shift every column to 0, scale to 1."""
    out = features.copy() if copy else features
    data = out['data']
    data = data - data.min(0)[np.newaxis, :]
    out['data'] = data / data.max(0)[np.newaxis, :]
    return out


def combine(feat_data, norm=True, feat_method_names=None):
    """features.py:98-139."""
    names = [list(d['names']) for d in feat_data]
    masks = [d['is_valid'] for d in feat_data if 'is_valid' in d]
    try:
        data = np.hstack([d['data'] for d in feat_data])
    except ValueError:
        raise ValueError('all features must contain the same number of spikes')
    if feat_method_names:
        for i, method in enumerate(feat_method_names):
            names[i] = [method + ':' + n for n in names[i]]
    out = {'data': data, 'names': list(np.concatenate(names))}
    if masks:
        m = masks[0]
        for other in masks[1:]:
            m = np.logical_and(m, other)
        out['is_valid'] = m
    if norm:
        normalize(out, copy=False)
    return out


# ---------------------------------------------------------------------------
# the multi-contact chain the bench times
# ---------------------------------------------------------------------------

def chain_per_contact(raw, filt, sp_win, edge='falling', align_type='min', ncomps=2,
                      thresh='auto', contacts=None, chunksize=1E6):
    """filter -> for every contact c: detect(c) -> align(c) -> extract(c) -> fetPCA.
    Returns (filtered dict, {c: (spt_detected, spt_aligned, sp_waves, features)})."""
    filtered = filter_proxy(raw, filt, chunksize) if filt is not None else raw
    res = {}
    for c in (range(raw['n_contacts']) if contacts is None else contacts):
        spt = detect_spikes(filtered, thresh=thresh, edge=edge, contact=c)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spa = align_spikes(filtered, spt, sp_win, type=align_type, contact=c)
        waves = extract_spikes(filtered, spa, sp_win, contacts=c)
        feats = fetPCA(waves, ncomps=ncomps) if len(spa['data']) > 1 else None
        res[c] = (spt, spa, waves, feats)
    return filtered, res
