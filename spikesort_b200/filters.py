"""Drop-in for ``spike_sort.core.filters`` (reference
``src/spike_sort/core/filters.py``): same names, arguments and dict conventions,
the arithmetic on the B200.

``filter_proxy`` keeps the reference's semantics - every contact is filtered in
independent chunks of ``chunksize`` samples, each with its own odd padding and
steady-state initial conditions (``filters.py:135-141``) - but runs all
(contact, chunk) units in three kernel launches (one for the FIR filter) and leaves the float64 result in
HBM inside a :class:`DeviceSignal` (the reference leaves it in an HDF5 CArray).
"""
import numpy as np

from . import _lib, _ops
from ._design import tf2sos_exact
from ._signal import DeviceSignal


class _IIRBase(object):
    """Coefficients on the host with SciPy (as the reference), application on the GPU."""

    def _design_filter(self, FS):                      # pragma: no cover - abstract
        raise NotImplementedError

    def sections(self, FS):
        """(sos, padlen) of the filter the reference would apply at this FS."""
        if FS not in self._sos_cache:
            b, a = self._design_filter(FS)
            self._sos_cache[FS] = tf2sos_exact(b, a)
        return self._sos_cache[FS]

    def apply(self, x, FS, chunksize):
        """CUDA recording (n_contacts, n_samples) -> CUDA float64, chunk by chunk."""
        sos, padlen = self.sections(FS)
        return _ops.filtfilt_sos(x, sos, padlen, chunksize)

    def __call__(self, x, FS):
        """filtfilt of one 1-D signal; returns a float64 ndarray (``filters.py:99-101``)."""
        arr = np.asarray(x)
        if arr.ndim != 1:
            raise ValueError("expected a 1-D signal")
        dev = _lib.to_device_2d(arr)
        return self.apply(dev, FS, max(dev.shape[1], 1))[0].cpu().numpy()


class Filter(_IIRBase):
    """``filters.py:77-101``: IIR design from pass/stop edges in Hz."""

    def __init__(self, fpass, fstop, gpass=1, gstop=10, ftype='butter'):
        self.ftype = ftype
        self.fp = np.asarray(fpass)
        self.fs = np.asarray(fstop)
        self._coefs_cache = {}
        self._sos_cache = {}
        self.gstop = gstop
        self.gpass = gpass

    def _design_filter(self, FS):
        if FS not in self._coefs_cache:
            from scipy import signal
            wp, ws = self.fp * 2 / FS, self.fs * 2 / FS
            self._coefs_cache[FS] = signal.iirdesign(wp=wp, ws=ws, gstop=self.gstop,
                                                     gpass=self.gpass, ftype=self.ftype)
        return self._coefs_cache[FS]


class ZeroPhaseFilter(_IIRBase):
    """``filters.py:10-38``: band given as ``fband`` with transition width ``tw``."""

    def __init__(self, ftype, fband, tw=200., stop=20):
        self.gstop = stop
        self.gpass = 1
        self.fband = fband
        self.tw = tw
        self.ftype = ftype
        self._coefs_cache = {}
        self._sos_cache = {}

    def _design_filter(self, FS):
        if FS not in self._coefs_cache:
            from scipy import signal
            wp = np.array(self.fband)
            ws = wp + np.array([-self.tw, self.tw])
            wp, ws = wp * 2.0 / FS, ws * 2.0 / FS
            self._coefs_cache[FS] = signal.iirdesign(wp=wp, ws=ws, gstop=self.gstop,
                                                     gpass=self.gpass, ftype=self.ftype)
        return self._coefs_cache[FS]


class FilterFir(object):
    """``filters.py:41-74``: remez FIR of `order` taps, applied forwards and backwards.
    (The reference passes the rate to remez as ``Hz=FS``; SciPy now calls it ``fs``.)"""

    def __init__(self, f_pass, f_stop, order):
        self._coefs_cache = {}
        self.fp = f_pass
        self.fs = f_stop
        self.order = order

    def _design_filter(self, FS):
        if FS not in self._coefs_cache:
            from scipy import signal
            bands = [0, min(self.fs, self.fp), max(self.fs, self.fp), FS / 2]
            gains = [int(self.fp < self.fs), int(self.fp > self.fs)]
            self._coefs_cache[FS] = (signal.remez(self.order, bands, gains, fs=FS), [1])
        return self._coefs_cache[FS]

    def apply(self, x, FS, chunksize):
        b, _ = self._design_filter(FS)
        return _ops.filtfilt_fir(x, b, chunksize)

    def __call__(self, x, FS):
        arr = np.asarray(x)
        if arr.ndim != 1:
            raise ValueError("expected a 1-D signal")
        dev = _lib.to_device_2d(arr)
        return self.apply(dev, FS, max(dev.shape[1], 1))[0].cpu().numpy()


def filter_proxy(spikes, filter_obj, chunksize=1E6):
    """``filters.py:104-143``.  Returns a copy of `spikes` whose 'data' is the
    float64 filtered recording, resident on the GPU."""
    if filter_obj is None:
        return spikes
    if not hasattr(filter_obj, "apply"):
        raise TypeError("filter_proxy on the CUDA path needs a spikesort_b200 filter object "
                        "(Filter / ZeroPhaseFilter / FilterFir); arbitrary Python callables "
                        "cannot run on the GPU and there is no CPU fallback")
    sp_dict = spikes.copy()
    x = _lib.to_device_2d(spikes['data'])
    out = filter_obj.apply(x, sp_dict['FS'], int(chunksize))
    sp_dict['data'] = DeviceSignal(out)
    return sp_dict


def fltLinearIIR(signal, fpass, fstop, gpass=1, gstop=10, ftype='butter'):
    """``filters.py:146-171``: acausal IIR filter of a raw-recording dict."""
    filt = Filter(fpass, fstop, gpass, gstop, ftype)
    return filter_proxy(signal, filt)


def clean_after_exit():
    """``filters.py:174-183`` removed the temporary HDF5 files; nothing to remove
    here (device buffers are freed with their DeviceSignal)."""
    return None
