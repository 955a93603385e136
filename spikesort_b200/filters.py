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


class _ZeroPhaseOnDevice(object):
    """A zero-phase (forward-backward) filter whose taps come from a host-side design
    routine and whose arithmetic runs on the GPU.

    Subclasses say what to design (`_spec`); this class memoises the taps per sampling
    rate (the reference's classes do the same, one design per FS), turns them into what
    the kernels take - exact second-order sections for a recursive filter, the plain taps
    for an FIR one - and applies them to a whole recording or to one 1-D signal."""

    def _spec(self, FS):                               # pragma: no cover - abstract
        """(b, a) for sampling rate FS."""
        raise NotImplementedError

    def _memo(self, name):
        return self.__dict__.setdefault(name, {})

    def _design_filter(self, FS):
        """(b, a), the pair the reference classes return from the method of this name."""
        taps = self._memo('_coefs_cache')
        if FS not in taps:
            taps[FS] = self._spec(FS)
        return taps[FS]

    def is_recursive(self, FS):
        """False for an FIR design (a == [1]): that one runs as a stencil, not as a scan."""
        return np.size(self._design_filter(FS)[1]) > 1

    def sections(self, FS):
        """(sos, padlen) of the recursive filter the reference would apply at this FS."""
        plans = self._memo('_sos_cache')
        if FS not in plans:
            plans[FS] = tf2sos_exact(*self._design_filter(FS))
        return plans[FS]

    def apply(self, x, FS, chunksize):
        """CUDA recording (n_contacts, n_samples) -> CUDA float64, chunk by chunk."""
        b, a = self._design_filter(FS)
        if not self.is_recursive(FS):                  # FIR: no recurrence, a stencil
            return _ops.filtfilt_fir(x, np.asarray(b, dtype=np.float64) / np.ravel(a)[0], chunksize)
        sos, padlen = self.sections(FS)
        return _ops.filtfilt_sos(x, sos, padlen, chunksize)

    def __call__(self, x, FS):
        """filtfilt of one 1-D signal -> float64 ndarray (``filters.py:36-38,72-74,99-101``)."""
        signal_1d = np.asarray(x)
        if signal_1d.ndim != 1:
            raise ValueError("expected a 1-D signal")
        dev = _lib.to_device_2d(signal_1d)
        return self.apply(dev, FS, max(dev.shape[1], 1))[0].cpu().numpy()


class Filter(_ZeroPhaseOnDevice):
    """``filters.py:77-101``: IIR design from pass / stop edges in Hz."""

    def __init__(self, fpass, fstop, gpass=1, gstop=10, ftype='butter'):
        self.fp, self.fs = np.asarray(fpass), np.asarray(fstop)
        self.gpass, self.gstop, self.ftype = gpass, gstop, ftype

    def _spec(self, FS):
        # the reference scales the edges as `f * 2 / FS` (filters.py:88); the same
        # expression here, so that the designed taps are the same bits
        return _scaled_iirdesign(self.fp * 2 / FS, self.fs * 2 / FS, self)


class ZeroPhaseFilter(_ZeroPhaseOnDevice):
    """``filters.py:10-38``: pass band `fband` (Hz) with transition width `tw` on both sides."""

    def __init__(self, ftype, fband, tw=200., stop=20):
        self.ftype, self.fband, self.tw = ftype, fband, tw
        self.gpass, self.gstop = 1, stop

    def _spec(self, FS):
        inner = np.array(self.fband)
        outer = inner + np.array([-self.tw, self.tw])
        return _scaled_iirdesign(inner * 2.0 / FS, outer * 2.0 / FS, self)


_DESIGNED = {}      # design arguments -> (b, a): fltLinearIIR designs the same filter on every call


def _scaled_iirdesign(wp, ws, owner):
    from scipy import signal
    key = (np.asarray(wp, dtype=np.float64).tobytes(), np.asarray(ws, dtype=np.float64).tobytes(),
           np.ndim(wp), float(owner.gpass), float(owner.gstop), str(owner.ftype))
    if key not in _DESIGNED:
        if len(_DESIGNED) > 256:
            _DESIGNED.clear()
        _DESIGNED[key] = signal.iirdesign(wp=wp, ws=ws, gpass=owner.gpass, gstop=owner.gstop,
                                          ftype=owner.ftype)
    return _DESIGNED[key]


class FilterFir(_ZeroPhaseOnDevice):
    """``filters.py:41-74``: `order`-tap remez FIR between f_pass and f_stop (Hz), applied
    forwards and backwards.  (The reference hands the rate to remez as ``Hz=FS``; SciPy now
    calls that keyword ``fs``.)"""

    def __init__(self, f_pass, f_stop, order):
        self.fp, self.fs, self.order = f_pass, f_stop, order

    def _spec(self, FS):
        from scipy import signal
        lo, hi = sorted((self.fs, self.fp))
        lowpass = self.fp < self.fs
        taps = signal.remez(self.order, [0, lo, hi, FS / 2], [int(lowpass), int(self.fp > self.fs)],
                            fs=FS)
        return taps, [1]


class _AdoptedFilter(_ZeroPhaseOnDevice):
    """A filter object of the REFERENCE package (anything with its `_design_filter(FS)` ->
    (b, a) method, e.g. an instance created before install()): its taps, our arithmetic."""

    def __init__(self, foreign):
        self._foreign = foreign

    def _spec(self, FS):
        return self._foreign._design_filter(FS)


def _on_device(filter_obj):
    if hasattr(filter_obj, "apply"):
        return filter_obj
    if hasattr(filter_obj, "_design_filter"):
        return _AdoptedFilter(filter_obj)
    raise TypeError("filter_proxy on the CUDA path needs a filter that exposes its taps "
                    "(Filter / ZeroPhaseFilter / FilterFir, or any object with "
                    "_design_filter(FS) -> (b, a)); an arbitrary Python callable cannot run "
                    "on the GPU and there is no CPU fallback")


def filter_proxy(spikes, filter_obj, chunksize=1E6):
    """``filters.py:104-143``.  Returns a copy of `spikes` whose 'data' is the
    float64 filtered recording, resident on the GPU."""
    if filter_obj is None:
        return spikes
    runner = _on_device(filter_obj)
    sp_dict = spikes.copy()
    x = _lib.to_device_2d(spikes['data'])
    out = runner.apply(x, sp_dict['FS'], int(chunksize))
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
