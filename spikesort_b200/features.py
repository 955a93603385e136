"""Drop-in for the feature stage of ``spike_sort.core.features`` (reference
``src/spike_sort/core/features.py``): ``PCA``, ``fetPCA``, the cheap per-spike
features and ``combine`` / ``normalize``.

The reference promotes whatever it is given to float64 (``features.py:224``); so do
the kernels - float32 blocks (``extract_spikes``), float64 blocks
(``resample_spikes``, ``extract_spikes(resample != 1)``) and int16 are read as they
are, any other dtype is widened to float64 on the host first (exact).  Scores agree
with the reference to float64 round-off up to the sign of each component (the sign
of an eigenvector is arbitrary in LAPACK too); blocks of >= 65536 spikes per contact
that are float32 take the tensor-core Gram (3xTF32, ~1e-7 relative on the covariance,
``SS_PCA_TC=0`` pins the float64 kernel).
"""
import functools

import numpy as np

from . import _lib, _ops

_MASK_KEY = 'is_valid'
_BAD_CONTACTS = "contacts must be either 'all' or a list of valid contact indices"


# ----------------------------------------------------------------------------
# waveform blocks on the device
# ----------------------------------------------------------------------------
def _select_contacts(block, contacts):
    """``features.py:298-307``: `contacts` is 'all' or an index (list) into the last axis."""
    if isinstance(contacts, str) and contacts == "all":
        return block
    try:
        return block[:, :, np.asarray(contacts)]
    except IndexError:
        raise IndexError(_BAD_CONTACTS)


def _device_block(block):
    """(n_win, n_spk, n_c) CUDA tensor of a waveform array in a dtype the kernels read
    (int16 / float32 / float64); anything else becomes float64, as ``astype(float64)``
    in the reference would make it."""
    t = _lib.require_cuda()
    if isinstance(block, t.Tensor):
        dev = block
        if dev.dtype not in (t.int16, t.float32, t.float64):
            dev = dev.to(t.float64)
    else:
        host = np.asarray(block)
        if host.dtype not in (np.int16, np.float32, np.float64):
            host = host.astype(np.float64)
        dev = t.from_numpy(np.ascontiguousarray(host))
    if dev.dim() == 2:
        dev = dev.unsqueeze(2)
    return _lib.upload(dev).contiguous()


def _float_block(block):
    """Device block for the min / peak-to-peak kernels: float32 and float64 as they are
    (np.ptp / np.min keep the dtype), the rest widened to float64."""
    dev = _device_block(block)
    return dev if dev.dtype.is_floating_point else dev.to(_lib.torch().float64)


def _principal_axes(dev, ncomps, leading_only):
    gram, ssum, _, counts = _ops.pca_gram(dev)
    evals, evecs = _ops.pca_eig(gram, ssum, counts, top=ncomps if leading_only else None)
    return evals, evecs, _ops.pca_project(dev, evals, evecs, ncomps)


# ----------------------------------------------------------------------------
# mask propagation (features.py:172-181)
# ----------------------------------------------------------------------------
def add_mask(feature_function):
    """Decorator: a feature set computed from waveforms that carry a validity mask carries
    the same mask."""

    @functools.wraps(feature_function, assigned=('__doc__',), updated=())
    def with_mask(source, *args, **kwargs):
        result = feature_function(source, *args, **kwargs)
        try:
            result[_MASK_KEY] = source[_MASK_KEY]
        except KeyError:
            pass
        return result
    return with_mask


def _named(prefix_fmt, n_channels, suffixes=(None,)):
    return [prefix_fmt % ((ch,) if sfx is None else (ch, sfx))
            for ch in range(n_channels) for sfx in suffixes]


# ----------------------------------------------------------------------------
# PCA
# ----------------------------------------------------------------------------
def PCA(data, ncomps=2):
    """``features.py:200-232``: (evals, evecs, score) of an (n_vars, n_obs) array of any
    real dtype; score is (ncomps, n_obs)."""
    dev = _device_block(np.asarray(data)[:, :, np.newaxis])
    evals, evecs, score = _principal_axes(dev, ncomps, leading_only=False)
    return (evals[0].cpu().numpy(), evecs[0].cpu().numpy(),
            np.ascontiguousarray(score.cpu().numpy().T))


@add_mask
def fetPCA(spikes_data, ncomps=2, contacts='all'):
    """``features.py:310-346``: the first `ncomps` whitened principal components of every
    contact; columns ``Ch{i}:PC{j}``, contact-major.  A 2-D block counts as one contact."""
    dev = _device_block(_select_contacts(spikes_data["data"], contacts))
    _, _, score = _principal_axes(dev, ncomps, leading_only=True)
    from .pipeline import to_host
    return {'data': to_host(score),
            'names': _named("Ch%d:PC%d", dev.shape[2], range(ncomps))}


# ----------------------------------------------------------------------------
# "next" rows: cheap features, combine(), normalize()
# ----------------------------------------------------------------------------
@add_mask
def fetP2P(spikes_data, contacts='all'):
    """``features.py:445-479``: peak-to-peak amplitude per contact (``np.ptp`` over time,
    in the waveforms' dtype)."""
    dev = _float_block(_select_contacts(spikes_data["data"], contacts))
    _, p2p = _ops.wave_minmax(dev, want_min=False, want_p2p=True)
    return {'data': p2p.cpu().numpy(), 'names': _named("Ch%d:P2P", dev.shape[2])}


@add_mask
def fetMin(spikes_data, contacts='all'):
    """``features.py:481-489``: minimum of the waveform per contact."""
    dev = _float_block(_select_contacts(spikes_data["data"], contacts))
    lowest, _ = _ops.wave_minmax(dev, want_min=True, want_p2p=False)
    return {'data': lowest.cpu().numpy(), 'names': _named("Ch%d:Min", dev.shape[2])}


@add_mask
def fetSpIdx(spikes_data):
    """``features.py:491-500``: running index of the spike (nothing to compute on a GPU)."""
    n_spikes = np.shape(spikes_data["data"])[1]
    return {'data': np.arange(n_spikes).reshape(-1, 1), 'names': ["SpIdx"]}


@add_mask
def fetSpTime(spt_dict):
    """``features.py:503-510``: spike time (ms) as a one-column feature."""
    return {'data': np.asarray(spt_dict['data'])[:, np.newaxis], 'names': ["SpTime"]}


def normalize(features, copy=True):
    """``features.py:184-197``: every column to [0, 1]: (x - min) / (max - min), float64."""
    out = dict(features) if copy else features
    t = _lib.require_cuda()
    table = np.ascontiguousarray(out['data'], dtype=np.float64)
    if table.size:
        table = _ops.normalize_columns(t.from_numpy(table).to(_lib.device())).cpu().numpy()
    out['data'] = table
    return out


def combine(feat_data, norm=True, feat_method_names=None):
    """``features.py:98-139``: feature sets side by side; the masks AND-ed; names prefixed by
    ``feat_method_names[i] + ':'`` when given; min-max normalised unless ``norm=False``."""
    sets = list(feat_data)
    prefixes = list(feat_method_names) if feat_method_names else []
    labels = []
    for k, fset in enumerate(sets):
        prefix = prefixes[k] + ':' if k < len(prefixes) else ''
        labels.extend(prefix + str(name) for name in fset['names'])
    try:
        table = np.hstack([fset['data'] for fset in sets])
        valid = None
        for fset in sets:
            if _MASK_KEY in fset:
                valid = fset[_MASK_KEY] if valid is None else np.logical_and(valid, fset[_MASK_KEY])
    except ValueError:
        raise ValueError('all features must contain the same number of spikes')
    combined = {"data": table, "names": labels}
    if valid is not None and len(valid):
        combined[_MASK_KEY] = valid
    return normalize(combined, copy=False) if norm else combined
