"""Drop-in for the PCA feature stage of ``spike_sort.core.features`` (reference
``src/spike_sort/core/features.py``): ``PCA``, ``fetPCA`` and the ``add_mask``
decorator.  Scores agree with the reference to float64 round-off up to the sign
of each component (the sign of an eigenvector is arbitrary in LAPACK too).
"""
import numpy as np

from . import _lib, _ops


def add_mask(feature_function):
    """``features.py:172-181``: copy 'is_valid' from the waveforms to the features."""

    def _decorated(spike_data, *args, **kwargs):
        feature_data = feature_function(spike_data, *args, **kwargs)
        if 'is_valid' in spike_data:
            feature_data['is_valid'] = spike_data['is_valid']
        return feature_data
    _decorated.__doc__ = feature_function.__doc__
    return _decorated


def _waves_on_device(spikes):
    """(n_win, n_spk, n_c) float32 CUDA tensor of a waveform array.  Other dtypes
    are cast to float32 only when that is exact (the reference promotes to
    float64); float64 input that is not float32-representable is rejected."""
    t = _lib.require_cuda()
    if isinstance(spikes, t.Tensor):
        w = spikes
    else:
        arr = np.asarray(spikes)
        if arr.dtype != np.float32:
            a32 = arr.astype(np.float32)
            if not np.array_equal(a32.astype(arr.dtype), arr):
                raise NotImplementedError("PCA on the CUDA path takes float32 waveforms (what "
                                          "extract_spikes produces)")
            arr = a32
        w = t.from_numpy(np.ascontiguousarray(arr))
    if w.dim() == 2:
        w = w.unsqueeze(2)
    return w.to(_lib.device()).to(t.float32).contiguous()


def _pca_device(w, ncomps, leading_only=False):
    gram, ssum, _, counts = _ops.pca_gram(w)
    evals, evecs = _ops.pca_eig(gram, ssum, counts, top=ncomps if leading_only else None)
    score = _ops.pca_project(w, evals, evecs, ncomps)
    return evals, evecs, score


def PCA(data, ncomps=2):
    """``features.py:200-232``: (evals, evecs, score) of an (n_vars, n_obs) array."""
    w = _waves_on_device(np.asarray(data)[:, :, np.newaxis])
    evals, evecs, score = _pca_device(w, ncomps)
    return (evals[0].cpu().numpy(), evecs[0].cpu().numpy(),
            np.ascontiguousarray(score.cpu().numpy().T))


def _get_data(spk_dict, contacts):
    """``features.py:298-307``."""
    spikes = spk_dict["data"]
    if not (isinstance(contacts, str) and contacts == "all"):
        contacts = np.asarray(contacts)
        try:
            spikes = spikes[:, :, contacts]
        except IndexError:
            raise IndexError("contacts must be either 'all' or a "
                             "list of valid contact indices")
    return spikes


@add_mask
def fetPCA(spikes_data, ncomps=2, contacts='all'):
    """``features.py:310-346``: per-contact principal components, columns
    ``Ch{i}:PC{j}`` contact-major."""
    spikes = _get_data(spikes_data, contacts)
    w = _waves_on_device(spikes)
    n_channels = w.shape[2]
    _, _, score = _pca_device(w, ncomps, leading_only=True)
    names = ["Ch%d:PC%d" % (i, j) for i in range(n_channels) for j in range(ncomps)]
    return {'data': score.cpu().numpy(), "names": names}


# --------------------------------------------------------------------------
# "next" rows: the cheap features every example combines with
# the PCA scores, and combine()/normalize()
# --------------------------------------------------------------------------

@add_mask
def fetP2P(spikes_data, contacts='all'):
    """``features.py:445-479``: peak-to-peak amplitude per contact (float32)."""
    w = _waves_on_device(_get_data(spikes_data, contacts))
    _, p2p = _ops.wave_minmax(w, want_min=False, want_p2p=True)
    p2p = p2p.cpu().numpy()
    names = ["Ch%d:P2P" % i for i in range(p2p.shape[1])]
    return {'data': p2p, 'names': names}


@add_mask
def fetMin(spikes_data, contacts='all'):
    """``features.py:481-489``: minimum of the waveform per contact (float32)."""
    w = _waves_on_device(_get_data(spikes_data, contacts))
    mn, _ = _ops.wave_minmax(w, want_min=True, want_p2p=False)
    mn = mn.cpu().numpy()
    names = ["Ch%d:Min" % i for i in range(mn.shape[1])]
    return {'data': mn, 'names': names}


@add_mask
def fetSpIdx(spikes_data):
    """``features.py:491-500``: sequential spike index (host side: nothing to compute)."""
    n_datapts = _get_data(spikes_data, [0]).shape[1]
    return {'data': np.arange(n_datapts)[:, np.newaxis], 'names': ["SpIdx"]}


@add_mask
def fetSpTime(spt_dict):
    """``features.py:503-510``: spike time in ms as a feature column."""
    spt = spt_dict['data']
    return {'data': spt[:, np.newaxis], 'names': ["SpTime"]}


def normalize(features, copy=True):
    """``features.py:184-197``: min-max normalisation of every column (float64)."""
    features_norm = features.copy() if copy else features
    t = _lib.require_cuda()
    data = np.ascontiguousarray(np.asarray(features_norm['data'], dtype=np.float64))
    if data.size:
        dev = t.from_numpy(data).to(_lib.device())
        data = _ops.normalize_columns(dev).cpu().numpy()
    features_norm['data'] = data
    return features_norm


def combine(feat_data, norm=True, feat_method_names=None):
    """``features.py:98-139``: stack feature sets column-wise, AND their masks, prefix
    names with the method names, normalise."""
    features = [d['data'] for d in feat_data]
    names = [list(d['names']) for d in feat_data]
    mask = [d['is_valid'] for d in feat_data if 'is_valid' in d]
    try:
        if mask:
            m = mask[0]
            for other in mask[1:]:
                m = np.logical_and(m, other)
            mask = m
        data = np.hstack(features)
    except ValueError:
        raise ValueError('all features must contain the same number of spikes')
    if feat_method_names:
        for i, method_name in enumerate(feat_method_names):
            for j, feature_name in enumerate(names[i]):
                names[i][j] = method_name + ':' + feature_name
    combined_features = {"data": data, "names": list(np.concatenate(names))}
    if list(mask):
        combined_features["is_valid"] = mask
    if norm:
        normalize(combined_features, copy=False)
    return combined_features
