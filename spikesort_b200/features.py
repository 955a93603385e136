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


def _pca_device(w, ncomps):
    gram, ssum, _, counts = _ops.pca_gram(w)
    evals, evecs = _ops.pca_eig(gram, ssum, counts)
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
    _, _, score = _pca_device(w, ncomps)
    names = ["Ch%d:PC%d" % (i, j) for i in range(n_channels) for j in range(ncomps)]
    return {'data': score.cpu().numpy(), "names": names}
