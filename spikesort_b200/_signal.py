"""Device-resident recording that quacks like the array the reference keeps in
``sp_dict['data']``.

The reference's ``filter_proxy`` returns a PyTables CArray, i.e. NOT an ndarray but
something with ``.shape`` and slicing that yields ndarrays
(``src/spike_sort/core/filters.py:127-142``).  ``DeviceSignal`` plays the same role
with the samples in HBM: slicing copies the requested part to the host, while the
functions of this package read the CUDA tensor directly (``device_tensor()``).
"""
import numpy as np

from . import _lib


class DeviceSignal(object):
    def __init__(self, tensor):
        if tensor.dim() != 2:
            raise ValueError("DeviceSignal wants a (n_contacts, n_samples) tensor")
        self._t = tensor

    def device_tensor(self):
        return self._t

    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def dtype(self):
        return np.dtype(_lib.np_dtype_of(self._t))

    @property
    def ndim(self):
        return 2

    def __len__(self):
        return self._t.shape[0]

    def _conv(self, k):
        if isinstance(k, np.ndarray):
            return _lib.torch().as_tensor(k, device=self._t.device)
        if isinstance(k, np.integer):
            return int(k)
        return k

    def __getitem__(self, key):
        key = tuple(self._conv(k) for k in key) if isinstance(key, tuple) else self._conv(key)
        out = self._t[key]
        if hasattr(out, "cpu"):
            out = out.cpu().numpy()
        return out

    def __array__(self, dtype=None, copy=None):
        arr = self._t.cpu().numpy()
        return arr.astype(dtype) if dtype is not None else arr

    def __repr__(self):
        return "DeviceSignal(shape=%s, dtype=%s, device=%s)" % (self.shape, self.dtype,
                                                                self._t.device)


def to_device(spike_data):
    """Copy of a raw-recording dict whose 'data' lives on the GPU (one H2D copy);
    pass the result to the functions of this package to avoid re-uploading."""
    out = dict(spike_data)
    out['data'] = DeviceSignal(_lib.to_device_2d(spike_data['data']))
    return out
