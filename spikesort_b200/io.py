"""Raw-recording ingest for the front end.

``BakerlabFilter`` mirrors the raw-waveform half of the reference's
``spike_sort.io.filters.BakerlabFilter`` (``src/spike_sort/io/filters.py:21-152``): one
int16 file per contact, paths described by a JSON ``.inf`` file with ``{subject}``,
``{ses_id}``, ``{el_id}``, ``{contact_id}`` placeholders, datasets addressed as
``/{subject}/s{ses_id}/el{el_id}``.  ``read_sp`` returns the reference's raw-recording
dict ``{'data', 'FS', 'n_contacts'}``; here ``'data'`` is a :class:`DeviceSignal` holding
the ``(n_contacts, n_samples)`` int16 recording in HBM, ready for ``filter_proxy`` /
``sort_frontend`` without another upload.

The reference copies every contact file chunk by chunk into one host array
(``io/filters.py:117-128``).  Here the chunks go through two pinned staging buffers: while
chunk k travels to the GPU on a copy stream, chunk k+1 is read from the file, so disk
and PCIe overlap and no pageable copy is made.
"""
import json
import os
import re

import numpy as np

from . import _lib
from ._signal import DeviceSignal

_DATASET = ("^/(?P<subject>[a-zA-z]+)"
            "/s(?P<ses_id>.+)"
            "/el(?P<el_id>[0-9]+)"
            "/?(?P<type>[a-zA-Z]+)?(?P<cell_id>[0-9]+)?$")


class LazyRecording(object):
    """A raw recording that still lives in its per-contact files: shape/dtype like an array,
    `read_slab(a, b, out)` fills a (n_contacts, b-a) buffer with samples [a, b) of every
    contact.  `sort_frontend` pulls time slabs out of it while earlier slabs are already on
    the GPU (files -> pinned buffers -> HBM -> chain, all overlapped); anything else that
    needs the whole array (`np.asarray`, slicing) reads it."""

    dtype = np.dtype(np.int16)
    ndim = 2

    def __init__(self, names, npts):
        self.names = list(names)
        self.shape = (len(self.names), int(npts))
        self._lens = [min(os.path.getsize(n) // 2, int(npts)) for n in self.names]

    def read_slab(self, a, b, out):
        """out: int16 (n_contacts, b-a) torch CPU tensor or ndarray view, rows unit-stride."""
        for i, name in enumerate(self.names):
            row = out[i]
            arr = row.numpy() if hasattr(row, "numpy") else row
            hi = min(b, self._lens[i])
            if hi < b:
                arr[max(hi - a, 0):] = 0                    # shorter file: zeros, as read_sp
            if hi > a:
                with open(name, 'rb', buffering=0) as fid:
                    fid.seek(2 * a)
                    view = memoryview(arr[:hi - a]).cast('B')
                    if fid.readinto(view) != 2 * (hi - a):
                        raise IOError("short read from %s" % name)

    def __array__(self, dtype=None, copy=None):
        out = np.empty(self.shape, dtype=np.int16)
        self.read_slab(0, self.shape[1], out)
        return out.astype(dtype) if dtype is not None else out

    def __getitem__(self, key):
        return np.asarray(self)[key]

    def __len__(self):
        return self.shape[0]


class BakerlabFilter(object):
    """Reader/writer of raw recordings in the Bakerlab layout (int16 file per contact)."""

    def __init__(self, conf_file, chunksize=16 * 1024 * 1024):
        self.conf_file = conf_file
        with open(conf_file) as fid:
            self.conf_dict = json.load(fid)
        self.chunksize = int(chunksize)          # elements per staged chunk (32 MB of int16)

    # ------------------------------------------------------------------ paths
    def _match_dataset(self, dataset):
        m = re.match(_DATASET, dataset)
        if not m:
            raise ValueError("dataset id could not be parsed")
        return m.groupdict()

    def contact_files(self, dataset, n_contacts=None):
        """Paths of the per-contact files of `dataset` (``io/filters.py:95-98,114-115``)."""
        conf = self.conf_dict
        rec = self._match_dataset(dataset)
        dirname = os.path.expandvars(conf['dirname'])
        full_path = os.path.join(dirname, conf['fspike'])
        n = conf['n_contacts'] if n_contacts is None else n_contacts
        names = []
        for i in range(n):
            rec['contact_id'] = i + 1
            names.append(full_path.format(**rec))
        return names

    # ------------------------------------------------------------------- read
    def read_sp(self, dataset, memmap=None, host=False, lazy=False):
        """``io/filters.py:72-129``.  `memmap` is accepted for compatibility (the recording
        goes to HBM, not to a host array).  host=True returns the plain NumPy array the
        reference returns instead (no GPU involved: plain file reads).  lazy=True reads
        nothing yet: 'data' is a LazyRecording that `sort_frontend` streams slab by slab
        while the chain already runs on the earlier slabs."""
        conf = self.conf_dict
        names = self.contact_files(dataset)
        n_contacts = len(names)
        npts = os.path.getsize(names[0]) // 2
        if lazy:
            return {'data': LazyRecording(names, npts), 'FS': conf['FS'], 'n_contacts': n_contacts}
        if host:
            data = np.empty((n_contacts, npts), dtype=np.int16)
            for i, name in enumerate(names):
                sp = np.memmap(name, dtype=np.int16, mode='r')
                k = min(sp.shape[0], npts)
                data[i, :k] = sp[:k]
                del sp
            return {'data': data, 'FS': conf['FS'], 'n_contacts': n_contacts}
        t = _lib.require_cuda()
        dev = t.empty((n_contacts, npts), dtype=t.int16, device=_lib.device())
        step = max(1, min(self.chunksize, npts))
        stage = [t.empty(step, dtype=t.int16).pin_memory() for _ in range(2)]
        views = [memoryview(s.numpy()).cast('B') for s in stage]
        done = [None, None]
        copy_stream = t.cuda.Stream()
        copy_stream.wait_stream(t.cuda.current_stream())
        k = 0
        for i, name in enumerate(names):
            _npts = min(os.path.getsize(name) // 2, npts)
            if _npts < npts:                                # shorter file: rest stays as allocated
                dev[i, _npts:].zero_()
            with open(name, 'rb', buffering=0) as fid:
                for lo in range(0, _npts, step):
                    hi = min(lo + step, _npts)
                    slot = k & 1
                    if done[slot] is not None:
                        done[slot].synchronize()            # staging buffer free again
                    got = fid.readinto(views[slot][:2 * (hi - lo)])
                    if got != 2 * (hi - lo):
                        raise IOError("short read from %s" % name)
                    with t.cuda.stream(copy_stream):
                        dev[i, lo:hi].copy_(stage[slot][:hi - lo], non_blocking=True)
                        done[slot] = t.cuda.Event()
                        done[slot].record(copy_stream)
                    k += 1
        t.cuda.current_stream().wait_stream(copy_stream)
        for ev in done:
            if ev is not None:
                ev.synchronize()
        return {'data': DeviceSignal(dev), 'FS': conf['FS'], 'n_contacts': n_contacts}

    # ------------------------------------------------------------------ write
    def write_sp(self, sp_dict, dataset):
        """``io/filters.py:131-152``: one int16 file per contact (values cast as NumPy's
        ``astype(np.int16)`` does)."""
        sp = sp_dict['data']
        n_contacts = sp.shape[0]
        for i, name in enumerate(self.contact_files(dataset, n_contacts)):
            row = sp[i, :]                                   # DeviceSignal slices come back as NumPy
            np.asarray(row).astype(np.int16).tofile(name)
