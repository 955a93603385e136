"""Drop-in for the hot-path functions of ``spike_sort.core.extract`` (reference
``src/spike_sort/core/extract.py``): ``detect_spikes``, ``filter_spt``,
``extract_spikes``, ``resample_spikes``, ``align_spikes``, ``remove_doubles``.  Same arguments, same
dict keys and dtypes; spike times and waveforms are bit-identical to the
reference given the same input signal.
"""
from warnings import warn

import numpy as np

from . import _lib, _ops, _spline

_EDGES = ('rising', 'max', 'falling', 'min')


def _rows_on_device(spike_data, rows):
    """CUDA tensor holding the requested contacts of spike_data['data'] plus the
    mapping from contact id to tensor row.  Device-resident recordings are used as
    they are; host arrays are uploaded row-wise (only what the call reads)."""
    data = spike_data['data']
    t = _lib.require_cuda()
    if hasattr(data, "device_tensor") or isinstance(data, t.Tensor):
        return _lib.to_device_2d(data), None
    rows = [int(r) for r in rows]
    sub = np.stack([np.asarray(data[r, :]) for r in rows])
    return _lib.to_device_2d(sub), {r: i for i, r in enumerate(rows)}


def _to_host(ten):
    """Big blocks through a pooled pinned buffer (pipeline.to_host), small ones plainly."""
    from .pipeline import to_host
    return to_host(ten)


def _n_pts(spike_data):
    sp_data = spike_data['data']
    try:
        return sp_data.shape[1]
    except IndexError:
        return len(sp_data)


def detect_spikes(spike_data, thresh='auto', edge="rising", contact=0):
    """``extract.py:43-94``: amplitude-threshold detection on one contact.
    Returns {'data': float64 ms, 'thresh': threshold used, 'contact': contact}."""
    FS = spike_data['FS']
    x, remap = _rows_on_device(spike_data, [contact])
    row = x[(remap[int(contact)] if remap else contact):][:1]
    falling = edge in _EDGES[2:]
    if isinstance(thresh, str):
        frac = 8.0 if thresh == 'auto' else float(thresh)
        thr_dev = _ops.var_threshold(row, int(10 * FS), -frac if falling else frac)
        thresh = np.float64(thr_dev.item())
    else:
        t = _lib.torch()
        # NumPy compares in result_type(row, thresh): a Python float against a
        # float32 row is rounded to float32 first (weak scalar promotion)
        cmp_dtype = np.result_type(_lib.np_dtype_of(row), thresh)
        thr_val = float(np.asarray(thresh, dtype=cmp_dtype)) if cmp_dtype.kind == 'f' \
            else float(thresh)
        thr_dev = t.tensor([thr_val], dtype=t.float64, device=row.device)
    if edge not in _EDGES:
        raise TypeError("'edge' parameter must be 'rising' or 'falling'")
    spt, _ = _ops.detect(row, thr_dev, falling, FS)
    return {'data': spt.cpu().numpy(), 'thresh': thresh, 'contact': contact}


def filter_spt(spike_data, spt_dict, sp_win):
    """``extract.py:97-111``: indices of the spikes whose window fits the recording."""
    spt = np.asarray(spt_dict['data'])
    t_min, t_max = _ops.time_bounds(_n_pts(spike_data), spike_data['FS'], sp_win)
    idx, = np.nonzero((spt >= t_min) & (spt <= t_max))
    return idx


def extract_spikes(spike_data, spt_dict, sp_win, resample=1, contacts='all'):
    """``extract.py:114-184``: float32 waveforms (n_win, n_spikes, n_contacts),
    'time' (ms) and 'FS'; 'is_valid' only when some window was clipped."""
    n_contacts = spike_data['n_contacts']
    if isinstance(contacts, str) and contacts == "all":
        contacts = np.arange(n_contacts)
    elif isinstance(contacts, (int, np.integer)):
        contacts = np.array([contacts])
    else:
        contacts = np.asarray(contacts)
    FS = spike_data['FS']
    t = _lib.require_cuda()
    x, remap = _rows_on_device(spike_data, np.unique(contacts))
    rows = contacts if remap is None else np.array([remap[int(c)] for c in contacts])
    rows = np.where(rows < 0, rows + x.shape[0], rows) if remap is None else rows
    if len(rows) and (rows.min() < 0 or rows.max() >= x.shape[0]):
        raise IndexError("contact index out of range")
    spt = t.as_tensor(np.ascontiguousarray(spt_dict['data'], dtype=np.float64)).to(x.device)
    cdev = t.as_tensor(np.ascontiguousarray(rows, dtype=np.int32)).to(x.device)
    _, _, time = _ops.window_params(sp_win, FS)
    waves, valid, n_invalid = _ops.extract(x, spt, FS, sp_win, contacts=cdev)
    if resample != 1:
        warn("resample argument is deprecated."
             "Please update your code to use function"
             "resample_spikes", DeprecationWarning)
        # extract.py:183 returns resample_spikes' fresh dict: no 'is_valid'
        plan = _spline.spline_plan(time, FS * resample)
        return {"data": _to_host(_ops.resample(waves, plan)), "time": plan.new_time.copy(),
                "FS": FS}
    wavedict = {"data": _to_host(waves), "time": time, "FS": FS}
    if int(n_invalid.item()) != 0:
        wavedict['is_valid'] = valid.cpu().numpy().astype(bool)
    return wavedict


def resample_spikes(spikes_dict, FS_new):
    """``extract.py:187-205``: cubic-spline upsampling of every (spike, contact)
    waveform to `FS_new`; float64 'data', new 'time', 'FS' unchanged."""
    t = _lib.require_cuda()
    sp_waves = np.asarray(spikes_dict['data'])
    if sp_waves.dtype not in (np.int16, np.float32, np.float64):
        sp_waves = sp_waves.astype(np.float64)
    time = np.asarray(spikes_dict['time'], dtype=np.float64)
    plan = _spline.spline_plan(time, FS_new)
    waves = t.as_tensor(np.ascontiguousarray(sp_waves)).to(_lib.device())
    return {"data": _to_host(_ops.resample(waves, plan)), "time": plan.new_time.copy(),
            "FS": spikes_dict['FS']}


def remove_doubles(spt_dict, tol):
    """``extract.py:277-288``."""
    t = _lib.require_cuda()
    new_dict = spt_dict.copy()
    spt = np.ascontiguousarray(spt_dict['data'], dtype=np.float64)
    if len(spt) > 0:
        kept, _ = _ops.remove_doubles(t.as_tensor(spt).to(_lib.device()), None, tol)
        spt = kept.cpu().numpy()
    new_dict['data'] = spt
    return new_dict


def align_spikes(spike_data, spt_dict, sp_win, type="max", resample=1,
                 contact=0, remove=True):
    """``extract.py:208-274``: moves every spike to the extremum of its window on
    `contact` (repeating while the extremum sits at a window edge), then removes
    doubles closer than one sample."""
    tol = 0.1
    if (sp_win[0] > -tol) or (sp_win[1] < tol):
        warn('You are using very short sp_win. '
             'This may lead to alignment problems.')
    if type not in ("max", "min"):
        raise ValueError("type must be 'max' or 'min'")
    t = _lib.require_cuda()
    FS = spike_data['FS']
    x, remap = _rows_on_device(spike_data, [contact])
    row = x[(remap[int(contact)] if remap else contact):][:1]
    spt = t.as_tensor(np.array(spt_dict['data'], dtype=np.float64, copy=True)).to(row.device)
    if resample != 1:
        _ops.align_resampled(row, spt, None, None, FS, sp_win, type == "min", resample)
    else:
        _ops.align(row, spt, None, None, FS, sp_win, type == "min")
    if remove:
        spt, _ = _ops.remove_doubles(spt, None, 1000.0 / FS) if spt.numel() > 0 else (spt, None)
    return {'data': spt.cpu().numpy()}
