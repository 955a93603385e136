"""Device-level wrappers of the C ABI: CUDA tensors in, CUDA tensors out.

Everything here only allocates buffers (torch) and calls ``libspikesort_b200.so``;
host<->device synchronisation happens only where a size must be known to allocate
the result (number of detected / kept spikes), exactly as documented in the header.
The exact float64 host expressions the reference evaluates with NumPy/Python
(window, time axis, bounds) are computed here the same way.
"""
import ctypes

import numpy as np

from . import _lib

MAX_ALIGN_ITER = 100000      # the reference loops until no spike moves; a spike
                             # advances by >= 1 sample per pass, so this is a safety net


# host-side ESTIMATE of the kernels launched through this module (kept for quick looks; the
# authoritative count, which bench.py reports, is the library's own ss_launch_count())
LAUNCHES = {"count": 0}


def launch_count():
    """Kernels of libspikesort_b200.so launched by this process so far (counted in the library
    at every launch site)."""
    return int(_lib.load().ss_launch_count())


def _launched(n):
    LAUNCHES["count"] += n


class Shard(object):
    """Position of a time shard in the whole recording (include/spikesort_b200.h,
    "time shards"): global index of its first sample, total length of the recording and
    the number of neighbour samples stored before / after it in the same buffer."""

    def __init__(self, index_offset=0, n_total=None, halo_before=0, halo_after=0):
        self.index_offset = int(index_offset)
        self.n_total = n_total
        self.halo_before = int(halo_before)
        self.halo_after = int(halo_after)


_WHOLE = Shard()


# ------------------------------------------------------------------ host math
def window_params(sp_win, FS):
    """win (int32, truncating), n_win and the float64 time axis, exactly as
    extract.py:151-152."""
    win = (np.asarray(sp_win) / 1000.0 * FS).astype(np.int32)
    time = np.arange(win[1] - win[0]) * 1000.0 / FS + sp_win[0]
    return int(win[0]), int(win[1]), time


def time_bounds(n_pts, FS, sp_win):
    """Inclusive [t_min, t_max] of filter_spt, extract.py:106-109."""
    max_time = n_pts * 1000.0 / FS
    t_min = np.max((-sp_win[0], 0))
    t_max = np.min((max_time, max_time - sp_win[1]))
    return float(t_min), float(t_max)


# --------------------------------------------------------------------- filter
def filtfilt_sos(x, sos, padlen, chunksize, out=None):
    """x: CUDA (n_contacts, n_samples) int16/float32/float64, unit time stride."""
    lib = _lib.load()
    t = _lib.torch()
    n_c, n = x.shape
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    n_sec = sos.shape[0]
    if out is None:
        out = t.empty((n_c, n), dtype=t.float64, device=x.device)
    nbytes = lib.ss_filtfilt_workspace_bytes(n_c, n, int(chunksize), int(padlen), n_sec)
    ws = _lib.workspace(nbytes)
    rc = lib.ss_filtfilt_sos(_lib.ptr(x), _lib.ss_dtype_of(x), n_c, n, x.stride(0),
                             sos.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n_sec,
                             int(padlen), int(chunksize), _lib.ptr(out), out.stride(0),
                             _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "ss_filtfilt_sos")
    _launched(3)          # summary, tilescan, apply
    return out


def filtfilt_fir(x, b, chunksize, out=None):
    """Zero-phase FIR filter (taps b, host float64) of every (contact, chunk) unit."""
    lib = _lib.load()
    t = _lib.torch()
    n_c, n = x.shape
    b = np.ascontiguousarray(b, dtype=np.float64)
    bd = t.from_numpy(b).to(x.device)
    if out is None:
        out = t.empty((n_c, n), dtype=t.float64, device=x.device)
    rc = lib.ss_filtfilt_fir(_lib.ptr(x), _lib.ss_dtype_of(x), n_c, n, x.stride(0), _lib.ptr(bd),
                             int(b.shape[0]), int(chunksize), _lib.ptr(out), out.stride(0),
                             _lib.stream_ptr())
    _lib.check(rc, "ss_filtfilt_fir")
    _launched(1)
    return out


def filtfilt_detect(x, sos, padlen, chunksize, FS, falling, thr=None, n0=None, factor=None,
                    out=None, ws=None, shard=_WHOLE, cap=None, peer=None):
    """Fused filter + (auto threshold) + crossing marking + emit.
    thr: CUDA float64 (n_contacts,) thresholds, or None to derive them from the first
    n0 filtered samples (factor carries the sign).  Returns (out, thr, spt, offsets).
    cap: size the spike list by this capacity instead of synchronising with the host to
    learn the count; the true count is offsets[-1] (more than cap: the list is truncated and
    the caller must redo the call without cap).
    peer: (left_from_right_ptr, right_from_left_ptr, H) - integer device addresses in the
    neighbour shards' arenas (0 = no neighbour): the tiles of sweep 2 that produce the first /
    last H samples store them there as well (halo exchange fused into the filter)."""
    lib = _lib.load()
    t = _lib.torch()
    n_c, n = x.shape
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    n_sec = sos.shape[0]
    if out is None:
        out = t.empty((n_c, n), dtype=t.float64, device=x.device)
    prepared = ws is not None                       # sweep 1 already done by filtfilt_prepare
    if ws is None:
        ws = _lib.workspace(lib.ss_filtfilt_workspace_bytes(n_c, n, int(chunksize), int(padlen),
                                                            n_sec))
    dws = _lib.workspace(lib.ss_detect_workspace_bytes(n_c, n))
    offsets = t.empty(n_c + 1, dtype=t.int64, device=x.device)
    auto = thr is None
    if auto:
        thr = t.empty(n_c, dtype=t.float64, device=x.device)
    args = (_lib.ptr(x), _lib.ss_dtype_of(x), n_c, n, x.stride(0),
            sos.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n_sec,
            int(padlen), int(chunksize), _lib.ptr(out), out.stride(0),
            _lib.ptr(ws), ws.numel(), int(auto),
            int(n0) if auto else 0, float(factor) if auto else 0.0,
            _lib.ptr(thr), int(bool(falling)), _lib.ptr(dws), dws.numel(),
            _lib.ptr(offsets), int(prepared))
    if peer is None:
        rc = lib.ss_filtfilt_detect_sos(*(args + (_lib.stream_ptr(),)))
    else:
        rc = lib.ss_filtfilt_detect_sos_peer(*(args + (ctypes.c_void_p(int(peer[0])),
                                                       ctypes.c_void_p(int(peer[1])), int(peer[2]),
                                                       _lib.stream_ptr())))
    _lib.check(rc, "ss_filtfilt_detect_sos")
    # summary, tilescan | apply(head), 2 threshold, mark(head), edge | apply+marks, edge | count, scan
    _launched((0 if prepared else 2) + (5 if auto else 0) + 4)
    total = int(offsets[-1].item()) if cap is None else int(cap)
    spt = t.empty(total, dtype=t.float64, device=x.device)
    rc = lib.ss_detect_emit(n_c, n, float(FS), shard.index_offset, _lib.ptr(dws), dws.numel(),
                            _lib.ptr(spt), total, _lib.stream_ptr())
    _lib.check(rc, "ss_detect_emit")
    _launched(1 if total > 0 else 0)
    return out, thr, spt, offsets


def filtfilt_prepare(x, sos, padlen, chunksize):
    """Sweep 1 + tile scan of every chunk (no threshold needed).  Returns the workspace
    to hand to filtfilt_head_threshold / filtfilt_detect."""
    lib = _lib.load()
    n_c, n = x.shape
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    n_sec = sos.shape[0]
    ws = _lib.workspace(lib.ss_filtfilt_workspace_bytes(n_c, n, int(chunksize), int(padlen), n_sec))
    rc = lib.ss_filtfilt_prepare_sos(_lib.ptr(x), _lib.ss_dtype_of(x), n_c, n, x.stride(0),
                                     sos.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n_sec,
                                     int(padlen), int(chunksize), _lib.ptr(ws), ws.numel(),
                                     _lib.stream_ptr())
    _lib.check(rc, "ss_filtfilt_prepare_sos")
    _launched(2)
    return ws


def filtfilt_head_threshold(x, sos, padlen, chunksize, out, ws, n0, factor):
    """With `ws` prepared: filter the chunks holding the first n0 samples into `out`
    and return factor*sqrt(var(out[:, :n0])) per contact."""
    lib = _lib.load()
    t = _lib.torch()
    n_c, n = x.shape
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    thr = t.empty(n_c, dtype=t.float64, device=x.device)
    tws = _lib.workspace(lib.ss_threshold_workspace_bytes(n_c))
    rc = lib.ss_filtfilt_head_threshold_sos(
        _lib.ptr(x), _lib.ss_dtype_of(x), n_c, n, x.stride(0),
        sos.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), sos.shape[0], int(padlen),
        int(chunksize), _lib.ptr(out), out.stride(0), _lib.ptr(ws), ws.numel(), int(n0),
        float(factor), _lib.ptr(thr), _lib.ptr(tws), tws.numel(), _lib.stream_ptr())
    _lib.check(rc, "ss_filtfilt_head_threshold_sos")
    _launched(3)
    return thr


# ------------------------------------------------------------------ threshold
def var_threshold(x, n0, factor):
    """factor * sqrt(var(x[r, :n0])) for every row of the CUDA tensor x -> CUDA float64."""
    lib = _lib.load()
    t = _lib.torch()
    n_rows = x.shape[0]
    n0 = min(int(n0), x.shape[1])
    thr = t.empty(n_rows, dtype=t.float64, device=x.device)
    ws = _lib.workspace(lib.ss_threshold_workspace_bytes(n_rows))
    rc = lib.ss_var_threshold(_lib.ptr(x), _lib.ss_dtype_of(x), n_rows, x.stride(0), n0,
                              float(factor), _lib.ptr(thr), _lib.ptr(ws), ws.numel(),
                              _lib.stream_ptr())
    _lib.check(rc, "ss_var_threshold")
    _launched(2)
    return thr


# --------------------------------------------------------------------- detect
def detect(x, thr, falling, FS, shard=_WHOLE, cap=None):
    """Threshold crossings of every row.  Returns (spt, offsets): CUDA float64 times
    (ms) concatenated row by row and CUDA int64 offsets[n_rows+1]."""
    lib = _lib.load()
    t = _lib.torch()
    n_rows, n = x.shape
    ws = _lib.workspace(lib.ss_detect_workspace_bytes(n_rows, n))
    offsets = t.empty(n_rows + 1, dtype=t.int64, device=x.device)
    rc = lib.ss_detect_mark(_lib.ptr(x), _lib.ss_dtype_of(x), n_rows, n, x.stride(0),
                            _lib.ptr(thr), int(bool(falling)), _lib.ptr(ws), ws.numel(),
                            _lib.ptr(offsets), _lib.stream_ptr())
    _lib.check(rc, "ss_detect_mark")
    _launched(2)          # mark, scan
    total = int(offsets[-1].item()) if cap is None else int(cap)   # the one sync: size of the result
    spt = t.empty(total, dtype=t.float64, device=x.device)
    rc = lib.ss_detect_emit(n_rows, n, float(FS), shard.index_offset, _lib.ptr(ws), ws.numel(),
                            _lib.ptr(spt), total, _lib.stream_ptr())
    _lib.check(rc, "ss_detect_emit")
    _launched(1 if total > 0 else 0)
    return spt, offsets


# ---------------------------------------------------------------------- align
def align(x, spt, offsets, rows, FS, sp_win, use_min, shard=_WHOLE, halo_miss=None):
    """In-place alignment of CUDA float64 times `spt`.  offsets: CUDA int64
    [n_rows+1] or None (single row); rows: CUDA int32 row ids or None.
    halo_miss: CUDA int32[1] counter of spikes whose window left the shard's halo."""
    lib = _lib.load()
    n = x.shape[1]
    win0, win1, _ = window_params(sp_win, FS)
    t_min, t_max = time_bounds(n if shard.n_total is None else shard.n_total, FS, sp_win)
    tol = 0.1                                       # extract.py:229
    lo, hi = sp_win[0] + tol, sp_win[1] - tol       # extract.py:263-264
    n_rows = 1 if offsets is None else offsets.numel() - 1
    rc = lib.ss_align(_lib.ptr(x), _lib.ss_dtype_of(x), n, x.stride(0), float(FS),
                      _lib.ptr(rows), n_rows, _lib.ptr(offsets), float(sp_win[0]), win0, win1,
                      t_min, t_max, float(lo), float(hi), int(bool(use_min)), _lib.ptr(spt),
                      spt.numel(), MAX_ALIGN_ITER, shard.index_offset, shard.halo_before,
                      shard.halo_after, _lib.ptr(halo_miss), _lib.stream_ptr())
    _lib.check(rc, "ss_align")
    _launched(1 if spt.numel() > 0 else 0)
    return spt


def remove_doubles(spt, offsets, tol, prev_last=None, also=None, cap=None):
    """Returns (kept times, new offsets); offsets None = one row.  `spt` may have spare
    capacity beyond offsets[-1].  also: optional CUDA int64 scalar fetched in the same host
    synchronisation that sizes the result; its value is then returned as a third item."""
    lib = _lib.load()
    t = _lib.torch()
    n_spk = spt.numel()
    single = offsets is None
    if single:
        offsets = t.tensor([0, n_spk], dtype=t.int64, device=spt.device)
    n_rows = offsets.numel() - 1
    ws = _lib.workspace(lib.ss_remove_doubles_workspace_bytes(n_spk))
    new_off = t.empty(n_rows + 1, dtype=t.int64, device=spt.device)
    rc = lib.ss_remove_doubles_mark(_lib.ptr(spt), n_spk, _lib.ptr(offsets), n_rows, float(tol),
                                    _lib.ptr(prev_last), _lib.ptr(ws), ws.numel(), _lib.ptr(new_off),
                                    _lib.stream_ptr())
    _lib.check(rc, "ss_remove_doubles_mark")
    _launched(3)          # mark, scan, row offsets
    if cap is not None:
        total = int(cap)                            # no host sync; true count = new_off[-1]
    elif also is None:
        total = int(new_off[-1].item())
    else:
        total, extra = (int(v) for v in t.stack((new_off[-1], also.reshape(()))).tolist())
    out = t.empty(total, dtype=t.float64, device=spt.device)
    rc = lib.ss_remove_doubles_emit(_lib.ptr(spt), n_spk, _lib.ptr(ws), ws.numel(), _lib.ptr(out),
                                    total, _lib.stream_ptr())
    _lib.check(rc, "ss_remove_doubles_emit")
    _launched(1 if (n_spk > 0 and total > 0) else 0)
    if also is not None:
        return out, (None if single else new_off), extra
    return out, (None if single else new_off)


# -------------------------------------------------------------------- extract
def extract(x, spt, FS, sp_win, contacts=None, offsets=None, rows=None, shard=_WHOLE,
            halo_miss=None):
    """Returns (waves float32 (n_win, n_spk, n_c), valid uint8[n_spk], n_invalid int32[1])
    as CUDA tensors.  contacts: CUDA int32 list (reference mode) or None (row mode)."""
    lib = _lib.load()
    t = _lib.torch()
    n = x.shape[1]
    win0, win1, _ = window_params(sp_win, FS)
    t_min, t_max = time_bounds(n if shard.n_total is None else shard.n_total, FS, sp_win)
    n_spk = spt.numel()
    n_c = 1 if contacts is None else contacts.numel()
    waves = t.empty((win1 - win0, n_spk, n_c), dtype=t.float32, device=x.device)
    valid = t.empty(n_spk, dtype=t.uint8, device=x.device)
    n_invalid = t.empty(1, dtype=t.int32, device=x.device)
    n_rows = 1 if offsets is None else offsets.numel() - 1
    rc = lib.ss_extract(_lib.ptr(x), _lib.ss_dtype_of(x), n, x.stride(0), float(FS),
                        _lib.ptr(contacts), n_c, _lib.ptr(rows), n_rows, _lib.ptr(offsets),
                        win0, win1, t_min, t_max, _lib.ptr(spt), n_spk, _lib.ptr(waves),
                        _lib.ptr(valid), _lib.ptr(n_invalid), shard.index_offset,
                        shard.halo_before, shard.halo_after, _lib.ptr(halo_miss),
                        _lib.stream_ptr())
    _lib.check(rc, "ss_extract")
    _launched(1 if n_spk > 0 else 0)
    return waves, valid, n_invalid


# ------------------------------------------------------------------------ PCA
def pca_first_wave(waves, offsets=None):
    """Reference waveform (float64 (G, n_win)) for the Gram shift: the first spike of every
    group (contact c's first spike without offsets)."""
    lib = _lib.load()
    t = _lib.torch()
    n_win, n_spk, n_c = waves.shape
    n_groups = n_c if offsets is None else offsets.numel() - 1
    ref = t.empty((n_groups, n_win), dtype=t.float64, device=waves.device)
    rc = lib.ss_pca_first_wave(_lib.ptr(waves), _lib.ss_dtype_of(waves), n_win, n_spk, n_c,
                               _lib.ptr(offsets), n_groups,
                               _lib.ptr(ref), _lib.stream_ptr())
    _lib.check(rc, "ss_pca_first_wave")
    _launched(1)
    return ref


GRAM_MODES = {"auto": _lib.SS_GRAM_AUTO, "f64": _lib.SS_GRAM_F64, "tc": _lib.SS_GRAM_TC}


def pca_gram(waves, offsets=None, ref=None, mode="auto"):
    """Gram / sums of CUDA waves (n_win, n_spk, n_c) (float32 / float64 / int16) per group.
    Returns (gram (G, n_win, n_win), sum (G, n_win), ref (G, n_win), counts (G,)).
    mode: 'f64' exact float64 products, 'tc' 3xTF32 tensor cores (eligible float32 blocks),
    'auto' = 'tc' from 65536 spikes per group on - a function of the call's shape only."""
    lib = _lib.load()
    t = _lib.torch()
    n_win, n_spk, n_c = waves.shape
    if offsets is None:
        n_groups = n_c
        counts = t.full((n_groups,), n_spk, dtype=t.int64, device=waves.device)
    else:
        n_groups = offsets.numel() - 1
        counts = offsets[1:] - offsets[:-1]
    if ref is None:
        ref = pca_first_wave(waves, offsets)
    gram = t.empty((n_groups, n_win, n_win), dtype=t.float64, device=waves.device)
    ssum = t.empty((n_groups, n_win), dtype=t.float64, device=waves.device)
    ws = _lib.workspace(lib.ss_pca_workspace_bytes(n_win, n_groups))
    rc = lib.ss_pca_gram(_lib.ptr(waves), _lib.ss_dtype_of(waves), n_win, n_spk, n_c,
                         _lib.ptr(offsets), n_groups, _lib.ptr(ref), _lib.ptr(gram), _lib.ptr(ssum),
                         GRAM_MODES[mode], _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "ss_pca_gram")
    _launched(2)          # partial Grams, ordered reduction
    return gram, ssum, ref, counts


def pca_eig(gram, ssum, counts, top=None):
    """Eigenpairs of the per-group covariance.  top=None: all of them (Jacobi), what PCA()
    returns; top=k: only the k leading pairs (ss_pca_eig_top; the other columns are zero),
    all that the scores of fetPCA need."""
    lib = _lib.load()
    t = _lib.torch()
    n_groups, n_win, _ = gram.shape
    evals = t.empty((n_groups, n_win), dtype=t.float64, device=gram.device)
    evecs = t.empty((n_groups, n_win, n_win), dtype=t.float64, device=gram.device)
    if top is None:
        rc = lib.ss_pca_eig(_lib.ptr(gram), _lib.ptr(ssum), _lib.ptr(counts), n_win, n_groups,
                            _lib.ptr(evals), _lib.ptr(evecs), _lib.stream_ptr())
        _lib.check(rc, "ss_pca_eig")
    else:
        rc = lib.ss_pca_eig_top(_lib.ptr(gram), _lib.ptr(ssum), _lib.ptr(counts), n_win, n_groups,
                                int(top), _lib.ptr(evals), _lib.ptr(evecs), _lib.stream_ptr())
        _lib.check(rc, "ss_pca_eig_top")
    _launched(1)
    return evals, evecs


def pca_project(waves, evals, evecs, ncomps, offsets=None):
    lib = _lib.load()
    t = _lib.torch()
    n_win, n_spk, n_c = waves.shape
    n_groups = evals.shape[0]
    width = ncomps * (n_c if offsets is None else 1)
    scores = t.empty((n_spk, width), dtype=t.float64, device=waves.device)
    rc = lib.ss_pca_project(_lib.ptr(waves), _lib.ss_dtype_of(waves), n_win, n_spk, n_c,
                            _lib.ptr(offsets), n_groups,
                            _lib.ptr(evals), _lib.ptr(evecs), int(ncomps), _lib.ptr(scores),
                            _lib.stream_ptr())
    _lib.check(rc, "ss_pca_project")
    _launched(1 if n_spk > 0 else 0)
    return scores


# ------------------------------------------------------------- cheap features
def wave_minmax(waves, want_min=True, want_p2p=True):
    """(min, max - min) over the time axis of CUDA float32 / float64 waves
    (n_win, n_spk, n_c); each (n_spk, n_c) in the waves' dtype or None."""
    lib = _lib.load()
    t = _lib.torch()
    n_win, n_spk, n_c = waves.shape
    mn = t.empty((n_spk, n_c), dtype=waves.dtype, device=waves.device) if want_min else None
    pp = t.empty((n_spk, n_c), dtype=waves.dtype, device=waves.device) if want_p2p else None
    rc = lib.ss_wave_minmax(_lib.ptr(waves), _lib.ss_dtype_of(waves), n_win, n_spk, n_c,
                            _lib.ptr(mn), _lib.ptr(pp), _lib.stream_ptr())
    _lib.check(rc, "ss_wave_minmax")
    _launched(1 if n_spk > 0 else 0)
    return mn, pp


def normalize_columns(x):
    """In-place (x - colmin) / (colmax - colmin) of a CUDA float64 (n_rows, n_cols) matrix."""
    lib = _lib.load()
    n_rows, n_cols = x.shape
    ws = _lib.workspace(lib.ss_normalize_workspace_bytes(n_cols))
    rc = lib.ss_normalize_columns(_lib.ptr(x), n_rows, n_cols, _lib.ptr(ws), ws.numel(),
                                  _lib.stream_ptr())
    _lib.check(rc, "ss_normalize_columns")
    _launched(3 if n_rows > 0 and n_cols > 0 else 0)
    return x


# ----------------------------------------------------------- spline upsampling
def _plan_on_device(plan, device):
    t = _lib.torch()
    key = str(device)
    cache = plan.__dict__.setdefault("_dev", {})
    if key not in cache:
        cache[key] = (t.from_numpy(plan.f64).to(device), t.from_numpy(plan.i32).to(device),
                      t.from_numpy(plan.new_time).to(device))
    return cache[key]


def resample(waves, plan):
    """CUDA waves (n_win, n_spk, n_c), any ss dtype -> CUDA float64 (n_out, n_spk, n_c);
    `plan` = _spline.spline_plan(time, FS_new)."""
    lib = _lib.load()
    t = _lib.torch()
    n_win, n_spk, n_c = waves.shape
    if n_win != plan.m:
        raise ValueError("waveforms have %d samples, the spline plan %d" % (n_win, plan.m))
    pf, pi, _ = _plan_on_device(plan, waves.device)
    out = t.empty((plan.n_out, n_spk, n_c), dtype=t.float64, device=waves.device)
    rc = lib.ss_resample(_lib.ptr(waves), _lib.ss_dtype_of(waves), n_win, n_spk, n_c, _lib.ptr(pf),
                         _lib.ptr(pi), plan.n_out, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "ss_resample")
    _launched(1 if n_spk > 0 and n_c > 0 and plan.n_out > 0 else 0)
    return out


def align_resampled(x, spt, offsets, rows, FS, sp_win, use_min, resample_factor, shard=_WHOLE,
                    halo_miss=None):
    """align() with the extremum searched on the spline-upsampled window
    (extract.py:245-258 with resample != 1)."""
    from ._spline import spline_plan
    lib = _lib.load()
    n = x.shape[1]
    win0, win1, time = window_params(sp_win, FS)
    t_min, t_max = time_bounds(n if shard.n_total is None else shard.n_total, FS, sp_win)
    tol = 0.1
    lo, hi = sp_win[0] + tol, sp_win[1] - tol
    plan = spline_plan(time, FS * resample_factor)
    pf, pi, rt = _plan_on_device(plan, x.device)
    n_rows = 1 if offsets is None else offsets.numel() - 1
    rc = lib.ss_align_resampled(_lib.ptr(x), _lib.ss_dtype_of(x), n, x.stride(0), float(FS),
                                _lib.ptr(rows), n_rows, _lib.ptr(offsets), win0, win1, t_min, t_max,
                                float(lo), float(hi), int(bool(use_min)), _lib.ptr(pf), _lib.ptr(pi),
                                plan.n_out, _lib.ptr(rt), _lib.ptr(spt), spt.numel(), MAX_ALIGN_ITER,
                                shard.index_offset, shard.halo_before, shard.halo_after,
                                _lib.ptr(halo_miss), _lib.stream_ptr())
    _lib.check(rc, "ss_align_resampled")
    _launched(1 if spt.numel() > 0 else 0)
    return spt
