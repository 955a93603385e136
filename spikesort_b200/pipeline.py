"""The multi-contact front-end chain, device-resident and batched over contacts.

The reference processes one contact per call (``detect_spikes(contact=c)``,
``extract.py:43-44``); a 32..384-channel recording is therefore a Python loop over
contacts (SURVEY.md section 8d "multi-contact semantics"):

    for c in contacts:
        spt   = detect_spikes(filtered, contact=c, thresh='auto', edge=edge)
        spt   = align_spikes(filtered, spt, sp_win, type=align_type, contact=c)
        waves = extract_spikes(filtered, spt, sp_win, contacts=c)
        feats = fetPCA(waves, ncomps)

``run_chain`` computes exactly that for all contacts at once: every stage is one
batched launch sequence over (contact, ...) with per-contact results concatenated
and delimited by an offsets array, and nothing leaves HBM between the stages.
"""
import numpy as np

from . import _lib, _ops

_FALLING = ('falling', 'min')


class ChainResult(object):
    """Device-resident result of run_chain (CUDA tensors)."""

    def __init__(self):
        self.filtered = None        # float64 (n_contacts, n_samples) or the input if unfiltered
        self.thresh = None          # float64 (n_contacts,)
        self.spt_detected = None    # float64, concatenated per contact
        self.offsets_detected = None
        self.spt = None             # aligned, doubles removed
        self.offsets = None         # int64 (n_contacts+1,)
        self.waves = None           # float32 (n_win, n_spk, 1)
        self.valid = None           # uint8 (n_spk,)
        self.n_invalid = None
        self.evals = None           # float64 (n_contacts, n_win)
        self.evecs = None           # float64 (n_contacts, n_win, n_win)
        self.scores = None          # float64 (n_spk, ncomps)
        self.time = None            # host float64 (n_win,)

    def per_contact(self, c):
        """Host copies shaped like the reference's per-contact results."""
        off = self.offsets.cpu().numpy()
        offd = self.offsets_detected.cpu().numpy()
        lo, hi = int(off[c]), int(off[c + 1])
        out = {
            'spt_detected': self.spt_detected[int(offd[c]):int(offd[c + 1])].cpu().numpy(),
            'thresh': float(self.thresh[c].item()),
            'spt': self.spt[lo:hi].cpu().numpy(),
            'waves': np.ascontiguousarray(self.waves[:, lo:hi, :].cpu().numpy()),
            'is_valid': self.valid[lo:hi].cpu().numpy().astype(bool),
        }
        if self.scores is not None:
            out['scores'] = self.scores[lo:hi].cpu().numpy()
            out['evals'] = self.evals[c].cpu().numpy()
            out['evecs'] = self.evecs[c].cpu().numpy()
        return out


def run_chain(x, FS, filt=None, sp_win=(-0.2, 0.8), edge='falling', align_type='min',
              thresh='auto', ncomps=2, chunksize=1E6, pca=True, thr_override=None, fused=True):
    """x: CUDA tensor (n_contacts, n_samples), int16/float32/float64.
    filt: a spikesort_b200.filters filter object or None.
    thresh: 'auto' / numeric string (multiple of the std of the first 10 s) or a number.
    thr_override: optional CUDA float64 (n_contacts,) thresholds (e.g. broadcast from the
    rank owning the first 10 s when the recording is time-sharded).
    fused: mark the crossings in the filter's second sweep (default) instead of a
    separate pass over the filtered signal; the results are identical."""
    t = _lib.require_cuda()
    res = ChainResult()
    sp_win = list(sp_win)
    falling = edge in _FALLING
    auto = thr_override is None and isinstance(thresh, str)
    frac = (8.0 if thresh == 'auto' else float(thresh)) if auto else None
    if filt is not None and fused and hasattr(filt, "sections"):
        # filter with the threshold and the crossing marks produced in the same sweep
        sos, padlen = filt.sections(FS)
        if auto:
            thr = None
        elif thr_override is not None:
            thr = thr_override
        else:
            thr = t.full((x.shape[0],), float(thresh), dtype=t.float64, device=x.device)
        sig, thr, spt, offsets = _ops.filtfilt_detect(
            x, sos, padlen, int(chunksize), FS, falling, thr=thr, n0=int(10 * FS),
            factor=(-frac if falling else frac) if auto else None)
        res.filtered, res.thresh = sig, thr
    else:
        if filt is not None:
            sig = filt.apply(x, FS, int(chunksize))      # unfused IIR, or the FIR stencil
        else:
            sig = x
        res.filtered = sig
        n_rows = sig.shape[0]
        if thr_override is not None:
            thr = thr_override
        elif auto:
            thr = _ops.var_threshold(sig, int(10 * FS), -frac if falling else frac)
        else:
            thr = t.full((n_rows,), float(thresh), dtype=t.float64, device=sig.device)
        res.thresh = thr
        spt, offsets = _ops.detect(sig, thr, falling, FS)
    res.spt_detected, res.offsets_detected = spt, offsets
    spt = spt.clone()
    _ops.align(sig, spt, offsets, None, FS, sp_win, align_type == 'min')
    spt, offsets = _ops.remove_doubles(spt, offsets, 1000.0 / FS)
    res.spt, res.offsets = spt, offsets
    waves, valid, n_invalid = _ops.extract(sig, spt, FS, sp_win, contacts=None, offsets=offsets)
    res.waves, res.valid, res.n_invalid = waves, valid, n_invalid
    res.time = _ops.window_params(sp_win, FS)[2]
    if pca and spt.numel() > 0:
        gram, ssum, _, counts = _ops.pca_gram(waves, offsets=offsets)
        res.evals, res.evecs = _ops.pca_eig(gram, ssum, counts)
        res.scores = _ops.pca_project(waves, res.evals, res.evecs, ncomps, offsets=offsets)
    return res


_PINNED = {}


def _pinned_like(name, tensor):
    """Grow-only pinned staging buffer (pinned allocation costs milliseconds)."""
    t = _lib.torch()
    need = tensor.numel()
    buf = _PINNED.get((name, tensor.dtype))
    if buf is None or buf.numel() < need:
        buf = t.empty(max(need, 1), dtype=tensor.dtype).pin_memory()
        _PINNED[(name, tensor.dtype)] = buf
    return buf[:need].view(tensor.shape)


def sort_frontend(spike_data, filt=None, sp_win=(-0.2, 0.8), edge='falling', align_type='min',
                  thresh='auto', ncomps=2, chunksize=1E6, contacts=None):
    """The whole front end for a multi-contact recording in ONE call, host in / host out.

    spike_data: the reference's raw-recording dict {'data', 'FS', 'n_contacts'}; 'data' may
    be a NumPy array, a (pinned) torch CPU tensor or a device-resident signal.  Every
    contact c gets exactly what the reference's per-contact calls would give
    (detect_spikes(contact=c) -> align_spikes(contact=c) -> extract_spikes(contacts=c) ->
    fetPCA), computed by the batched device chain; results come back as NumPy in the
    reference's dict conventions:

        {'FS', 'n_contacts', 'contacts': [c, ...],
         'spt':      [ {'data', 'thresh', 'contact'}, ... ],          # detected
         'spt_aligned': [ {'data'}, ... ],
         'sp_waves': [ {'data' (n_win, n_spk_c, 1) float32, 'time', 'FS'[, 'is_valid']}, ... ],
         'features': [ {'data' (n_spk_c, ncomps), 'names'[, 'is_valid']}, ... ]}
    """
    t = _lib.require_cuda()
    FS = spike_data['FS']
    x = _lib.to_device_2d(spike_data['data'])            # one H2D copy (pinned: ~55 GB/s)
    if contacts is not None:
        x = x[t.as_tensor(list(contacts), device=x.device)]
    res = run_chain(x, FS, filt=filt, sp_win=sp_win, edge=edge, align_type=align_type,
                    thresh=thresh, ncomps=ncomps, chunksize=chunksize)
    # one D2H per result array, through pinned staging, then zero-copy per-contact views
    def host(name, ten):
        if ten is None:
            return None
        buf = _pinned_like(name, ten)
        buf.copy_(ten, non_blocking=True)
        return buf
    h = {k: host(k, getattr(res, k)) for k in ('thresh', 'spt_detected', 'offsets_detected', 'spt',
                                               'offsets', 'waves', 'valid', 'scores')}
    t.cuda.current_stream().synchronize()
    h = {k: (v.numpy() if v is not None else None) for k, v in h.items()}
    n_rows = x.shape[0]
    ids = list(range(n_rows)) if contacts is None else list(contacts)
    out = {'FS': FS, 'n_contacts': n_rows, 'contacts': ids, 'spt': [], 'spt_aligned': [],
           'sp_waves': [], 'features': []}
    names = ["Ch0:PC%d" % j for j in range(ncomps)]
    off, offd = h['offsets'], h['offsets_detected']
    for r, c in enumerate(ids):
        lo, hi = int(off[r]), int(off[r + 1])
        out['spt'].append({'data': h['spt_detected'][int(offd[r]):int(offd[r + 1])].copy(),
                           'thresh': float(h['thresh'][r]), 'contact': c})
        out['spt_aligned'].append({'data': h['spt'][lo:hi].copy()})
        valid = h['valid'][lo:hi].astype(bool)
        wd = {'data': np.ascontiguousarray(h['waves'][:, lo:hi, :]), 'time': res.time, 'FS': FS}
        fd = {'data': h['scores'][lo:hi].copy() if h['scores'] is not None else None,
              'names': list(names)}
        if not valid.all():
            wd['is_valid'] = valid
            fd['is_valid'] = valid
        out['sp_waves'].append(wd)
        out['features'].append(fd)
    return out
