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
import os

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
        self.evals = None           # float64 (n_contacts, n_win): the ncomps leading eigenvalues, then zeros
        self.evecs = None           # float64 (n_contacts, n_win, n_win): leading ncomps columns, rest zero
        self.scores = None          # float64 (n_spk, ncomps)
        self.time = None            # host float64 (n_win,)
        self._pending = None        # run_chain(lazy=True): (capacity, exact rerun, hint key)

    def finalize(self):
        """A result of run_chain(lazy=True) holds capacity-sized arrays whose true lengths are
        still on the device.  finalize() is the one host synchronisation of the chain: it reads
        the two counts, trims the arrays to views of their true length and, if the capacity
        was too small, replaces the result by the exact rerun.  Idempotent; returns self."""
        if self._pending is None:
            return self
        cap, rerun, key = self._pending
        self._pending = None
        n_det, _ = _trim_to_counts(self, cap)
        if n_det > cap:
            exact, n_det, _ = rerun()
            self.__dict__.update(exact.__dict__)
        _CAPACITY_HINTS[key] = n_det
        return self

    def per_contact(self, c):
        """Host copies shaped like the reference's per-contact results."""
        self.finalize()
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


_CAPACITY_HINTS = {}


def run_chain(x, FS, filt=None, sp_win=(-0.2, 0.8), edge='falling', align_type='min',
              thresh='auto', ncomps=2, chunksize=1E6, pca=True, thr_override=None, fused=True,
              speculate=True, resample=1, lazy=False):
    """x: CUDA tensor (n_contacts, n_samples), int16/float32/float64.
    filt: a spikesort_b200.filters filter object or None.
    thresh: 'auto' / numeric string (multiple of the std of the first 10 s) or a number.
    thr_override: optional CUDA float64 (n_contacts,) thresholds (e.g. broadcast from the
    rank owning the first 10 s when the recording is time-sharded).
    fused: mark the crossings in the filter's second sweep (default) instead of a
    separate pass over the filtered signal; the results are identical.
    speculate: the chain needs the number of detected / kept spikes on the HOST twice, to
    size arrays - two synchronisations that leave the GPU waiting for the launches of every
    kernel behind them.  When a call with the same shape has been seen before, the arrays
    are sized by that call's count + 25 % instead, every kernel bounds its work by the
    device-side offsets, and the counts are read once at the end; if the capacity turns out
    too small the chain is simply run again with exact sizes.  Results are identical.
    lazy: with a capacity-sized chain, do not even read the counts: the result comes back
    with everything queued and ChainResult.finalize() (called by per_contact(), or by the
    caller once the NEXT recording's chain has been queued) does the one synchronisation.
    Back-to-back calls then keep the GPU busy across calls - otherwise it idles ~0.1 ms per
    call while the host gets from the synchronisation to the first launch of the next one.
    resample: align on the cubic-spline upsampled window (align_spikes(resample=...),
    extract.py:245-258; the tutorial uses 10); waveforms are extracted at the original rate."""
    key = (tuple(x.shape), x.dtype, float(FS), tuple(sp_win), edge, align_type, str(thresh),
           id(type(filt)), int(chunksize), bool(fused), int(resample))
    args = (x, FS, filt, sp_win, edge, align_type, thresh, ncomps, chunksize, pca, thr_override,
            fused)
    hint = _CAPACITY_HINTS.get(key) if speculate else None
    if hint is not None:
        cap = int(hint * 1.25) + 1024
        res = _run_chain(*args, cap, resample, defer=True)
        res._pending = (cap, lambda: _run_chain(*args, None, resample), key)
        return res if lazy else res.finalize()
    res, n_det, n_kept = _run_chain(*args, None, resample)
    if speculate:
        if len(_CAPACITY_HINTS) > 64:
            _CAPACITY_HINTS.clear()
        _CAPACITY_HINTS[key] = n_det
    return res


def _trim_to_counts(res, cap):
    """The one host sync of the capacity-sized chain: both counts at once, then trim (views).
    Returns (n_detected, n_kept); nothing is trimmed when the capacity was exceeded."""
    t = _lib.torch()
    n_det, n_kept = (int(v) for v in t.stack((res.offsets_detected[-1], res.offsets[-1])).tolist())
    if n_det <= cap:
        res.spt_detected = res.spt_detected[:n_det]
        res.spt, res.valid = res.spt[:n_kept], res.valid[:n_kept]
        res.waves = res.waves[:, :n_kept, :]
        if res.scores is not None:
            if n_kept == 0:
                res.scores = res.evals = res.evecs = None
            else:
                res.scores = res.scores[:n_kept]
    return n_det, n_kept


def _scan_filter(filt, FS):
    """True when `filt` runs through the fused recursive-filter kernels (an IIR design)."""
    probe = getattr(filt, "is_recursive", None)
    return bool(probe and probe(FS))


def _run_chain(x, FS, filt, sp_win, edge, align_type, thresh, ncomps, chunksize, pca, thr_override,
               fused, cap, resample=1, defer=False):
    """The chain; cap=None: exact sizes (two host syncs), else capacity-sized arrays and one
    sync at the end.  Returns (ChainResult, n_detected, n_kept) - or, with defer, only the
    ChainResult with everything queued and nothing read back (_trim_to_counts finishes it)."""
    t = _lib.require_cuda()
    res = ChainResult()
    sp_win = list(sp_win)
    falling = edge in _FALLING
    auto = thr_override is None and isinstance(thresh, str)
    frac = (8.0 if thresh == 'auto' else float(thresh)) if auto else None
    if filt is not None and fused and _scan_filter(filt, FS):
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
            factor=(-frac if falling else frac) if auto else None, cap=cap)
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
        spt, offsets = _ops.detect(sig, thr, falling, FS, cap=cap)
    res.spt_detected, res.offsets_detected = spt, offsets
    spt = spt.clone()
    if resample != 1:
        _ops.align_resampled(sig, spt, offsets, None, FS, sp_win, align_type == 'min', resample)
    else:
        _ops.align(sig, spt, offsets, None, FS, sp_win, align_type == 'min')
    spt, offsets = _ops.remove_doubles(spt, offsets, 1000.0 / FS, cap=cap)
    res.spt, res.offsets = spt, offsets
    waves, valid, n_invalid = _ops.extract(sig, spt, FS, sp_win, contacts=None, offsets=offsets)
    res.waves, res.valid, res.n_invalid = waves, valid, n_invalid
    res.time = _ops.window_params(sp_win, FS)[2]
    if pca and (cap is not None or spt.numel() > 0):
        gram, ssum, _, counts = _ops.pca_gram(waves, offsets=offsets, mode="f64")
        res.evals, res.evecs = _ops.pca_eig(gram, ssum, counts, top=ncomps)   # leading pairs only
        res.scores = _ops.pca_project(waves, res.evals, res.evecs, ncomps, offsets=offsets)
    if cap is None:
        return res, int(res.spt_detected.numel()), int(res.spt.numel())
    if defer:
        return res
    n_det, n_kept = _trim_to_counts(res, cap)
    return res, n_det, n_kept


def sort_frontend(spike_data, filt=None, sp_win=(-0.2, 0.8), edge='falling', align_type='min',
                  thresh='auto', ncomps=2, chunksize=1E6, contacts=None, slabs='auto', resample=1,
                  halo=None):
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

    slabs: a host-resident recording is cut into time slabs (multiples of `chunksize`) whose
    upload overlaps the chain on the previous slab ('auto': up to 6 slabs; 1: one upload, one
    chain).  Slabs are time shards on one GPU (dist.ShardedChain over a LocalFabric): spike
    times and waveforms are identical to the unsliced chain, PCA scores to round-off.
    halo: samples a slab keeps of each neighbour (default 4 windows).  A spike whose alignment
    walk needs more is counted on the device; the call then reruns unsliced, so the result
    never depends on the slicing.
    """
    t = _lib.require_cuda()
    FS = spike_data['FS']
    if contacts is None and slabs != 1:
        piped = _sort_frontend_slabs(spike_data, filt, sp_win, edge, align_type, thresh, ncomps,
                                     chunksize, slabs, resample, halo)
        if piped is not None:
            return piped
    x = _lib.to_device_2d(spike_data['data'])            # one H2D copy (pinned: ~55 GB/s)
    if contacts is not None:
        x = x[t.as_tensor(list(contacts), device=x.device)]
    res = run_chain(x, FS, filt=filt, sp_win=sp_win, edge=edge, align_type=align_type,
                    thresh=thresh, ncomps=ncomps, chunksize=chunksize, resample=resample)
    n_rows = x.shape[0]
    ids = list(range(n_rows)) if contacts is None else list(contacts)
    return results_to_host(res, ids, FS, ncomps)


def _merge_and_fetch(results, ids, FS, ncomps, time, status=None):
    """ChainResults of one or more time slabs/shards (in time order) -> the reference's
    per-contact dicts on the host.  The device first brings everything into the host's
    layout (ss_merge_shard: lists contact-major across slabs, one contiguous waveform block
    per contact), one D2H per array through pinned staging follows, and what is left for the
    host is one contiguous copy per contact and array."""
    t = _lib.torch()
    lib = _lib.load()
    n_c = len(ids)
    r0 = results[0]
    dev = r0.spt.device
    n_win = int(r0.waves.shape[0])
    n_kept = sum(int(r.spt.numel()) for r in results)
    have_scores = n_kept > 0 and all(r.scores is not None for r in results)
    n_det = sum(int(r.spt_detected.numel()) for r in results)

    def layout(offs):
        counts = t.stack([o[1:] - o[:-1] for o in offs])           # (S, n_c)
        n_tot = counts.sum(0)
        block = t.cumsum(n_tot, 0) - n_tot
        before = t.cumsum(counts, 0) - counts
        ends = t.cat((block, n_tot.sum(0, keepdim=True)))          # host-side offsets [n_c+1]
        return (block[None, :] + before).contiguous(), block.contiguous(), n_tot.contiguous(), ends

    base_k, block_k, ntot_k, ends_k = layout([r.offsets for r in results])
    base_d, _, _, ends_d = layout([r.offsets_detected for r in results])
    spt_all = t.empty(n_kept, dtype=t.float64, device=dev)
    det_all = t.empty(n_det, dtype=t.float64, device=dev)
    val_all = t.empty(n_kept, dtype=t.uint8, device=dev)
    sc_all = t.empty((n_kept, ncomps), dtype=t.float64, device=dev) if have_scores else None
    wav_all = t.empty(n_kept * n_win, dtype=t.float32, device=dev)
    st = _lib.stream_ptr()
    for k, r in enumerate(results):
        rc = lib.ss_merge_shard(_lib.ptr(r.spt), _lib.ptr(r.valid),
                                _lib.ptr(r.scores if have_scores else None), int(ncomps),
                                _lib.ptr(r.waves), n_win, int(r.waves.stride(0)), int(r.spt.numel()),
                                _lib.ptr(r.offsets), n_c,
                                _lib.ptr(base_k[k]), _lib.ptr(block_k), _lib.ptr(ntot_k),
                                _lib.ptr(spt_all), _lib.ptr(val_all), _lib.ptr(sc_all),
                                _lib.ptr(wav_all), st)
        _lib.check(rc, "ss_merge_shard")
        rc = lib.ss_merge_shard(_lib.ptr(r.spt_detected), None, None, 0, None, 0, 0,
                                int(r.spt_detected.numel()), _lib.ptr(r.offsets_detected), n_c,
                                _lib.ptr(base_d[k]), None, None, _lib.ptr(det_all), None, None, None, st)
        _lib.check(rc, "ss_merge_shard")
        _ops._launched((1 if r.spt.numel() else 0) + (1 if r.spt_detected.numel() else 0))
    dev_arrays = {'thresh': r0.thresh, 'det': det_all, 'spt': spt_all, 'valid': val_all,
                  'scores': sc_all, 'waves': wav_all, 'ends_k': ends_k, 'ends_d': ends_d,
                  'status': status}               # status word of a sharded step (dist.py)
    # One pinned host block per call, leased from a pool and handed out as the result arrays
    # themselves (views): no second host copy, no page faults on fresh memory.  A block
    # returns to the pool when the caller has dropped every array that points into it.
    spans, total = {}, 0
    for name, ten in dev_arrays.items():
        if ten is not None:
            spans[name] = (total, ten.numel() * ten.element_size())
            total = (total + spans[name][1] + 63) // 64 * 64
    pinned, block = _lease_host_block(max(total, 64))
    h = {}
    for name, ten in dev_arrays.items():
        if ten is None:
            h[name] = None
            continue
        off, nbytes = spans[name]
        if nbytes:
            pinned[off:off + nbytes].view(ten.dtype).copy_(ten.reshape(-1), non_blocking=True)
        h[name] = block[off:off + nbytes].view(_lib.np_dtype_full(ten.dtype))
    t.cuda.current_stream().synchronize()
    out = {'FS': FS, 'n_contacts': n_c, 'contacts': ids, 'spt': [None] * n_c,
           'spt_aligned': [None] * n_c, 'sp_waves': [None] * n_c, 'features': [None] * n_c}
    if h['status'] is not None:
        out['_status'] = int(h['status'][0])
    names = ["Ch0:PC%d" % j for j in range(ncomps)]
    ok, od = h['ends_k'], h['ends_d']
    valid_all = h['valid'].view(np.bool_)                 # uint8 0/1 -> bool, same bytes
    scores_all = h['scores'].reshape(-1, ncomps) if have_scores else None
    for r in range(n_c):
        lo, hi = int(ok[r]), int(ok[r + 1])
        out['spt'][r] = {'data': h['det'][int(od[r]):int(od[r + 1])],
                         'thresh': float(h['thresh'][r]), 'contact': ids[r]}
        out['spt_aligned'][r] = {'data': h['spt'][lo:hi]}
        valid = valid_all[lo:hi]
        wd = {'data': h['waves'][lo * n_win:hi * n_win].reshape(n_win, hi - lo, 1),
              'time': time, 'FS': FS}
        fd = {'data': scores_all[lo:hi] if have_scores else None, 'names': list(names)}
        if not valid.all():
            wd['is_valid'] = valid
            fd['is_valid'] = valid
        out['sp_waves'][r] = wd
        out['features'][r] = fd
    return out


_HOST_BLOCKS = []                  # [[pinned uint8 tensor, its ndarray, leased?]]
_MAX_HOST_BLOCKS = 6


class _Lease(object):
    """Exports a pooled pinned block through the array interface.  Every result array handed
    to the caller is a view whose base chain ends in the ndarray made from this object, so the
    object lives exactly as long as somebody can still see the block's bytes; its finalizer
    hands the block back to the pool."""

    def __init__(self, arr):
        self.__array_interface__ = dict(arr.__array_interface__)


def _release(entry):
    entry[2] = False


def _lease_host_block(nbytes):
    """(pinned uint8 tensor, ndarray over the same memory) of at least nbytes.  Leases are
    tracked explicitly (weakref.finalize on the lease object), not by reference counts."""
    import weakref
    t = _lib.torch()
    entry = next((e for e in _HOST_BLOCKS if not e[2] and e[1].nbytes >= nbytes), None)
    if entry is None:
        for i, e in enumerate(_HOST_BLOCKS):                       # free but too small: replace
            if not e[2]:
                del _HOST_BLOCKS[i]
                break
        if len(_HOST_BLOCKS) >= _MAX_HOST_BLOCKS:
            # the caller keeps many results alive: plain (pageable) memory, owned by the arrays
            arr = np.empty(nbytes, dtype=np.uint8)
            return t.from_numpy(arr), arr
        ten = t.empty(int(nbytes * 1.25) + 4096, dtype=t.uint8).pin_memory()
        entry = [ten, ten.numpy(), False]
        _HOST_BLOCKS.append(entry)
    entry[2] = True
    lease = _Lease(entry[1])
    weakref.finalize(lease, _release, entry)
    return entry[0], np.asarray(lease)


def to_host(ten, min_bytes=8 << 20):
    """NumPy copy of a CUDA tensor.  Big results (>= min_bytes) come back through a pooled PINNED
    block (one DMA at PCIe speed instead of a staged copy into fresh pageable memory: 640 MB of
    scores of 10^7 waveforms 12 ms instead of ~0.25 s); the array is a view of the lease and
    the block returns to the pool when the caller drops it."""
    t = _lib.torch()
    nbytes = ten.numel() * ten.element_size()
    if nbytes < min_bytes or not ten.is_cuda:
        return ten.cpu().numpy()
    ten = ten.contiguous()
    pinned, block = _lease_host_block(nbytes)
    pinned[:nbytes].view(ten.dtype).copy_(ten.reshape(-1), non_blocking=True)
    t.cuda.current_stream().synchronize()
    return block[:nbytes].view(_lib.np_dtype_full(ten.dtype)).reshape(tuple(ten.shape))


def results_to_host(res, ids, FS, ncomps):
    """ChainResult -> the reference's per-contact dicts on the host."""
    return _merge_and_fetch([res], ids, FS, ncomps, res.time)


def _slab_bounds(n, chunksize, n0, want):
    """Time slabs in whole chunks: [(a, b), ...]; the first one holds the n0 threshold
    samples, a short trailing chunk joins the last slab.  None if slicing is pointless."""
    cs = int(chunksize)
    n_chunks = -(-n // cs)
    full = n // cs                                        # chunks of full length
    if full < 2:
        return None
    S = min(int(want), full)
    per = -(-full // S)
    first = max(per, -(-min(n0, n) // cs))
    edges = [0]
    k = first
    while k < full:
        edges.append(k * cs)
        k += per
    edges.append(n)
    if len(edges) < 3:
        return None
    del n_chunks
    return list(zip(edges[:-1], edges[1:]))


def _sort_frontend_slabs(spike_data, filt, sp_win, edge, align_type, thresh, ncomps, chunksize,
                         slabs, resample=1, halo=None):
    """sort_frontend with the upload of slab s+1 overlapping the chain on slab s.  Returns
    None when the input does not qualify (device-resident, FIR / no filter, too short)."""
    t = _lib.torch()
    data = spike_data['data']
    if hasattr(data, "device_tensor") or filt is None or not _scan_filter(filt, spike_data['FS']):
        return None
    lazy = hasattr(data, "read_slab")                     # io.LazyRecording: slabs come from files
    if lazy:
        host = None
        shape, tdtype = tuple(data.shape), {np.dtype(np.int16): t.int16,
                                            np.dtype(np.float32): t.float32,
                                            np.dtype(np.float64): t.float64}.get(np.dtype(data.dtype))
        if tdtype is None or len(shape) != 2:
            return None
    elif isinstance(data, t.Tensor):
        if data.is_cuda or data.dim() != 2:
            return None
        host = data
    else:
        arr = np.asarray(data)
        if arr.ndim != 2 or arr.dtype not in (np.int16, np.float32, np.float64):
            return None
        host = t.from_numpy(np.ascontiguousarray(arr))
    if not lazy:
        if host.dtype not in (t.int16, t.float32, t.float64) or host.stride(1) != 1:
            return None
        shape, tdtype = tuple(host.shape), host.dtype
        # big pageable recordings go through the feeder's pinned double buffer (is_pinned asks the
        # driver about the pointer, so NumPy views of pinned memory are recognised too)
        if host.numel() * host.element_size() >= (64 << 20) and not host.is_pinned() and \
                os.environ.get("SS_PAGEABLE_STAGING", "1")[:1] != "0":
            data = _PageableSource(host.numpy())
            lazy, host = True, None
    FS = spike_data['FS']
    n_c, n = shape
    win0, win1, time = _ops.window_params(list(sp_win), FS)
    n_win = win1 - win0
    H = 4 * n_win if halo is None else int(halo)
    bounds = _slab_bounds(n, chunksize, int(10 * FS) if isinstance(thresh, str) else 0,
                          6 if slabs == 'auto' else slabs)
    if bounds is None or min(b - a for a, b in bounds) < max(H, 1):
        return None
    from . import dist as ssd
    import time as _time
    prof = [] if os.environ.get("SS_E2E_PROFILE") else None

    def mark(name):
        if prof is not None:
            prof.append((name, _time.perf_counter()))

    mark("start")
    S = len(bounds)
    dev = _lib.device()
    cur = t.cuda.current_stream()
    up, _ = _side_streams()
    up.wait_stream(cur)
    xdev = t.empty((n_c, n), dtype=tdtype, device=dev)
    arrived = []
    lib = _lib.load()
    esz = xdev.element_size()

    def upload(src_ptr, src_pitch, a, b):
        # one strided DMA per slab (torch's copy_ would stage a non-contiguous slice
        # through pageable memory)
        with t.cuda.stream(up):
            rc = lib.ss_copy_2d_async(xdev.data_ptr() + a * esz, xdev.stride(0) * esz, src_ptr,
                                      src_pitch, (b - a) * esz, n_c, 1, up.cuda_stream)
            _lib.check(rc, "ss_copy_2d_async")
            ev = t.cuda.Event()
            ev.record(up)
        return ev

    feeder = None
    if lazy:
        # files -> two pinned slab buffers (reader thread, one slab ahead) -> HBM
        feeder = _SlabFeeder(data, bounds, n_c, tdtype)
    else:
        for a, b in bounds:                               # all uploads queued up front
            arrived.append(upload(host.data_ptr() + a * esz, host.stride(0) * esz, a, b))
    mark("uploads queued")
    key = (n_c, H, n_win, S, str(dev))
    if key not in _SLAB_FABRICS:
        _SLAB_FABRICS[key] = ssd.LocalFabric(n_c, H, n_win, S, dev)
    fab = _SLAB_FABRICS[key]
    chains = [ssd.ShardedChain(xdev[:, a:b], FS, filt, sp_win, edge, align_type, thresh, ncomps,
                               chunksize, rank=k, n_ranks=S, index_offset=a, n_total=n, halo=H,
                               fabric=fab, resample=resample)
              for k, (a, b) in enumerate(bounds)]

    def finish(k):
        # slab k's neighbours are filtered: align, remove doubles, extract, partial Gram
        chains[k].fab_align()
        chains[k].fab_extract_gram()

    thr = None
    for k in range(S):
        if feeder is not None:
            a, b = bounds[k]
            buf = feeder.get(k)                           # blocks until slab k is read
            ev = upload(buf.data_ptr(), buf.stride(0) * esz, a, b)
            feeder.release_after(k, ev)
            arrived.append(ev)
        cur.wait_event(arrived[k])
        chains[k].prepare()
        if k == 0:
            thr = chains[0].head_threshold()
        chains[k].fab_filter_detect(thr)
        mark("filter %d" % k)
        if k > 0:
            finish(k - 1)
            mark("finish %d" % (k - 1))
    finish(S - 1)
    mark("finish last")
    res0 = chains[0].fab_project()                        # Gram of all slabs, ONE eigensolve
    for k in range(1, S):
        r = chains[k].res
        r.scores = _ops.pca_project(r.waves, res0.evals, res0.evecs, ncomps, offsets=r.offsets)
    mark("project queued")
    out = _merge_and_fetch([c.res for c in chains], list(range(n_c)), FS, ncomps, time,
                           status=chains[0]._status)
    if feeder is not None:
        feeder.close()
    mark("merged, fetched, assembled")
    if out.pop('_status', 0) >= ssd.SHARD_HALO_MISS:
        # a spike's alignment walk (or its window afterwards) left the H samples a slab keeps
        # of its neighbours: the slab result differs from the reference's.  The recording is on
        # the device in one piece by now - run it unsliced, where every sample is readable.
        res = run_chain(xdev, FS, filt=filt, sp_win=sp_win, edge=edge, align_type=align_type,
                        thresh=thresh, ncomps=ncomps, chunksize=chunksize, resample=resample)
        out = results_to_host(res, list(range(n_c)), FS, ncomps)
        mark("halo miss: unsliced rerun")
    if prof is not None:
        print("e2e profile (ms): " + ", ".join("%s %.2f" % (prof[i][0], 1e3 * (prof[i][1] - prof[i - 1][1]))
                                               for i in range(1, len(prof))), flush=True)
    return out


class _PageableSource(object):
    """A pageable in-memory recording (the usual NumPy array) as a slab source for _SlabFeeder:
    the slab is copied into the feeder's pinned buffer by a few threads (NumPy's copy releases
    the GIL), one slab ahead of the DMA - a strided 2-D copy straight out of pageable memory is
    staged by the CUDA runtime itself, synchronously and far below PCIe speed."""

    # (B200 box, 32 x 18 M int16: 2 / 4 / 8 / 16 threads -> 89 / 57 / 46 / 41 ms per call)
    THREADS = int(os.environ.get("SS_STAGING_THREADS", "0")) or max(1, min(8, (os.cpu_count() or 2) // 2))

    def __init__(self, arr):
        self.arr = arr
        self.shape = tuple(arr.shape)
        self.dtype = arr.dtype

    def read_slab(self, a, b, out):
        import concurrent.futures
        dst = out.numpy() if hasattr(out, "numpy") else out
        n_c = self.shape[0]
        step = max(1, (n_c + self.THREADS - 1) // self.THREADS)

        def copy(c0):
            np.copyto(dst[c0:c0 + step, :b - a], self.arr[c0:c0 + step, a:b])

        if n_c <= 1:
            copy(0)
            return
        with concurrent.futures.ThreadPoolExecutor(max_workers=self.THREADS) as pool:
            list(pool.map(copy, range(0, n_c, step)))


class _SlabFeeder(object):
    """Reads the time slabs of a lazy (file-backed) recording into two pinned buffers on a
    background thread, one slab ahead of the upload: disk, PCIe and the kernels overlap."""

    def __init__(self, source, bounds, n_c, tdtype):
        import queue
        import threading
        t = _lib.torch()
        self.source, self.bounds = source, bounds
        width = max(b - a for a, b in bounds)
        key = (n_c, width, tdtype)
        if key not in _FEEDER_BUFS:                        # pinned allocation costs milliseconds
            if len(_FEEDER_BUFS) > 2:
                _FEEDER_BUFS.clear()
            _FEEDER_BUFS[key] = [t.empty((n_c, width), dtype=tdtype).pin_memory() for _ in range(2)]
        self.bufs = _FEEDER_BUFS[key]
        self.free = [threading.Event(), threading.Event()]
        for f in self.free:
            f.set()
        self.uploaded = [None, None]
        self.ready = queue.Queue()
        self.error = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        try:
            for k, (a, b) in enumerate(self.bounds):
                slot = k & 1
                self.free[slot].wait()
                self.free[slot].clear()
                if self.uploaded[slot] is not None:
                    self.uploaded[slot].synchronize()      # the DMA out of this buffer is done
                self.source.read_slab(a, b, self.bufs[slot][:, :b - a])
                self.ready.put(k)
        except Exception as exc:                           # surfaced by get()
            self.error = exc
            self.ready.put(-1)

    def get(self, k):
        got = self.ready.get()
        if got != k:
            raise self.error if self.error is not None else RuntimeError("slab feeder out of order")
        return self.bufs[k & 1]

    def release_after(self, k, event):
        self.uploaded[k & 1] = event
        self.free[k & 1].set()

    def close(self):
        self.thread.join(timeout=10)


_FEEDER_BUFS = {}
_SLAB_FABRICS = {}
_STREAMS = {}


def _side_streams():
    t = _lib.torch()
    dev = t.cuda.current_device()
    if dev not in _STREAMS:
        _STREAMS[dev] = (t.cuda.Stream(), t.cuda.Stream())
    return _STREAMS[dev]
