"""Time-sharded multi-GPU chain: one process per GPU, ``torch.distributed`` (NCCL over
NVLink on the B200 box, gloo in the CPU tests) for the few real exchange steps.

The reference filters every (contact, chunk) unit independently
(``src/spike_sort/core/filters.py:137-141``), so a recording sharded by TIME in
multiples of ``chunksize`` needs no filter halo and no communication for the filter
(SURVEY.md section 8e).  What does cross shard boundaries:

* the ``'auto'`` threshold uses only the first ``int(10*FS)`` samples
  (``extract.py:79``): the rank owning them computes it, everyone gets it by
  **broadcast**;
* a crossing between the last sample of shard r and the first of shard r+1
  (``extract.py:91``): one float64 per contact from the right neighbour
  (**send/recv**), merged into the spike list of shard r;
* PCA is over ALL spikes of a contact (``features.py:224-231``): per-contact Gram
  matrices, sums and counts are **all-reduced** (float64, a few MB), the reference
  waveform that conditions the Gram is **broadcast** from rank 0, the tiny eigensolve
  is replicated, projection stays local;
* per-shard spike counts are **all-gathered** so every rank knows the global layout
  of the (time-ordered) concatenated spike list.

The collective helpers below work on whatever device the tensors live on, so the
world_size-2 gloo tests exercise exactly this code on CPU tensors.
"""
from . import _lib, _ops

_FALLING = ('falling', 'min')


def _dist():
    import torch.distributed as dist
    return dist


def world():
    d = _dist()
    if d.is_available() and d.is_initialized():
        return d.get_rank(), d.get_world_size()
    return 0, 1


def share_thresholds(thr, src=0):
    """Broadcast the per-contact thresholds from the rank owning the first 10 s."""
    rank, n = world()
    if n > 1:
        _dist().broadcast(thr, src=src)
    return thr


def exchange_first_samples(first):
    """`first`: (n_contacts,) float64, the first filtered sample of every contact of
    this shard.  Returns the right neighbour's `first` (NaN on the last rank: no
    crossing can be formed with NaN)."""
    t = _lib.torch()
    rank, n = world()
    nxt = t.full_like(first, float('nan'))
    if n == 1:
        return nxt
    d = _dist()
    ops = []
    if rank > 0:
        ops.append(d.P2POp(d.isend, first.contiguous(), rank - 1))
    if rank < n - 1:
        ops.append(d.P2POp(d.irecv, nxt, rank + 1))
    for req in d.batch_isend_irecv(ops):
        req.wait()
    return nxt


def boundary_crossings(last, nxt, thr, falling):
    """Mask (n_contacts,) of crossings between this shard's last sample and the next
    shard's first sample, the rule of extract.py:86-91."""
    if falling:
        return (last > thr) & (nxt < thr)
    return (last < thr) & (nxt > thr)


def append_boundary_spikes(spt, offsets, mask, t_last):
    """Insert time `t_last` at the END of the list of every contact flagged in `mask`
    (it is the largest index of the shard, so order is kept).  Returns (spt, offsets)."""
    t = _lib.torch()
    add = mask.to(t.int64)
    if int(add.sum()) == 0:
        return spt, offsets
    counts = offsets[1:] - offsets[:-1]
    new_counts = counts + add
    new_off = t.zeros_like(offsets)
    new_off[1:] = t.cumsum(new_counts, 0)
    out = t.empty(int(new_off[-1]), dtype=spt.dtype, device=spt.device)
    rows = t.repeat_interleave(t.arange(counts.numel(), device=spt.device), counts)
    pos = t.arange(spt.numel(), device=spt.device) - offsets[:-1][rows] + new_off[:-1][rows]
    out[pos] = spt
    out[(new_off[1:] - 1)[mask]] = t_last
    return out, new_off


def merge_gram(gram, ssum, counts):
    """All-reduce (sum) of per-contact Gram matrices, sums and spike counts."""
    rank, n = world()
    if n > 1:
        d = _dist()
        t = _lib.torch()
        flat = t.cat([gram.reshape(-1), ssum.reshape(-1), counts.to(gram.dtype).reshape(-1)])
        d.all_reduce(flat, op=d.ReduceOp.SUM)
        g, s = gram.numel(), ssum.numel()
        gram = flat[:g].reshape(gram.shape)
        ssum = flat[g:g + s].reshape(ssum.shape)
        counts = flat[g + s:].round().to(counts.dtype)
    return gram, ssum, counts


def share_reference(ref, has_spikes):
    """Every rank must shift its waveforms by the SAME reference before the Gram
    (ss_pca_gram): take, per contact, the reference of the lowest rank that has a
    spike there.  Implemented as an all-gather of (flag, ref) and a local pick."""
    rank, n = world()
    if n == 1:
        return ref
    d = _dist()
    t = _lib.torch()
    packed = t.cat([has_spikes.to(ref.dtype)[:, None], ref], dim=1).contiguous()
    allp = [t.empty_like(packed) for _ in range(n)]
    d.all_gather(allp, packed)
    stack = t.stack(allp)                                   # (world, n_contacts, 1 + n_win)
    flags = stack[:, :, 0] > 0.5
    first = t.argmax(flags.to(t.int64), dim=0)              # lowest rank with a spike (0 if none)
    idx = first[None, :, None].expand(1, stack.shape[1], stack.shape[2] - 1)
    return t.gather(stack[:, :, 1:], 0, idx)[0].contiguous()


def gather_counts(offsets):
    """(world, n_contacts) spike counts of every shard, on every rank."""
    t = _lib.torch()
    rank, n = world()
    counts = (offsets[1:] - offsets[:-1]).contiguous()
    if n == 1:
        return counts[None, :]
    allc = [t.empty_like(counts) for _ in range(n)]
    _dist().all_gather(allc, counts)
    return t.stack(allc)


def run_chain_sharded(x, FS, filt, sp_win=(-0.2, 0.8), edge='falling', align_type='min',
                      thresh='auto', ncomps=2, chunksize=1E6):
    """The chain of pipeline.run_chain on this rank's time shard `x` (CUDA tensor,
    n_samples a multiple of chunksize), with the exchange steps listed in the module
    docstring.  Spike times are relative to the start of the shard; `all_counts`
    gives the global layout."""
    from .pipeline import ChainResult
    t = _lib.require_cuda()
    rank, n_ranks = world()
    sp_win = list(sp_win)
    res = ChainResult()
    falling = edge in _FALLING
    if filt is not None:
        sos, padlen = filt.sections(FS)
        sig = _ops.filtfilt_sos(x, sos, padlen, int(chunksize))
    else:
        sig = x
    res.filtered = sig
    n_rows, n = sig.shape
    if isinstance(thresh, str):
        frac = 8.0 if thresh == 'auto' else float(thresh)
        if rank == 0:
            thr = _ops.var_threshold(sig, int(10 * FS), -frac if falling else frac)
        else:
            thr = t.empty(n_rows, dtype=t.float64, device=sig.device)
        thr = share_thresholds(thr, 0)
    else:
        thr = t.full((n_rows,), float(thresh), dtype=t.float64, device=sig.device)
    res.thresh = thr
    spt, offsets = _ops.detect(sig, thr, falling, FS)
    nxt = exchange_first_samples(sig[:, 0].to(t.float64).contiguous())
    mask = boundary_crossings(sig[:, n - 1].to(t.float64), nxt, thr, falling)
    spt, offsets = append_boundary_spikes(spt, offsets, mask, (n - 1) * 1000.0 / FS)
    res.spt_detected, res.offsets_detected = spt, offsets
    spt = spt.clone()
    _ops.align(sig, spt, offsets, None, FS, sp_win, align_type == 'min')
    spt, offsets = _ops.remove_doubles(spt, offsets, 1000.0 / FS)
    res.spt, res.offsets = spt, offsets
    res.all_counts = gather_counts(offsets)
    waves, valid, n_invalid = _ops.extract(sig, spt, FS, sp_win, contacts=None, offsets=offsets)
    res.waves, res.valid, res.n_invalid = waves, valid, n_invalid
    res.time = _ops.window_params(sp_win, FS)[2]
    # PCA over all shards: shared reference, all-reduced Gram, replicated eigensolve
    n_win = waves.shape[0]
    counts = (offsets[1:] - offsets[:-1]).contiguous()
    if spt.numel() > 0:
        first = offsets[:-1].clamp(max=spt.numel() - 1)
        ref = waves[:, first, 0].t().to(t.float64).contiguous()
    else:
        ref = t.zeros((n_rows, n_win), dtype=t.float64, device=sig.device)
    ref = share_reference(ref, counts > 0)
    gram, ssum, _, _ = _ops.pca_gram(waves, offsets=offsets, ref=ref)
    gram, ssum, counts = merge_gram(gram, ssum, counts)
    res.evals, res.evecs = _ops.pca_eig(gram.contiguous(), ssum.contiguous(), counts.contiguous())
    res.scores = _ops.pca_project(waves, res.evals, res.evecs, ncomps, offsets=offsets)
    return res
