"""Time-sharded multi-GPU chain: one process per GPU, ``torch.distributed`` (NCCL over
NVLink on the B200 box, gloo in the CPU tests) for the few real exchange steps.

The reference filters every (contact, chunk) unit independently
(``src/spike_sort/core/filters.py:137-141``), so a recording sharded by TIME in
multiples of ``chunksize`` needs no filter halo and no communication for the filter
(SURVEY.md section 8e).  What does cross shard boundaries, and how it is exchanged:

====================================  ===========================================
``'auto'`` threshold: first            computed by rank 0 (which owns them) from its
``int(10*FS)`` samples                 head chunks, **broadcast** (n_contacts float64)
(``extract.py:79``)
crossing between the last sample of    first filtered sample of every contact from
shard r and the first of r+1           the right neighbour (**send/recv**)
(``extract.py:91``)
windows of spikes near a shard edge    ``H`` filtered samples per contact from both
(``extract.py:160-163,245``)           neighbours (**send/recv**) into the margins of
                                       the shard's own buffer (kernels take
                                       ``halo_before/halo_after``)
``np.diff`` across the boundary in     last ALIGNED time of every contact from the
``remove_doubles`` (``:281``)          left neighbour (**send/recv**)
PCA over ALL spikes of a contact       Gram + sums + counts re-expressed for the zero
(``features.py:224-231``)              reference (**all-reduce**, float64, 2.9 MB at 384
                                       contacts); eigensolve replicated
global layout of the spike list        per-shard counts ride in the same all-reduce
====================================  ===========================================

Spike times are GLOBAL (``index_offset``): the concatenation of the per-rank lists in
rank order is, per contact, exactly the list the single-GPU chain produces
(``tests/test_gpu_sharded.py`` checks this bit for bit by running the phases of
several shards on one GPU).

``ShardedChain`` splits the chain into phases separated by exactly those exchanges.  Two
transports carry them:

* **peer memory** (default under NCCL worlds; ``PeerFabric``): every rank owns a small arena
  in torch symmetric memory; the library's own kernels (``csrc/exchange.cu``) store
  thresholds, halos, last spike times and Gram partials straight into the neighbours'
  arenas over NVLink and order themselves with a flag barrier kernel - no NCCL call, no
  torch op and no extra host sync between the stages.  ``LocalFabric`` gives the same
  kernels plain arenas on ONE device, which is how the tests (and a slab-pipelined
  single-GPU run) drive several shards in lock-step;
* **collectives** (``transport='nccl'``, and gloo on CPU tensors for the helpers,
  ``tests/test_dist_gloo.py``): broadcast / send-recv / all-reduce through
  ``torch.distributed``.
"""
import ctypes

from . import _lib, _ops
from .pipeline import ChainResult

_FALLING = ('falling', 'min')


def _dist():
    import torch.distributed as dist
    return dist


def world():
    d = _dist()
    if d.is_available() and d.is_initialized():
        return d.get_rank(), d.get_world_size()
    return 0, 1


# ------------------------------------------------------------------ collectives
def share_thresholds(thr, src=0):
    """Broadcast the per-contact thresholds from the rank owning the first 10 s."""
    rank, n = world()
    if n > 1:
        _dist().broadcast(thr, src=src)
    return thr


def exchange_with_neighbours(to_left, to_right):
    """Send `to_left` to rank-1 and `to_right` to rank+1; returns (from_left, from_right)
    (None where there is no neighbour).  Tensors are contiguous and of equal shape on
    every rank."""
    t = _lib.torch()
    rank, n = world()
    if n == 1:
        return None, None
    d = _dist()
    from_left = t.empty_like(to_right) if rank > 0 else None
    from_right = t.empty_like(to_left) if rank < n - 1 else None
    ops = []
    if rank > 0:
        ops.append(d.P2POp(d.isend, to_left.contiguous(), rank - 1))
        ops.append(d.P2POp(d.irecv, from_left, rank - 1))
    if rank < n - 1:
        ops.append(d.P2POp(d.isend, to_right.contiguous(), rank + 1))
        ops.append(d.P2POp(d.irecv, from_right, rank + 1))
    for req in d.batch_isend_irecv(ops):
        req.wait()
    return from_left, from_right


def exchange_first_samples(first):
    """`first`: (n_contacts,) float64, the first filtered sample of every contact of
    this shard.  Returns the right neighbour's `first` (NaN on the last rank: no
    crossing can be formed with NaN)."""
    t = _lib.torch()
    _, from_right = exchange_with_neighbours(first, first.new_zeros(first.shape))
    return from_right if from_right is not None else t.full_like(first, float('nan'))


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
    n_add = int(add.sum())                                  # the one host sync of this step
    if n_add == 0:
        return spt, offsets
    counts = offsets[1:] - offsets[:-1]
    new_off = t.zeros_like(offsets)
    new_off[1:] = t.cumsum(counts + add, 0)
    n_old = spt.numel()
    out = t.full((n_old + n_add,), float(t_last), dtype=spt.dtype, device=spt.device)
    # old element k of row r moves right by the number of boundary spikes of rows < r
    shift = t.cumsum(add, 0) - add                          # exclusive
    rows = t.repeat_interleave(t.arange(counts.numel(), device=spt.device), counts,
                               output_size=n_old)
    out[t.arange(n_old, device=spt.device) + shift[rows]] = spt
    return out, new_off


def merge_gram(gram, ssum, counts, with_layout=False):
    """All-reduce (sum) of per-contact Gram matrices, sums and spike counts (all expressed
    for the same reference).  with_layout: the same collective also carries every rank's
    per-contact counts (each rank fills its own slot of a zero vector), returned as a
    (world, n_contacts) tensor - the layout of the concatenated spike list."""
    rank, n = world()
    t = _lib.torch()
    if n == 1:
        return (gram, ssum, counts, counts[None, :]) if with_layout else (gram, ssum, counts)
    d = _dist()
    parts = [gram.reshape(-1), ssum.reshape(-1), counts.to(gram.dtype).reshape(-1)]
    if with_layout:
        slots = t.zeros((n, counts.numel()), dtype=gram.dtype, device=gram.device)
        slots[rank] = counts.to(gram.dtype)
        parts.append(slots.reshape(-1))
    flat = t.cat(parts)
    d.all_reduce(flat, op=d.ReduceOp.SUM)
    g, s_, c = gram.numel(), ssum.numel(), counts.numel()
    out = (flat[:g].reshape(gram.shape), flat[g:g + s_].reshape(ssum.shape),
           flat[g + s_:g + s_ + c].round().to(counts.dtype))
    if with_layout:
        out = out + (flat[g + s_ + c:].round().to(counts.dtype).reshape(n, c),)
    return out


def pick_reference(stack_flags, stack_refs):
    """Per contact, the reference waveform of the lowest shard that has a spike there.
    stack_flags: (world, n_contacts) bool, stack_refs: (world, n_contacts, n_win)."""
    t = _lib.torch()
    first = t.argmax(stack_flags.to(t.int64), dim=0)        # 0 if none
    idx = first[None, :, None].expand(1, stack_refs.shape[1], stack_refs.shape[2])
    return t.gather(stack_refs, 0, idx)[0].contiguous()


def share_reference(ref, has_spikes):
    """Every rank must shift its waveforms by the SAME reference before the Gram
    (ss_pca_gram): all-gather of (flag, ref) and a local pick."""
    rank, n = world()
    if n == 1:
        return ref
    d = _dist()
    t = _lib.torch()
    packed = t.cat([has_spikes.to(ref.dtype)[:, None], ref], dim=1).contiguous()
    allp = [t.empty_like(packed) for _ in range(n)]
    d.all_gather(allp, packed)
    stack = t.stack(allp)                                   # (world, n_contacts, 1 + n_win)
    return pick_reference(stack[:, :, 0] > 0.5, stack[:, :, 1:])


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



# ------------------------------------------------------------------ peer memory
class _Fabric(object):
    """Arenas of all shards as seen from this process (include/spikesort_b200.h, "time
    shards over peer memory")."""

    def __init__(self, n_c, H, n_win, world):
        lib = _lib.load()
        self.n_c, self.H, self.n_win, self.world = int(n_c), int(H), int(n_win), int(world)
        self.nbytes = int(lib.ss_shard_arena_bytes(self.n_c, self.H, self.n_win, self.world))
        off = (ctypes.c_size_t * 7)()
        _lib.check(lib.ss_shard_arena_offsets(self.n_c, self.H, self.n_win, self.world, off),
                   "ss_shard_arena_offsets")
        (self.off_flags, self.off_thr, self.off_from_left, self.off_from_right,
         self.off_prev_last, self.off_slots, self.slot_stride) = (int(v) for v in off)
        self.epoch = 0

    def _set_ptrs(self, ptrs):
        self.ptrs = (ctypes.c_void_p * self.world)(*[int(p) for p in ptrs])

    def local(self, rank, offset, count):
        """float64 view of `count` values at byte `offset` of rank's arena (own arenas only)."""
        a = self.arena(rank)
        return a[offset // 8:offset // 8 + count]

    def key(self):
        return (self.n_c, self.H, self.n_win, self.world)


class LocalFabric(_Fabric):
    """All shards on one device and one stream: plain allocations, stream order is the barrier."""

    def __init__(self, n_c, H, n_win, world, device=None):
        _Fabric.__init__(self, n_c, H, n_win, world)
        t = _lib.require_cuda()
        dev = device if device is not None else _lib.device()
        self._arenas = [t.zeros(self.nbytes // 8, dtype=t.float64, device=dev)
                        for _ in range(self.world)]
        self._set_ptrs([a.data_ptr() for a in self._arenas])

    def arena(self, rank):
        return self._arenas[rank]

    def barrier(self, rank):
        pass


class PeerFabric(_Fabric):
    """One process per GPU: the arena lives in torch symmetric memory, peers store into it
    over NVLink; ss_peer_barrier (flags in the arena itself) orders the steps."""

    def __init__(self, n_c, H, n_win, group=None):
        t = _lib.require_cuda()
        d = _dist()
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else d.group.WORLD
        self.rank, world = d.get_rank(group), d.get_world_size(group)
        _Fabric.__init__(self, n_c, H, n_win, world)
        self._arena = symm.empty(self.nbytes // 8, dtype=t.float64, device=_lib.device())
        self._arena.zero_()
        self._hdl = symm.rendezvous(self._arena, group)
        self._set_ptrs(self._hdl.buffer_ptrs)
        t.cuda.synchronize()
        d.barrier(group)                       # every arena is zeroed before anyone signals

    def arena(self, rank):
        if rank != self.rank:
            raise ValueError("only the local arena is addressable from the host side")
        return self._arena

    def barrier(self, rank):
        self.epoch += 1
        rc = _lib.load().ss_peer_barrier(self.ptrs, rank, self.world, self.epoch, 10.0,
                                         _lib.stream_ptr())
        _lib.check(rc, "ss_peer_barrier")
        _ops._launched(1)


_FABRICS = {}
_CAPACITY_HINTS = {}
SHARD_OVERFLOW, SHARD_HALO_MISS = 1, 2          # status word of a sharded step (header)


class HaloError(RuntimeError):
    """An alignment walk left even the widest halo a shard can hold of its neighbour."""


def peer_fabric(n_c, H, n_win, slot=0):
    """Cached PeerFabric of the default process group (the rendezvous costs milliseconds).
    slot: independent arena sets for steps that are in flight at the same time on different
    streams (every rank must use the same slot for the same step)."""
    key = (int(n_c), int(H), int(n_win), world()[1], int(_lib.torch().cuda.current_device()), int(slot))
    if key not in _FABRICS:
        _FABRICS[key] = PeerFabric(n_c, H, n_win)
    return _FABRICS[key]


# ---------------------------------------------------------------- phased chain
class ShardedChain(object):
    """One time shard of the chain, as phases separated by the exchange steps."""

    def __init__(self, x, FS, filt, sp_win=(-0.2, 0.8), edge='falling', align_type='min',
                 thresh='auto', ncomps=2, chunksize=1E6, rank=0, n_ranks=1, index_offset=0,
                 n_total=None, halo=None, fabric=None, resample=1):
        t = _lib.require_cuda()
        self.t = t
        self.x, self.FS, self.filt = x, FS, filt
        self.sp_win = list(sp_win)
        self.falling = edge in _FALLING
        self.use_min = align_type == 'min'
        self.thresh, self.ncomps, self.chunksize = thresh, ncomps, int(chunksize)
        self.resample = int(resample)
        self.rank, self.n_ranks = rank, n_ranks
        self.n_c, self.n = x.shape
        win0, win1, self.time = _ops.window_params(self.sp_win, FS)
        self.n_win = win1 - win0
        self.H = min(4 * self.n_win if halo is None else int(halo), self.n)
        self.shard = _ops.Shard(index_offset, self.n if n_total is None else n_total,
                                self.H if rank > 0 else 0, self.H if rank < n_ranks - 1 else 0)
        self.sos, self.padlen = filt.sections(FS)
        self.res = ChainResult()
        # spikes whose alignment / extraction window left the H neighbour samples (device counter)
        self.halo_miss = t.zeros(1, dtype=t.int32, device=x.device)
        # filtered signal with room for the neighbours' samples on both sides
        self.buf = t.empty((self.n_c, self.n + 2 * self.H), dtype=t.float64, device=x.device)
        self.sig = self.buf[:, self.H:self.H + self.n]
        self.fabric = fabric
        if fabric is not None and fabric.key() != (self.n_c, self.H, self.n_win, n_ranks):
            raise ValueError("fabric %r does not match the shard %r"
                             % (fabric.key(), (self.n_c, self.H, self.n_win, n_ranks)))

    def _align(self, spt):
        if self.resample != 1:
            _ops.align_resampled(self.sig, spt, self._offsets, None, self.FS, self.sp_win,
                                 self.use_min, self.resample, shard=self.shard,
                                 halo_miss=self.halo_miss)
        else:
            _ops.align(self.sig, spt, self._offsets, None, self.FS, self.sp_win, self.use_min,
                       shard=self.shard, halo_miss=self.halo_miss)

    # 1. sweep 1 of the filter (needs nothing from other ranks)
    def prepare(self):
        self.ws = _ops.filtfilt_prepare(self.x, self.sos, self.padlen, self.chunksize)

    # 2. rank 0: thresholds from its head chunks (then broadcast)
    def head_threshold(self):
        if not isinstance(self.thresh, str):
            return self.t.full((self.n_c,), float(self.thresh), dtype=self.t.float64,
                               device=self.x.device)
        frac = 8.0 if self.thresh == 'auto' else float(self.thresh)
        return _ops.filtfilt_head_threshold(self.x, self.sos, self.padlen, self.chunksize,
                                            self.sig, self.ws, int(10 * self.FS),
                                            -frac if self.falling else frac)

    # 3. sweep 2 with fused crossing marks; returns what the neighbours need
    def filter_detect(self, thr, slabs=True, cap=None, peer=None):
        r = self.res
        r.thresh = thr
        _, _, spt, offsets = _ops.filtfilt_detect(
            self.x, self.sos, self.padlen, self.chunksize, self.FS, self.falling, thr=thr,
            out=self.sig, ws=self.ws, shard=self.shard, cap=cap, peer=peer)
        r.filtered = self.sig
        self._spt, self._offsets = spt, offsets
        if not slabs:
            return None
        H = self.H
        left_slab = self.sig[:, :H].contiguous()            # becomes rank-1's right halo
        right_slab = self.sig[:, self.n - H:].contiguous()  # becomes rank+1's left halo
        return left_slab, right_slab

    # 4. halos in place, boundary crossing, alignment; returns the last aligned times
    def align(self, halo_left, halo_right):
        t, r = self.t, self.res
        H = self.H
        if halo_left is not None:
            self.buf[:, :H] = halo_left
        if halo_right is not None:
            self.buf[:, H + self.n:] = halo_right
            nxt = halo_right[:, 0].contiguous()
            mask = boundary_crossings(self.sig[:, self.n - 1], nxt, r.thresh, self.falling)
            t_last = (self.shard.index_offset + self.n - 1) * 1000.0 / self.FS
            self._spt, self._offsets = append_boundary_spikes(self._spt, self._offsets, mask,
                                                              t_last)
        r.spt_detected, r.offsets_detected = self._spt, self._offsets
        spt = self._spt.clone()
        self._align(spt)
        self._aligned = spt
        off = self._offsets
        last = t.full((self.n_c,), float('nan'), dtype=t.float64, device=spt.device)
        if spt.numel() > 0:
            has = off[1:] > off[:-1]
            last = t.where(has, spt[(off[1:] - 1).clamp(min=0)], last)    # no host sync
        return last

    # 5. duplicate removal (knowing the left neighbour's last times), extraction
    def extract(self, prev_last):
        t, r = self.t, self.res
        spt, offsets = _ops.remove_doubles(self._aligned, self._offsets, 1000.0 / self.FS,
                                           prev_last=prev_last)
        r.spt, r.offsets = spt, offsets
        waves, valid, n_invalid = _ops.extract(self.sig, spt, self.FS, self.sp_win, contacts=None,
                                               offsets=offsets, shard=self.shard,
                                               halo_miss=self.halo_miss)
        r.waves, r.valid, r.n_invalid, r.time = waves, valid, n_invalid, self.time
        counts = (offsets[1:] - offsets[:-1]).contiguous()
        if spt.numel() > 0:
            first = offsets[:-1].clamp(max=spt.numel() - 1)
            ref = waves[:, first, 0].t().to(t.float64).contiguous()
        else:
            ref = t.zeros((self.n_c, self.n_win), dtype=t.float64, device=spt.device)
        return ref, counts > 0

    # 6. partial Gram.  The kernel shifts by a LOCAL reference waveform (conditioning of the
    #    float32/tf32 products); the result is re-expressed for the zero reference in float64
    #    - sum (d+r)(d+r)^T = G + S r^T + r S^T + n r r^T - so that shards simply add up
    #    and no reference has to be agreed on between ranks
    def gram(self, ref):
        t = self.t
        gram, ssum, _, counts = _ops.pca_gram(self.res.waves, offsets=self.res.offsets, ref=ref, mode="f64")
        n = counts.to(t.float64)
        sr = ssum[:, :, None] * ref[:, None, :]
        gram0 = gram + sr + sr.transpose(1, 2) + n[:, None, None] * ref[:, :, None] * ref[:, None, :]
        sum0 = ssum + n[:, None] * ref
        return gram0, sum0, counts

    # 7. replicated eigensolve + local projection
    def project(self, gram, ssum, counts):
        r = self.res
        r.evals, r.evecs = _ops.pca_eig(gram.contiguous(), ssum.contiguous(), counts.contiguous(),
                                        top=self.ncomps)
        r.scores = _ops.pca_project(r.waves, r.evals, r.evecs, self.ncomps, offsets=r.offsets)
        return r

    # ---- the same phases over a fabric (peer memory): no torch op, no extra host sync
    def _xargs(self):
        f = self.fabric
        return self.n_c, self.H, self.n_win, f.ptrs

    def fab_put_thresholds(self, thr):
        """Rank owning the first 10 s: thresholds into every arena (then barrier)."""
        n_c, H, n_win, ptrs = self._xargs()
        rc = _lib.load().ss_shard_put_thresholds(_lib.ptr(thr), n_c, H, n_win, ptrs, self.n_ranks,
                                                 _lib.stream_ptr())
        _lib.check(rc, "ss_shard_put_thresholds")
        _ops._launched(1)

    def fab_thresholds(self):
        f = self.fabric
        return f.local(self.rank, f.off_thr, self.n_c).clone()

    def fab_filter_detect(self, thr, cap=None):
        """Sweep 2 with fused crossing marks AND fused halo put: the boundary samples go
        into the neighbours' arenas from the filter kernel itself (then barrier).  cap: capacity of the detected-spike list (no host sync; see
        pipeline.run_chain `speculate`)."""
        self.cap = cap
        self._overflow = None
        # the halo exchange is fused into sweep 2: its edge tiles store the first / last H
        # samples into the neighbours' arenas as they write them to HBM
        f = self.fabric
        left = int(f.ptrs[self.rank - 1]) + f.off_from_right if self.rank > 0 else 0
        right = int(f.ptrs[self.rank + 1]) + f.off_from_left if self.rank < self.n_ranks - 1 else 0
        peer = (left, right, self.H) if self.H > 0 and self.n_ranks > 1 else None
        self.filter_detect(thr, slabs=False, cap=cap, peer=peer)
        if cap is not None:
            self._overflow = (self._offsets[-1] > cap).to(self.t.int64)     # device flag

    def fab_align(self):
        """Halos into the margins, boundary crossing appended, alignment, last aligned time of
        every contact into the right neighbour's arena (then barrier)."""
        t, r, f, lib = self.t, self.res, self.fabric, _lib.load()
        n_c, H, n_win, ptrs = self._xargs()
        dev = self.sig.device
        flags = t.empty(n_c, dtype=t.int32, device=dev)
        rc = lib.ss_shard_install_halos(_lib.ptr(self.sig), self.sig.stride(0), n_c, self.n, H, n_win,
                                        _lib.ptr(f.arena(self.rank)), self.rank, self.n_ranks,
                                        _lib.ptr(r.thresh), int(self.falling), _lib.ptr(flags),
                                        _lib.stream_ptr())
        _lib.check(rc, "ss_shard_install_halos")
        _ops._launched(1)
        self._cap_detected = None
        if self.rank < self.n_ranks - 1:
            n_old = self._spt.numel()
            spt = t.empty(n_old + n_c, dtype=t.float64, device=dev)     # spare capacity
            offsets = t.empty(n_c + 1, dtype=t.int64, device=dev)
            t_last = (self.shard.index_offset + self.n - 1) * 1000.0 / self.FS
            rc = lib.ss_shard_append_boundary(_lib.ptr(self._spt), _lib.ptr(self._offsets), n_c,
                                              n_old, _lib.ptr(flags), float(t_last), _lib.ptr(spt),
                                              _lib.ptr(offsets), _lib.stream_ptr())
            _lib.check(rc, "ss_shard_append_boundary")
            _ops._launched(1)
            self._spt, self._offsets = spt, offsets
            self._cap_detected = offsets[-1]       # true length, read with remove_doubles' sync
        r.spt_detected, r.offsets_detected = self._spt, self._offsets
        spt = self._spt.clone()
        self._align(spt)
        self._aligned = spt
        rc = lib.ss_shard_put_last(_lib.ptr(spt), _lib.ptr(self._offsets), n_c, H, n_win, ptrs,
                                   self.rank, self.n_ranks, _lib.stream_ptr())
        _lib.check(rc, "ss_shard_put_last")
        _ops._launched(1 if self.rank < self.n_ranks - 1 else 0)

    def fab_extract_gram(self):
        """Duplicate removal (left neighbour's last times from the arena), extraction, partial
        Gram re-expressed for the zero reference into slot `rank` of every arena (then
        barrier)."""
        t, r, f, lib = self.t, self.res, self.fabric, _lib.load()
        n_c, H, n_win, ptrs = self._xargs()
        prev_last = f.local(self.rank, f.off_prev_last, n_c) if self.rank > 0 else None
        if self.cap is not None:
            # capacity mode: no host sync here, lists keep their spare room until finalize()
            spt, offsets = _ops.remove_doubles(self._aligned, self._offsets, 1000.0 / self.FS,
                                               prev_last=prev_last, cap=self._aligned.numel())
        elif self._cap_detected is not None:
            spt, offsets, n_det = _ops.remove_doubles(self._aligned, self._offsets, 1000.0 / self.FS,
                                                      prev_last=prev_last, also=self._cap_detected)
            r.spt_detected = r.spt_detected[:n_det]
        else:
            spt, offsets = _ops.remove_doubles(self._aligned, self._offsets, 1000.0 / self.FS,
                                               prev_last=prev_last)
        r.spt, r.offsets = spt, offsets
        waves, valid, n_invalid = _ops.extract(self.sig, spt, self.FS, self.sp_win, contacts=None,
                                               offsets=offsets, shard=self.shard,
                                               halo_miss=self.halo_miss)
        r.waves, r.valid, r.n_invalid, r.time = waves, valid, n_invalid, self.time
        gram, ssum, ref, _ = _ops.pca_gram(waves, offsets=offsets, mode="f64")
        rc = lib.ss_shard_gram_put(_lib.ptr(gram), _lib.ptr(ssum), _lib.ptr(ref), _lib.ptr(offsets),
                                   n_c, H, n_win, ptrs, self.rank, self.n_ranks,
                                   _lib.ptr(self._overflow), _lib.ptr(self.halo_miss),
                                   _lib.stream_ptr())
        _lib.check(rc, "ss_shard_gram_put")
        _ops._launched(1)

    def fab_project(self):
        """Sum of all shards' partials in rank order, replicated eigensolve, local projection."""
        t, r, f, lib = self.t, self.res, self.fabric, _lib.load()
        n_c, H, n_win, _ = self._xargs()
        dev = self.sig.device
        gram = t.empty((n_c, n_win, n_win), dtype=t.float64, device=dev)
        ssum = t.empty((n_c, n_win), dtype=t.float64, device=dev)
        counts = t.empty(n_c, dtype=t.int64, device=dev)
        all_counts = t.empty((self.n_ranks, n_c), dtype=t.int64, device=dev)
        self._status = t.empty(1, dtype=t.int64, device=dev)
        rc = lib.ss_shard_gram_reduce(_lib.ptr(f.arena(self.rank)), n_c, H, n_win, self.n_ranks,
                                      _lib.ptr(gram), _lib.ptr(ssum), _lib.ptr(counts),
                                      _lib.ptr(all_counts), _lib.ptr(self._status),
                                      _lib.stream_ptr())
        _lib.check(rc, "ss_shard_gram_reduce")
        _ops._launched(1)
        r.all_counts = all_counts
        return self.project(gram, ssum, counts)

    def finalize(self):
        """Capacity mode: the one host sync of the step - true list lengths and whether ANY
        shard overflowed its capacity (the status word that travelled with the Gram partials).
        Trims the results to views of their true length.  Returns (ok, n_detected)."""
        t, r = self.t, self.res
        n_det, n_kept, bad = (int(v) for v in t.stack(
            (r.offsets_detected[-1], r.offsets[-1], self._status[0])).tolist())
        self.status = bad                       # 0, SHARD_OVERFLOW or SHARD_HALO_MISS (any rank)
        if bad:
            return False, n_det
        r.spt_detected = r.spt_detected[:n_det]
        r.spt, r.valid = r.spt[:n_kept], r.valid[:n_kept]
        r.waves = r.waves[:, :n_kept, :]
        if r.scores is not None:
            r.scores = r.scores[:n_kept]
        return True, n_det


class PendingShard(object):
    """run_chain_sharded(lazy=True): this rank's step is queued, nothing has been read back.
    finalize() is the step's one host synchronisation (true list lengths + the status word
    every rank received with the Gram partials); if any rank overflowed its capacity or lost a
    window beyond the halo, ALL ranks see it and redo the step together."""

    def __init__(self, chain, redo, remember):
        self._chain, self._redo, self._remember = chain, redo, remember
        self._res = None

    def finalize(self):
        if self._res is None:
            ok, n_det = self._chain.finalize()
            self._res = self._chain.res if ok else self._redo(self._chain.status)
            if ok:
                self._remember(n_det)
            self._chain = None
        return self._res


def run_chain_sharded(x, FS, filt, sp_win=(-0.2, 0.8), edge='falling', align_type='min',
                      thresh='auto', ncomps=2, chunksize=1E6, index_offset=None, n_total=None,
                      transport=None, speculate=True, halo=None, lazy=False, slot=0):
    """The chain on this rank's time shard `x` (CUDA tensor, every rank the same number
    of samples unless index_offset/n_total say otherwise), with the exchange steps of
    the module docstring.  Spike times are global; `all_counts` gives the layout of the
    concatenated list.  transport: 'peer' (NVLink peer memory, default) or 'nccl'
    (torch.distributed collectives); SS_TRANSPORT overrides the default.  speculate: as
    pipeline.run_chain - capacity-sized lists from the previous call's count, one host sync
    at the end, exact rerun (agreed on by all ranks) if any shard overflowed.
    halo: neighbour samples kept on each side of the shard (default 4 windows).  A spike
    whose alignment walk leaves them is COUNTED on the device (ss_align / ss_extract
    `d_halo_miss`), every rank learns of it through the status word, and the step is redone
    with a halo four times as wide (up to the whole neighbour shard; beyond that HaloError) -
    the result always equals the unsharded chain's.
    lazy: (peer transport, a capacity hint from an earlier call) return a PendingShard with
    the step queued and no host synchronisation; its finalize() gives the result.
    slot: arena set of the step (peer transport).  Steps queued on DIFFERENT CUDA streams must
    use different slots - their exchange kernels then run independently of each other; all
    ranks pass the same slot for the same step."""
    import os
    t = _lib.require_cuda()
    rank, n_ranks = world()
    n = x.shape[1]
    if index_offset is None:
        index_offset = rank * n
    if n_total is None:
        n_total = n_ranks * n
    if transport is None:
        transport = os.environ.get("SS_TRANSPORT", "peer")
    if transport not in ("peer", "nccl"):
        raise ValueError("transport must be 'peer' or 'nccl'")
    win0, win1, _ = _ops.window_params(list(sp_win), FS)
    n_win = win1 - win0
    H0 = min(4 * n_win if halo is None else int(halo), n)
    use_fabric = transport == "peer" and n_ranks > 1
    prof = os.environ.get("SS_DIST_PROFILE") and rank == 0
    marks = []

    def mark(name):
        if prof:
            import time
            t.cuda.synchronize()
            marks.append((name, time.perf_counter()))

    def new_chain(H, fabric):
        return ShardedChain(x, FS, filt, sp_win, edge, align_type, thresh, ncomps, chunksize, rank,
                            n_ranks, index_offset, n_total, halo=H, fabric=fabric)

    def fabric_queue(H, cap):
        """Queue the whole step on the stream; returns the chain (nothing read back)."""
        fabric = peer_fabric(x.shape[0], H, n_win, slot)
        ch = new_chain(H, fabric)
        mark("start")
        ch.prepare()
        mark("prepare")
        if rank == 0:
            thr = ch.head_threshold()
            ch.fab_put_thresholds(thr)
        fabric.barrier(rank)
        if rank != 0:
            thr = ch.fab_thresholds()
        mark("threshold+put")
        ch.fab_filter_detect(thr, cap=cap)
        fabric.barrier(rank)
        mark("filter_detect+halo put")
        ch.fab_align()
        fabric.barrier(rank)
        mark("align+last put")
        ch.fab_extract_gram()
        fabric.barrier(rank)
        mark("extract+gram put")
        ch.fab_project()
        mark("reduce+project")
        return ch

    def collective_pass(H):
        ch = new_chain(H, None)
        mark("start")
        ch.prepare()
        mark("prepare")
        thr = ch.head_threshold() if rank == 0 else t.empty(x.shape[0], dtype=t.float64,
                                                            device=x.device)
        thr = share_thresholds(thr, 0)
        mark("threshold+bcast")
        left_slab, right_slab = ch.filter_detect(thr)
        mark("filter_detect")
        halo_left, halo_right = exchange_with_neighbours(left_slab, right_slab)
        mark("halo exchange")
        last = ch.align(halo_left, halo_right)
        mark("align")
        prev_last, _ = exchange_with_neighbours(t.zeros_like(last), last)
        mark("last exchange")
        ref, _ = ch.extract(prev_last)
        mark("extract")
        gram, ssum, counts = ch.gram(ref)
        gram, ssum, counts, all_counts = merge_gram(gram, ssum, counts, with_layout=True)
        mark("gram+allreduce")
        res = ch.project(gram, ssum, counts)
        res.all_counts = all_counts
        mark("project")
        missed = ch.halo_miss.to(t.float64)
        if n_ranks > 1:
            _dist().all_reduce(missed, op=_dist().ReduceOp.MAX)
        return res, (SHARD_HALO_MISS if float(missed.item()) > 0 else 0)

    key = (tuple(x.shape), x.dtype, float(FS), tuple(sp_win), edge, align_type, str(thresh),
           id(type(filt)), int(chunksize), n_ranks, rank, H0)

    def remember(n_det):
        if speculate:
            if len(_CAPACITY_HINTS) > 64:
                _CAPACITY_HINTS.clear()
            _CAPACITY_HINTS[key] = n_det

    def exact_until_clean(H, status):
        """Exact-size passes, the halo widened after every miss; all ranks take the same
        branches (the status word is the same on every rank)."""
        while True:
            if status >= SHARD_HALO_MISS:
                if H >= n:
                    raise HaloError("an alignment walk crossed a whole neighbour shard (%d "
                                    "samples): shard the recording into longer pieces" % n)
                H = min(4 * H, n)
            marks[:] = []
            if use_fabric:
                ch = fabric_queue(H, None)
                status = int(ch._status.item())
                res, n_det = ch.res, int(ch.res.spt_detected.numel())
            else:
                res, status = collective_pass(H)
                n_det = int(res.spt_detected.numel())
            if status == 0:
                remember(n_det)
                return res

    if use_fabric:
        # all ranks decide alike: the hint exists on every rank or on none (same call
        # history); an overflow or a halo miss on any rank is seen by all (status word)
        hint = _CAPACITY_HINTS.get(key) if speculate else None
        if hint is None:
            res = exact_until_clean(H0, 0)
        else:
            cap = int(hint * 1.25) + 1024 + x.shape[0]
            ch = fabric_queue(H0, cap)
            pending = PendingShard(ch, lambda status: exact_until_clean(H0, status), remember)
            if lazy:
                return pending
            res = pending.finalize()
    else:
        res = exact_until_clean(H0, 0)
    if prof:
        print("dist profile (ms): " + ", ".join(
            "%s %.3f" % (marks[i][0], 1000 * (marks[i][1] - marks[i - 1][1]))
            for i in range(1, len(marks))), flush=True)
    return PendingShard_done(res) if lazy else res


class PendingShard_done(object):
    """A finished step behind the PendingShard interface (first call of a shape, or the
    collective transport: nothing was deferred)."""

    def __init__(self, res):
        self._res = res

    def finalize(self):
        return self._res


def sort_frontend_sharded(spike_data, filt, sp_win=(-0.2, 0.8), edge='falling', align_type='min',
                          thresh='auto', ncomps=2, chunksize=1E6, transport=None):
    """Host in / host out on one rank of a time-sharded recording: `spike_data['data']` is
    THIS rank's shard (n_contacts, n_samples_per_rank) on the host (pinned for full PCIe
    speed).  Uploads it, runs run_chain_sharded (exchanges over NVLink) and returns this
    shard's part of every contact's results as the reference's dicts, spike times global -
    concatenating the ranks' lists in rank order gives the whole recording's lists
    ('all_counts' holds every rank's per-contact counts)."""
    from .pipeline import results_to_host
    x = _lib.to_device_2d(spike_data['data'])
    FS = spike_data['FS']
    res = run_chain_sharded(x, FS, filt, sp_win, edge, align_type, thresh, ncomps, chunksize,
                            transport=transport)
    out = results_to_host(res, list(range(x.shape[0])), FS, ncomps)
    out['all_counts'] = res.all_counts.cpu().numpy()
    return out
