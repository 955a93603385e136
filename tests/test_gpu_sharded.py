"""GPU tier: the time-sharded chain (spikesort_b200/dist.py) equals the unsharded chain.

The phases of several shards are run in lock-step on ONE GPU, with the exchange steps
(thresholds, neighbour halos, last aligned times, Gram sum) done
by hand exactly as run_chain_sharded does them over NCCL.  Concatenated per contact in
shard order, spike times and waveforms must be bit-identical to pipeline.run_chain on
the whole recording; PCA scores agree to round-off (the Gram is summed in another order).
"""
import numpy as np
import pytest

from util import sign_normalise, rel_err

pytestmark = pytest.mark.gpu


def _run_sharded(torch, dist_mod, x_full, S, FS, flt, sp_win, chunksize, thresh):
    n = x_full.shape[1] // S
    chains = [dist_mod.ShardedChain(x_full[:, r * n:(r + 1) * n], FS, flt, sp_win, 'falling',
                                    'min', thresh, 2, chunksize, rank=r, n_ranks=S,
                                    index_offset=r * n, n_total=S * n) for r in range(S)]
    for c in chains:
        c.prepare()
    thr = chains[0].head_threshold()                                   # broadcast
    slabs = [c.filter_detect(thr) for c in chains]
    lasts = []
    for r, c in enumerate(chains):                                     # neighbour send/recv
        hl = slabs[r - 1][1] if r > 0 else None
        hr = slabs[r + 1][0] if r < S - 1 else None
        lasts.append(c.align(hl, hr))
    refs = [c.extract(lasts[r - 1] if r > 0 else None) for r, c in enumerate(chains)]
    parts = [c.gram(ref) for c, (ref, _) in zip(chains, refs)]       # each shard its own reference
    gram = sum(p[0] for p in parts)                                    # all-reduce
    ssum = sum(p[1] for p in parts)
    counts = sum(p[2] for p in parts)
    return [c.project(gram, ssum, counts) for c in chains], thr


def _run_sharded_fabric(torch, dist_mod, x_full, S, FS, flt, sp_win, chunksize, thresh, cap=None):
    """The peer-memory transport (csrc/exchange.cu) with all arenas on one device: the same
    kernels and call order as run_chain_sharded(transport='peer'), stream order standing in
    for the NVLink flag barrier."""
    from spikesort_b200 import _ops
    n = x_full.shape[1] // S
    win0, win1, _ = _ops.window_params(list(sp_win), FS)
    fab = dist_mod.LocalFabric(x_full.shape[0], min(4 * (win1 - win0), n), win1 - win0, S)
    chains = [dist_mod.ShardedChain(x_full[:, r * n:(r + 1) * n], FS, flt, sp_win, 'falling',
                                    'min', thresh, 2, chunksize, rank=r, n_ranks=S,
                                    index_offset=r * n, n_total=S * n, fabric=fab)
              for r in range(S)]
    for c in chains:
        c.prepare()
    thr0 = chains[0].head_threshold()
    chains[0].fab_put_thresholds(thr0)
    thrs = [thr0] + [c.fab_thresholds() for c in chains[1:]]
    for c, thr in zip(chains, thrs):
        c.fab_filter_detect(thr, cap=cap)
    for c in chains:
        c.fab_align()
    for c in chains:
        c.fab_extract_gram()
    out = [c.fab_project() for c in chains]
    if cap is not None:
        # capacity mode: lists carry spare room until the one host sync at the end; a capacity
        # that is too small on ANY shard is reported to ALL of them
        oks = [c.finalize()[0] for c in chains]
        assert all(oks) or not any(oks)
        if not all(oks):
            return None, thr0
    return out, thr0


@pytest.mark.parametrize("transport", ["collective", "fabric", "fabric-capacity"])
@pytest.mark.parametrize("S,chunksize", [(2, 20000), (3, 10000), (4, 30000)])
def test_sharded_chain_equals_whole(cuda, S, chunksize, transport):
    from spikesort_b200 import filters
    _check_sharded(cuda, S, chunksize, transport, filters.Filter(300., 40.))


@pytest.mark.parametrize("transport", ["collective", "fabric-capacity"])
def test_sharded_chain_equals_whole_band_pass(cuda, transport):
    """The same with a higher-order band-pass: the int16 shards go through the tile-per-lane
    summary sweep (separate prepare call) and the section-by-section second sweep."""
    import numpy as np
    from spikesort_b200 import filters
    flt = filters.Filter(np.array([150., 900.]), np.array([60., 1200.]))
    assert flt.sections(3000)[0].shape[0] >= 2
    _check_sharded(cuda, 3, 10000, transport, flt)


def _check_sharded(cuda, S, chunksize, transport, flt):
    import torch
    from spikesort_b200 import dist as ssd, filters, pipeline, synth
    FS, n_c = 3000, 5
    n = 30000 * 2                              # per shard; 10*FS = 30000 <= n
    if n % chunksize:
        n = (n // chunksize) * chunksize
    x = synth.make_recording(n_c, S * n, FS, seed=31 + S, rate_hz=30.0, noise=6.0,
                             amp=(70.0, 160.0), burst=True, negative=True).to(cuda)
    # spikes sitting on every shard boundary: crossing, window and doubles across the edge
    tpl = torch.tensor([0, -40, -120, -200, -120, -40, 20, 30, 10, 0], dtype=torch.int16, device=cuda)
    for r in range(1, S):
        for c, d in enumerate((-7, -3, -1, 0, 2)):
            x[c, r * n + d:r * n + d + 10] += tpl
        x[1, r * n + 1:r * n + 11] += tpl                              # a double near the edge
    sp_win = [-1.0, 2.0]                                               # 3 + 6 samples at 3 kHz
    whole = pipeline.run_chain(x, FS, filt=flt, sp_win=sp_win, edge='falling', align_type='min',
                               thresh='4', ncomps=2, chunksize=chunksize)
    run = _run_sharded if transport == "collective" else _run_sharded_fabric
    if transport == "fabric-capacity":
        # far too small on every shard -> all report overflow; then with room to spare
        none, _ = run(torch, ssd, x, S, FS, flt, sp_win, chunksize, '4', cap=7)
        assert none is None
        shards, thr = run(torch, ssd, x, S, FS, flt, sp_win, chunksize, '4', cap=4000)
    else:
        shards, thr = run(torch, ssd, x, S, FS, flt, sp_win, chunksize, '4')
    assert torch.equal(thr, whole.thresh)
    assert torch.equal(torch.cat([s.filtered for s in shards], dim=1), whole.filtered)
    tot = 0
    for c in range(n_c):
        w = whole.per_contact(c)
        parts = [s.per_contact(c) for s in shards]
        det = np.concatenate([p['spt_detected'] for p in parts])
        spt = np.concatenate([p['spt'] for p in parts])
        wav = np.concatenate([p['waves'] for p in parts], axis=1)
        val = np.concatenate([p['is_valid'] for p in parts])
        sc = np.concatenate([p['scores'] for p in parts], axis=0)
        assert np.array_equal(det, w['spt_detected']), c
        assert np.array_equal(spt, w['spt']), c
        assert np.array_equal(wav, w['waves']) and np.array_equal(val, w['is_valid']), c
        assert rel_err(sign_normalise(sc, w['scores']), w['scores']) < 1e-9, c
        tot += len(spt)
    assert tot > 200
    if transport.startswith("fabric"):
        # layout of the concatenated list, identical on every shard
        layout = torch.stack([s.offsets[1:] - s.offsets[:-1] for s in shards])
        for s in shards:
            assert torch.equal(s.all_counts, layout)
        assert torch.equal(shards[0].evecs, shards[-1].evecs)          # bit-identical replicas
    # something really happened at the boundaries
    t_edges = [r * n * 1000.0 / FS for r in range(1, S)]
    near = sum(int(((whole.spt > te - 3.0) & (whole.spt < te + 3.0)).sum()) for te in t_edges)
    assert near >= S - 1
