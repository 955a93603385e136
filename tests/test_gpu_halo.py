"""GPU tier: alignment walks that leave the neighbour samples a time shard keeps (the halo).

The reference is one process and simply keeps walking (``extract.py:241-265``).  A shard can
only read `halo` samples of its neighbours; the kernels count every spike whose window
needed more (``d_halo_miss`` of ss_align / ss_extract), the count travels with the Gram
partials to every shard, and the step is redone - unsliced on one GPU, with a wider halo
across GPUs - so the result always equals the unsharded chain's.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ramp_recording(torch, cuda, n, at, length):
    """float64 row: flat baseline with one long ASCENDING ramp of negative values starting at
    `at` (its minimum); aligning on the minimum walks down it to the left, 3 samples per pass
    (with sp_win = [-1, 2] ms at 3 kHz only the window's first sample re-queues a spike)."""
    x = torch.zeros((1, n), dtype=torch.float64, device=cuda)
    x[0, at:at + length] = -torch.arange(length, 0, -1, dtype=torch.float64, device=cuda)
    return x


def test_align_counts_windows_beyond_the_halo(cuda):
    import torch
    from spikesort_b200 import _ops
    FS, sp_win = 3000.0, [-1.0, 2.0]                       # window = 9 samples
    n, split, length = 4000, 2000, 140
    at = split - 110
    x = _ramp_recording(torch, cuda, n, at, length)
    t0 = (split + 9) * 1000.0 / FS                         # in the right shard, on the ramp
    whole = torch.tensor([t0], dtype=torch.float64, device=cuda)
    _ops.align(x, whole, None, None, FS, sp_win, True)
    assert float(whole[0]) < (split - 100) * 1000.0 / FS   # walked down the whole ramp

    def shard_align(H):
        # right shard = samples [split, n) with H samples of the left neighbour before it
        buf = x[:, split - H:].contiguous()
        sh = _ops.Shard(index_offset=split, n_total=n, halo_before=H, halo_after=0)
        spt = torch.tensor([t0], dtype=torch.float64, device=cuda)
        miss = torch.zeros(1, dtype=torch.int32, device=cuda)
        _ops.align(buf[:, H:], spt, None, None, FS, sp_win, True, shard=sh, halo_miss=miss)
        return float(spt[0]), int(miss[0])

    t_short, miss_short = shard_align(36)                  # the default 4 windows: not enough
    assert miss_short == 1 and t_short != float(whole[0])
    t_wide, miss_wide = shard_align(144)
    assert miss_wide == 0 and t_wide == float(whole[0])
    # without a neighbour on that side (halo 0 = start of the recording) nothing is counted
    sh = _ops.Shard(index_offset=0, n_total=n - split, halo_before=0, halo_after=0)
    spt = torch.tensor([9 * 1000.0 / FS], dtype=torch.float64, device=cuda)
    miss = torch.zeros(1, dtype=torch.int32, device=cuda)
    _ops.align(x[:, split:].contiguous(), spt, None, None, FS, sp_win, True, shard=sh, halo_miss=miss)
    assert int(miss[0]) == 0


def test_extract_counts_windows_beyond_the_halo(cuda):
    import torch
    from spikesort_b200 import _ops
    FS, sp_win = 3000.0, [-1.0, 2.0]
    n, split, H = 4000, 2000, 4
    x = torch.arange(n, dtype=torch.float64, device=cuda)[None, :].contiguous()
    buf = x[:, :split + H].contiguous()
    sh = _ops.Shard(index_offset=0, n_total=n, halo_before=0, halo_after=H)
    spt = torch.tensor([(split - 20) * 1000.0 / FS, (split - 1) * 1000.0 / FS],
                       dtype=torch.float64, device=cuda)
    off = torch.tensor([0, 2], dtype=torch.int64, device=cuda)
    miss = torch.zeros(1, dtype=torch.int32, device=cuda)
    _ops.extract(buf[:, :split], spt, FS, sp_win, contacts=None, offsets=off, shard=sh, halo_miss=miss)
    assert int(miss[0]) == 1                               # the second window needs 6 > 4 samples
    cdev = torch.tensor([0], dtype=torch.int32, device=cuda)
    miss.zero_()
    _ops.extract(buf[:, :split], spt, FS, sp_win, contacts=cdev, shard=sh, halo_miss=miss)
    assert int(miss[0]) >= 1


def _boundary_recording(n_c, n, FS, bounds, seed=3):
    """int16 recording with spikes sitting on every slab / shard boundary in `bounds`."""
    import torch
    from spikesort_b200 import synth
    x = synth.make_recording(n_c, n, FS, seed=seed, rate_hz=30.0, noise=6.0, amp=(70.0, 160.0),
                             burst=True, negative=True)
    tpl = torch.tensor([0, -40, -120, -200, -120, -40, 20, 30, 10, 0], dtype=torch.int16)
    for b in bounds:
        for c, d in enumerate((-7, -3, -1, 0, 2)[:n_c]):
            x[c, b + d:b + d + 10] += tpl
    return x


def test_sort_frontend_slabs_fall_back_when_a_window_leaves_the_halo(cuda, monkeypatch):
    """With a halo too narrow for the spikes that sit on the slab boundaries the slab pipeline
    must notice (status word travelling with the Gram partials), rerun unsliced and return
    exactly what the unsliced call returns; with the default halo it stays sliced.  (For
    high-pass filtered data a real walk beyond 4 windows across a chunk boundary needs a slope
    that stays monotonic through filtfilt's chunk-edge transient - the narrow halo stands in.)"""
    import spikesort_b200 as ss
    from spikesort_b200 import filters, pipeline
    FS, cs = 3000, 30000
    n = 4 * cs
    x = _boundary_recording(5, n, FS, [cs, 2 * cs, 3 * cs]).numpy()
    raw = {'data': x, 'FS': FS, 'n_contacts': 5}
    kw = dict(filt=filters.Filter(300., 40.), sp_win=[-1.0, 2.0], thresh='4', chunksize=cs)
    calls = []
    real = pipeline.run_chain
    monkeypatch.setattr(pipeline, "run_chain", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    narrow = ss.sort_frontend(raw, slabs=4, halo=2, **kw)
    assert len(calls) == 1, "the slab pipeline did not fall back to the unsliced chain"
    calls[:] = []
    sliced = ss.sort_frontend(raw, slabs=4, **kw)
    assert len(calls) == 0, "the default halo should have been enough"
    whole = ss.sort_frontend(raw, slabs=1, **kw)
    for got in (narrow, sliced):
        for c in range(5):
            assert np.array_equal(got['spt'][c]['data'], whole['spt'][c]['data'])
            assert np.array_equal(got['spt_aligned'][c]['data'], whole['spt_aligned'][c]['data'])
            assert np.array_equal(got['sp_waves'][c]['data'], whole['sp_waves'][c]['data'])
