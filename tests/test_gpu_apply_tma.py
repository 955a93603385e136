"""Sweep 2 through the async-copy engine (filt_apply_tma_kernel: bulk load of the int16 tile,
SWIZZLE_128B shared tile, one tensor store per tile) against the element kernel it replaces for
orders 1-2 on int16 recordings: bit-identical outputs and spike lists, whatever the chunk
geometry puts on the fast path."""
import os

import numpy as np
import pytest
from scipy import signal

from util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(cuda):
    import torch
    import spikesort_b200
    from spikesort_b200 import _ops, _lib, pipeline, synth
    _lib.load()
    return dict(torch=torch, ss=spikesort_b200, ops=_ops, pipeline=pipeline, synth=synth, dev=cuda)


class tma_mode:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.old = os.environ.get("SS_APPLY_TMA")
        os.environ["SS_APPLY_TMA"] = str(self.mode)

    def __exit__(self, *exc):
        if self.old is None:
            del os.environ["SS_APPLY_TMA"]
        else:
            os.environ["SS_APPLY_TMA"] = self.old


DESIGNS = {
    'hp1': lambda: signal.iirdesign(800 * 2 / 25000, 100 * 2 / 25000, 1, 10, ftype='butter'),
    'bp2': lambda: signal.butter(1, [0.02, 0.4], 'bandpass'),
    'lp2': lambda: signal.butter(2, 0.3),
}

# (n_contacts, chunksize, chunks, last chunk, view offset, row padding)
GEOMETRIES = [
    (3, 16000, 4, 16000, 0, 0),        # everything on the fast path
    (3, 16000, 4, 16000, 1, 7),        # input 2 B off a 16-byte boundary, rows 128-bit aligned or not
    (2, 16000, 3, 16000, 5, 3),
    (2, 16000, 3, 5000, 6, 2),         # short last chunk: other padding, other alignment
    (4, 16016, 3, 1234, 3, 0),         # rows of odd length: contacts 1.. misaligned or not
    (2, 9999, 5, 9999, 2, 1),          # odd chunk size: every other chunk off the 128-byte rows
    (1, 1000000, 2, 1000000, 0, 0),    # the bench geometry, several CTAs per unit
    (2, 600, 3, 600, 0, 0),            # chunks of two tiles: no interior tile at all
    (3, 16000, 4, 6004, 0, 0),         # last chunk with another alignment: its own (element) launch
    (4, 16000, 2, 7500, 0, 4),         # rows of 23 500 + 4 samples (configs[0]-like): per-contact map rows
    (5, 16000, 3, 16000, 3, 5),        # rows of odd length in the PARENT (input misalignment differs per contact)
    (3, 16000, 3, 8002, 1, 0),         # output rows of even, non-multiple-of-16 length
]


@pytest.mark.parametrize("name", sorted(DESIGNS))
@pytest.mark.parametrize("geom", GEOMETRIES)
def test_tma_sweep_equals_element_sweep(env, oracle, name, geom):
    from spikesort_b200._design import tf2sos_exact
    torch, ops = env['torch'], env['ops']
    n_c, cs, n_chunks, last, off, rowpad = geom
    b, a = DESIGNS[name]()
    sos, padlen = tf2sos_exact(b, a)
    n = (n_chunks - 1) * cs + last
    rng = np.random.default_rng(n_c * 1000 + cs + off)
    parent = np.round(rng.standard_normal((n_c, off + n + rowpad)) * 400).astype(np.int16)
    parent[:, off] = 30000                                   # the odd extension wraps around
    view = torch.from_numpy(parent).to(env['dev'])[:, off:off + n]
    with tma_mode(0):
        y0 = ops.filtfilt_sos(view, sos, padlen, cs)
    with tma_mode(1):
        y1 = ops.filtfilt_sos(view, sos, padlen, cs)
    assert torch.equal(y0, y1)
    if n <= 200000:
        x = parent[:, off:off + n]
        ref = np.stack([np.concatenate([oracle.filtfilt(b, a, x[c, lo:min(lo + cs, n)])
                                        for lo in range(0, n, cs)]) for c in range(n_c)])
        assert rel_err(y1.cpu().numpy(), ref) < 1e-12


def test_tma_path_is_taken_and_required_mode_reports(env):
    """SS_APPLY_TMA=2 fails instead of falling back: an output whose tiles start an odd number
    of samples from a 16-byte boundary has no tensor map; the bench geometry has one."""
    from spikesort_b200._design import tf2sos_exact
    torch, ops = env['torch'], env['ops']
    b, a = DESIGNS['hp1']()
    sos, padlen = tf2sos_exact(b, a)
    x = torch.zeros((2, 3 * 16000), dtype=torch.int16, device=env['dev'])
    with tma_mode(2):
        ops.filtfilt_sos(x, sos, padlen, 16000)             # tensor map exists
        # chunk + 2 * padlen odd: the left padding + edge is odd, tiles start on odd elements
        with pytest.raises(Exception):
            ops.filtfilt_sos(x[:, :3 * 15999], sos, padlen, 15999)
        # float input never takes it and is not an error
        ops.filtfilt_sos(x.to(torch.float32), sos, padlen, 16000)
        # rows of any even length and any input alignment have a map (per-contact rows of the
        # tensor map, per-unit realignment): 4 contacts x 40 004 samples, views 1..3 samples in
        parent = torch.zeros((4, 40004 + 8), dtype=torch.int16, device=env['dev'])
        for off in (0, 1, 2, 3):
            ops.filtfilt_sos(parent[:, off:off + 40004], sos, padlen, 16000)


@pytest.mark.parametrize("cs,n", [(16000, 80000), (60000, 200000), (9999, 54995), (16000, 70004),
                                  (16000, 72000)])
def test_tma_fused_filter_detect(env, cs, n):
    """crossing marks fused into the TMA kernel: outputs, thresholds and spike lists as with
    the element kernel, auto and given thresholds, rising and falling."""
    torch, ops = env['torch'], env['ops']
    FS = 1000 if n < 100000 else 30000
    x = env['synth'].make_recording(5, n, FS, seed=12, rate_hz=max(40.0, FS / 100.0), noise=7.0,
                                    amp=(60.0, 160.0)).to(env['dev'])
    flt = env['ss'].filters.Filter(FS * 0.03, FS * 0.004)
    sos, padlen = flt.sections(FS)
    for falling in (True, False):
        res = []
        for mode in (0, 1):
            with tma_mode(mode):
                res.append(ops.filtfilt_detect(x, sos, padlen, cs, FS, falling, thr=None,
                                               n0=int(10 * FS), factor=-1.5 if falling else 1.5))
                low = torch.full((5,), -2.0 if falling else 2.0, dtype=torch.float64, device=env['dev'])
                res.append(ops.filtfilt_detect(x, sos, padlen, cs, FS, falling, thr=low))
        for a_, b_ in ((res[0], res[2]), (res[1], res[3])):
            for t0, t1 in zip(a_, b_):
                assert torch.equal(t0, t1)
        assert res[1][2].numel() > n // 10


@pytest.mark.parametrize("dtype", ["int16", "float32"])
def test_no_padding_last_tile_is_a_whole_tile(env, dtype):
    """padlen = 0 and chunks of a whole number of tiles: the unit's LAST tile holds no extension at
    all, yet its last-output functional is the one the backward pass starts from - sweep 1 leaves
    that functional out for interior tiles, so this tile must not count as one."""
    from spikesort_b200._design import tf2sos_exact
    torch, ops = env['torch'], env['ops']
    b, a = DESIGNS['hp1']()
    sos, _ = tf2sos_exact(b, a)
    rng = np.random.default_rng(3)
    for n, cs in [(4096, 2048), (8192, 8192), (3 * 4096 + 1024, 4096)]:
        x = np.round(rng.standard_normal((2, n)) * 300 + 50)
        x = x.astype(np.int16) if dtype == "int16" else x.astype(np.float32)
        got = ops.filtfilt_sos(torch.from_numpy(x).to(env['dev']), sos, 0, cs).cpu().numpy()
        ref = np.stack([np.concatenate([signal.filtfilt(b, a, x[c, lo:lo + cs].astype(np.float64), padlen=0)
                                        for lo in range(0, n, cs)]) for c in range(2)])
        assert rel_err(got, ref) < 1e-12, (n, cs)
