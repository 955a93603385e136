"""Host-logic tier: the NumPy statement (tools/proto_filtfilt_sections.py) of K1's algebra as
built - section cascade, constant left padding to whole tiles, tile summaries, tile scan with
the M term, and the SECTION-BY-SECTION second sweep (lane-local pass, lane scan with the
section's 2x2 powers, correction) - against the sequential cascade and scipy.signal.filtfilt
(what Filter.__call__ runs, filters.py:99-101).  The GPU tier checks the kernels themselves."""
import importlib.util
import os

import numpy as np
import pytest
from scipy import signal

from spikesort_b200._design import tf2sos_exact

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools",
                     "proto_filtfilt_sections.py")
_spec = importlib.util.spec_from_file_location("proto_filtfilt_sections", _PATH)
proto = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(proto)

DESIGNS = {
    'hp1': (lambda: signal.iirdesign(800 * 2 / 25000, 100 * 2 / 25000, 1, 10, ftype='butter'), 1e-12),
    'lp3': (lambda: signal.butter(3, 0.2), 1e-12),
    'bp4': (lambda: signal.butter(2, [300 / 15000, 6000 / 15000], 'bandpass'), 1e-10),
    'bp8': (lambda: signal.iirdesign(np.array([300, 6000]) * 2 / 30000,
                                     np.array([100, 8000]) * 2 / 30000, 1, 10, ftype='butter'), 1e-8),
}


@pytest.mark.parametrize("name", sorted(DESIGNS))
@pytest.mark.parametrize("n,S", [(700, 16), (1500, 16), (333, 4)])
def test_tiled_sectionwise_scan_is_the_filter(name, n, S):
    design, bound = DESIGNS[name]
    b, a = design()
    sos, padlen = tf2sos_exact(b, a)
    assert padlen == 3 * max(len(a), len(b))
    rng = np.random.default_rng(n + S)
    x = np.round(rng.standard_normal(n) * 300 + 40)
    seq = proto.filtfilt_sequential(sos, padlen, x)
    tiled = proto.filtfilt_tiled(sos, padlen, x, S=S)
    ref = signal.filtfilt(b, a, x)
    scale = np.abs(ref).max()
    assert np.abs(tiled - seq).max() <= 1e-12 * scale          # the scan is the cascade
    assert np.abs(tiled - ref).max() <= bound * scale          # and the cascade is filtfilt(b, a)
