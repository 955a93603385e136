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


# ---- the single-pass span kernel (filt_span_kernel) --------------------------------------
_PATH2 = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools",
                      "proto_filtfilt_span.py")
_spec2 = importlib.util.spec_from_file_location("proto_filtfilt_span", _PATH2)
span = importlib.util.module_from_spec(_spec2)
_spec2.loader.exec_module(span)

SPAN_DESIGNS = {
    'hp1_30k': lambda: signal.iirdesign(800 * 2 / 30000, 100 * 2 / 30000, 1, 10, ftype='butter'),
    'hp1_25k': lambda: signal.iirdesign(800 * 2 / 25000, 100 * 2 / 25000, 1, 10, ftype='butter'),
    'lp3': lambda: signal.butter(3, 0.2),
    'bp4': lambda: signal.butter(2, [0.1, 0.4], 'bandpass'),
}


@pytest.mark.parametrize("name", sorted(SPAN_DESIGNS))
@pytest.mark.parametrize("n,W,K", [(9000, 16, 1), (3000, 3, 1), (512 * 4 + 300, 4, 1), (700, 16, 1),
                                   (512 * 17 - 12 - 40, 16, 1)])
def test_span_formulation_is_the_filter(name, n, W, K):
    """Fast-decaying filters: spans with K warm-up tiles reproduce filtfilt to round-off, with
    a large DC offset in the data (the steady-state warm start removes it from what the
    truncation drops), and the host bound says so."""
    b, a = SPAN_DESIGNS[name]()
    sos, padlen = tf2sos_exact(b, a)
    assert span.span_error_bound(sos, K) <= 1e-17
    rng = np.random.default_rng(n + W)
    x = np.round(rng.standard_normal(n) * 300 + 5000)
    got = span.filtfilt_span(sos, padlen, x, W=W, K=K)
    ref = signal.filtfilt(b, a, x)
    assert np.abs(got - ref).max() <= 1e-12 * max(np.abs(ref).max(), np.abs(x).max())


def test_span_bound_rejects_slow_filters():
    """The order-8 band-pass of configs[1]'s variant (max |pole| 0.98) does not decay within two
    tiles: it must stay on the exact two-sweep scan."""
    b, a = signal.iirdesign(np.array([300, 6000]) * 2 / 30000, np.array([100, 8000]) * 2 / 30000,
                            1, 10, ftype='butter')
    sos, _ = tf2sos_exact(b, a)
    assert span.span_error_bound(sos, 1) > 1e-17 and span.span_error_bound(sos, 2) > 1e-17
    b, a = signal.iirdesign(3 * 2 / 3000, 0.5 * 2 / 3000, 1, 10, ftype='butter')    # pole ~0.994
    sos, _ = tf2sos_exact(b, a)
    assert span.span_error_bound(sos, 2) > 1e-17
