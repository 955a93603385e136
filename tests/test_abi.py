"""CPU tier: the C-ABI library loads and exports exactly what include/spikesort_b200.h
declares; argument validation that needs no GPU; the Python binding covers the header."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "spikesort_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from spikesort_b200 import _build, _lib
    _build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n


def test_binding_covers_header(lib):
    from spikesort_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_version_and_limits(lib):
    assert lib.ss_version() == 100
    assert lib.ss_max_filter_order() >= 12
    assert lib.ss_max_window() >= 64


def test_workspace_queries(lib):
    assert lib.ss_filtfilt_workspace_bytes(4, 1000000, 1000000, 6, 1) > 0
    assert lib.ss_filtfilt_workspace_bytes(0, 10, 10, 6, 1) == 0
    assert lib.ss_detect_workspace_bytes(2, 100000) >= 2 * 100000 // 8
    assert lib.ss_threshold_workspace_bytes(3) > 0
    assert lib.ss_remove_doubles_workspace_bytes(0) > 0
    assert lib.ss_pca_workspace_bytes(30, 4) > 0
    assert lib.ss_pca_workspace_bytes(1000, 4) == 0


def test_argument_errors_without_gpu(lib):
    from spikesort_b200 import _lib
    sos = (ctypes.c_double * 6)(1, -1, 0, 1, -0.9, 0)
    one = ctypes.c_void_p(16)                      # never dereferenced: validation fails first
    rc = lib.ss_filtfilt_sos(one, 0, 1, 5, 5, sos, 1, 6, 1000, one, 5, one, 1 << 20, None)
    assert rc == _lib.SS_ERR_PADLEN
    assert "greater than padlen" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "ss_filtfilt_sos")
    rc = lib.ss_filtfilt_sos(one, 0, 1, 500, 500, sos, 9, 6, 1000, one, 500, one, 1 << 20, None)
    assert rc == _lib.SS_ERR_UNSUPPORTED
    rc = lib.ss_filtfilt_sos(one, 0, 1, 500, 500, sos, 1, 6, 1000, one, 500, one, 8, None)
    assert rc == -4                                # workspace too small
    rc = lib.ss_align(one, 7, 100, 100, 25000.0, None, 1, None, -0.2, -5, 20, 0.2, 3.0, -0.1, 0.7,
                      0, one, 3, 10, 0, 0, 0, None, None)
    assert rc == -1


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    from spikesort_b200 import extract, _lib
    with pytest.raises(_lib.SpikeSortB200Error):
        extract.detect_spikes({'data': np.zeros((1, 100)), 'FS': 1000, 'n_contacts': 1}, thresh=1.0)
