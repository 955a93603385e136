"""ctypes binding of ``libspikesort_b200.so`` (``include/spikesort_b200.h``).

This is the binding a SpikeSort maintainer would add.  There
is NO CPU fallback: if the library is missing or no CUDA device is present the
calls raise.  PyTorch is used only to own device buffers and streams.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPIKESORT_B200_LIB: another build of the same library (kernel A/B runs)
LIB_PATH = os.environ.get("SPIKESORT_B200_LIB") or os.path.join(_HERE, "libspikesort_b200.so")

SS_I16, SS_F32, SS_F64 = 0, 1, 2
SS_GRAM_AUTO, SS_GRAM_F64, SS_GRAM_TC = 0, 1, 2
SS_ERR_PADLEN = -3
SS_ERR_UNSUPPORTED = -5

_c = ctypes
_i64, _int, _dbl, _vp, _sz = _c.c_int64, _c.c_int, _c.c_double, _c.c_void_p, _c.c_size_t
_dp = _c.POINTER(_c.c_double)

# name -> (restype, argtypes); must list EVERY symbol declared in the header
SIGNATURES = {
    "ss_version": (_int, []),
    "ss_last_error": (_c.c_char_p, []),
    "ss_launch_count": (_i64, []),
    "ss_max_filter_order": (_int, []),
    "ss_max_window": (_int, []),
    "ss_max_pca_window": (_int, []),
    "ss_filtfilt_workspace_bytes": (_sz, [_i64, _i64, _i64, _int, _int]),
    "ss_filtfilt_sos": (_int, [_vp, _int, _i64, _i64, _i64, _dp, _int, _int, _i64, _vp, _i64,
                               _vp, _sz, _vp]),
    "ss_filtfilt_detect_sos": (_int, [_vp, _int, _i64, _i64, _i64, _dp, _int, _int, _i64, _vp, _i64,
                                      _vp, _sz, _int, _i64, _dbl, _vp, _int, _vp, _sz, _vp, _int,
                                      _vp]),
    "ss_filtfilt_detect_sos_peer": (_int, [_vp, _int, _i64, _i64, _i64, _dp, _int, _int, _i64, _vp, _i64,
                                           _vp, _sz, _int, _i64, _dbl, _vp, _int, _vp, _sz, _vp, _int,
                                           _vp, _vp, _int, _vp]),
    "ss_filtfilt_prepare_sos": (_int, [_vp, _int, _i64, _i64, _i64, _dp, _int, _int, _i64, _vp, _sz,
                                       _vp]),
    "ss_filtfilt_head_threshold_sos": (_int, [_vp, _int, _i64, _i64, _i64, _dp, _int, _int, _i64,
                                              _vp, _i64, _vp, _sz, _i64, _dbl, _vp, _vp, _sz, _vp]),
    "ss_threshold_workspace_bytes": (_sz, [_i64]),
    "ss_var_threshold": (_int, [_vp, _int, _i64, _i64, _i64, _dbl, _vp, _vp, _sz, _vp]),
    "ss_detect_workspace_bytes": (_sz, [_i64, _i64]),
    "ss_detect_mark": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _int, _vp, _sz, _vp, _vp]),
    "ss_detect_emit": (_int, [_i64, _i64, _dbl, _i64, _vp, _sz, _vp, _i64, _vp]),
    "ss_align": (_int, [_vp, _int, _i64, _i64, _dbl, _vp, _i64, _vp, _dbl, _int, _int, _dbl,
                        _dbl, _dbl, _dbl, _int, _vp, _i64, _int, _i64, _i64, _i64, _vp, _vp]),
    "ss_remove_doubles_workspace_bytes": (_sz, [_i64]),
    "ss_remove_doubles_mark": (_int, [_vp, _i64, _vp, _i64, _dbl, _vp, _vp, _sz, _vp, _vp]),
    "ss_remove_doubles_emit": (_int, [_vp, _i64, _vp, _sz, _vp, _i64, _vp]),
    "ss_extract": (_int, [_vp, _int, _i64, _i64, _dbl, _vp, _int, _vp, _i64, _vp, _int, _int,
                          _dbl, _dbl, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "ss_pca_workspace_bytes": (_sz, [_int, _i64]),
    "ss_pca_gram": (_int, [_vp, _int, _int, _i64, _int, _vp, _i64, _vp, _vp, _vp, _int, _vp, _sz, _vp]),
    "ss_pca_eig": (_int, [_vp, _vp, _vp, _int, _i64, _vp, _vp, _vp]),
    "ss_pca_eig_top": (_int, [_vp, _vp, _vp, _int, _i64, _int, _vp, _vp, _vp]),
    "ss_pca_project": (_int, [_vp, _int, _int, _i64, _int, _vp, _i64, _vp, _vp, _int, _vp, _vp]),
    "ss_wave_minmax": (_int, [_vp, _int, _int, _i64, _int, _vp, _vp, _vp]),
    "ss_normalize_workspace_bytes": (_sz, [_int]),
    "ss_normalize_columns": (_int, [_vp, _i64, _int, _vp, _sz, _vp]),
    "ss_copy_2d_async": (_int, [_vp, _sz, _vp, _sz, _sz, _sz, _int, _vp]),
    "ss_filtfilt_fir": (_int, [_vp, _int, _i64, _i64, _i64, _vp, _int, _i64, _vp, _i64, _vp]),
    "ss_shard_arena_bytes": (_sz, [_int, _int, _int, _int]),
    "ss_shard_arena_offsets": (_int, [_int, _int, _int, _int, _vp]),
    "ss_peer_barrier": (_int, [_vp, _int, _int, _c.c_uint64, _dbl, _vp]),
    "ss_shard_put_thresholds": (_int, [_vp, _int, _int, _int, _vp, _int, _vp]),
    "ss_shard_put_halos": (_int, [_vp, _i64, _int, _i64, _int, _int, _vp, _int, _int, _vp]),
    "ss_shard_install_halos": (_int, [_vp, _i64, _int, _i64, _int, _int, _vp, _int, _int, _vp, _int,
                                      _vp, _vp]),
    "ss_shard_append_boundary": (_int, [_vp, _vp, _int, _i64, _vp, _dbl, _vp, _vp, _vp]),
    "ss_shard_put_last": (_int, [_vp, _vp, _int, _int, _int, _vp, _int, _int, _vp]),
    "ss_shard_gram_put": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _vp, _int, _int, _vp, _vp, _vp]),
    "ss_shard_gram_reduce": (_int, [_vp, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ss_merge_shard": (_int, [_vp, _vp, _vp, _int, _vp, _int, _i64, _i64, _vp, _int, _vp, _vp, _vp, _vp, _vp,
                              _vp, _vp, _vp]),
    "ss_pca_first_wave": (_int, [_vp, _int, _int, _i64, _int, _vp, _i64, _vp, _vp]),
    "ss_resample": (_int, [_vp, _int, _int, _i64, _int, _vp, _vp, _int, _vp, _vp]),
    "ss_align_resampled": (_int, [_vp, _int, _i64, _i64, _dbl, _vp, _i64, _vp, _int, _int, _dbl,
                                  _dbl, _dbl, _dbl, _int, _vp, _vp, _int, _vp, _vp, _i64, _int,
                                  _i64, _i64, _i64, _vp]),
}

_lib = None


class SpikeSortB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SpikeSortB200Error(
            "%s not found: build it with `python -m spikesort_b200._build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.ss_version() != 100:
        raise SpikeSortB200Error("libspikesort_b200.so version mismatch")
    _lib = lib
    return lib


def last_error():
    return load().ss_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc == 0:
        return
    msg = last_error()
    if rc == SS_ERR_PADLEN:
        raise ValueError(msg)                 # what scipy.signal.filtfilt raises
    if rc == SS_ERR_UNSUPPORTED:
        raise NotImplementedError("%s: %s" % (what, msg))
    raise SpikeSortB200Error("%s failed (%d): %s" % (what, rc, msg))


# --------------------------------------------------------------------------
# torch glue: device buffers and streams only
# --------------------------------------------------------------------------

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise SpikeSortB200Error("spikesort_b200 needs a CUDA device (B200); there is no "
                                 "CPU fallback")
    return t


def device():
    t = require_cuda()
    return t.device("cuda", t.cuda.current_device())


def stream_ptr():
    return ctypes.c_void_p(torch().cuda.current_stream().cuda_stream)


def ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr()) if tensor is not None else ctypes.c_void_p(0)


_NP2SS = {np.dtype(np.int16): SS_I16, np.dtype(np.float32): SS_F32, np.dtype(np.float64): SS_F64}


def ss_dtype_of(tensor):
    t = torch()
    return {t.int16: SS_I16, t.float32: SS_F32, t.float64: SS_F64}[tensor.dtype]


def np_dtype_of(tensor):
    t = torch()
    return {t.int16: np.int16, t.float32: np.float32, t.float64: np.float64}[tensor.dtype]


def np_dtype_full(torch_dtype):
    """NumPy dtype of any torch dtype this package moves to the host."""
    t = torch()
    return np.dtype({t.int16: np.int16, t.float32: np.float32, t.float64: np.float64,
                     t.int64: np.int64, t.int32: np.int32, t.uint8: np.uint8}[torch_dtype])


_STAGE = {}                      # device index -> two pinned uint8 staging buffers
_STAGE_LOCK = __import__("threading").Lock()   # the staging buffers serve one upload at a time


def upload(host):
    """CPU torch tensor -> CUDA tensor of the same shape.  Pinned and small tensors are one
    (asynchronous) copy.  A big PAGEABLE tensor - the usual NumPy array - would be staged by the
    CUDA runtime itself at a fraction of the PCIe rate; it goes through two pinned 32 MB buffers
    instead: a few threads copy chunk k + 1 into one of them (NumPy's copy releases the GIL) while
    the DMA of chunk k runs out of the other.  SS_PAGEABLE_STAGING=0 switches that off."""
    t = torch()
    nbytes = host.numel() * host.element_size()
    if host.is_cuda or nbytes < (64 << 20) or not host.is_contiguous() or host.is_pinned() or \
            os.environ.get("SS_PAGEABLE_STAGING", "1")[:1] == "0":
        return host.to(device(), non_blocking=True)
    with _STAGE_LOCK:
        return _staged_upload(t, host, nbytes)


def _staged_upload(t, host, nbytes):
    import concurrent.futures
    dev = t.empty(host.shape, dtype=host.dtype, device=device())
    src = host.reshape(-1).view(t.uint8).numpy()
    dst = dev.reshape(-1).view(t.uint8)
    chunk = 32 << 20
    key = t.cuda.current_device()
    if key not in _STAGE:
        _STAGE[key] = [t.empty(chunk, dtype=t.uint8).pin_memory() for _ in range(2)]
    bufs = _STAGE[key]
    views = [b.numpy() for b in bufs]
    n_thr = int(os.environ.get("SS_STAGING_THREADS", "0")) or max(1, min(8, (os.cpu_count() or 2) // 2))
    done = [None, None]
    with concurrent.futures.ThreadPoolExecutor(max_workers=n_thr) as pool:
        for i, off in enumerate(range(0, nbytes, chunk)):
            slot = i & 1
            n = min(chunk, nbytes - off)
            if done[slot] is not None:
                done[slot].synchronize()                  # the DMA out of this buffer has finished
            part = (n + n_thr - 1) // n_thr

            def copy(p0, slot=slot, off=off, n=n, part=part):
                p1 = min(p0 + part, n)
                np.copyto(views[slot][p0:p1], src[off + p0:off + p1])

            list(pool.map(copy, range(0, n, part)))
            dst[off:off + n].copy_(bufs[slot][:n], non_blocking=True)
            ev = t.cuda.Event()
            ev.record()
            done[slot] = ev
    for ev in done:                                       # the staging buffers are reused by the next call
        if ev is not None:
            ev.synchronize()
    return dev


def to_device_2d(data):
    """Recording (n_contacts, n_samples) -> CUDA tensor of a supported dtype with
    unit stride along time.  Accepts DeviceSignal, torch tensors and anything
    NumPy can view (ndarray, memmap, PyTables-like with __getitem__)."""
    t = require_cuda()
    if hasattr(data, "device_tensor"):
        return data.device_tensor()
    if isinstance(data, t.Tensor):
        x = data
    else:
        arr = np.asarray(data[:] if not isinstance(data, np.ndarray) else data)
        if arr.dtype not in _NP2SS:
            # NumPy promotes every other real dtype to float64 before filtering/comparing
            arr = arr.astype(np.float64)
        x = t.from_numpy(np.ascontiguousarray(arr))
    if x.dtype not in (t.int16, t.float32, t.float64):
        x = x.to(t.float64)
    if x.dim() == 1:
        x = x.unsqueeze(0)
    x = upload(x)
    if x.stride(-1) != 1:
        x = x.contiguous()
    return x


def empty(shape, dtype):
    return torch().empty(shape, dtype=dtype, device=device())


def workspace(nbytes):
    t = torch()
    return t.empty(max(int(nbytes), 8), dtype=t.uint8, device=device())
