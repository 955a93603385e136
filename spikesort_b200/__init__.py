"""spikesort_b200 - B200-native drop-in for the front end of btel/SpikeSort.

    from spikesort_b200 import filters, extract, features

mirrors ``spike_sort.core.{filters,extract,features}`` for the path
filter -> detect -> align/extract -> PCA.  ``install()`` rebinds the reference's
own modules to this implementation so unmodified ``spike_beans`` components and
example scripts pick it up.  The arithmetic lives in ``libspikesort_b200.so``
(hand-written sm_100a CUDA behind a C ABI, ``include/spikesort_b200.h``); there is
no CPU fallback.
"""
from . import filters, extract, features            # noqa: F401
from ._signal import DeviceSignal, to_device        # noqa: F401
from .pipeline import sort_frontend                 # noqa: F401

__version__ = "0.1.0"
__all__ = ["filters", "extract", "features", "DeviceSignal", "to_device", "install",
           "sort_frontend"]

_HOT_PATH = {
    "filters": ("Filter", "ZeroPhaseFilter", "FilterFir", "filter_proxy", "fltLinearIIR", "clean_after_exit"),
    "extract": ("detect_spikes", "filter_spt", "extract_spikes", "resample_spikes",
                "align_spikes", "remove_doubles"),
    "features": ("PCA", "fetPCA", "add_mask", "fetP2P", "fetMin", "fetSpIdx", "fetSpTime",
                 "normalize", "combine"),
}


def install(target=None):
    """Rebind the hot-path names of the reference package to this implementation.

    `target` is the imported ``spike_sort`` package (default: ``import spike_sort``);
    ``spike_beans`` looks filters and features up by name at call time
    (``components.py:79,197-198``) and calls ``sort.extract.*`` directly
    (``components.py:132-141,171``), so attribute substitution is enough.
    Returns the dict of replaced originals so the caller can undo it."""
    if target is None:
        import spike_sort as target                 # the reference package
    originals = {}
    mine = {"filters": filters, "extract": extract, "features": features}
    for mod_name, names in _HOT_PATH.items():
        ref_mod = getattr(target, mod_name)
        for name in names:
            if hasattr(ref_mod, name):
                originals[(mod_name, name)] = getattr(ref_mod, name)
            setattr(ref_mod, name, getattr(mine[mod_name], name))
    return originals
