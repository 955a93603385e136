"""TEST INFRASTRUCTURE - loader for the UNMODIFIED reference hot-path files.

Works only where ``/root/reference`` exists (the build container); the GPU box
has no reference tree, so nothing under ``tests -m gpu``, ``smoke()`` or
``bench.py`` may import this module.  It is used by

* ``tests/golden/make_golden.py`` to generate the committed golden vectors, and
* ``tests/test_oracle_vs_reference.py`` (skipped when the tree is absent) to pin
  the numpy/C restatement in ``oracle/port.py`` against the real reference.

The reference is Python 2 code (``setup.py:39``); its three hot-path modules
(``src/spike_sort/core/{filters,extract,features}.py``) load under Python 3 once
three names exist:

1. ``builtins.basestring``   (``extract.py:73``)
2. a ``tables`` module with ``Atom.from_dtype`` and
   ``openFile().createCArray()`` (``filters.py:128-131``) - here NumPy-backed,
   in RAM (no PyTables in this image),
3. a ``spike_sort`` / ``spike_sort.stats`` stub package (``features.py:23``).

No reference source is copied: the files are executed from where they lie.
"""
import builtins
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SPIKESORT_REFERENCE", "/root/reference")
_CORE = os.path.join(REFERENCE_ROOT, "src", "spike_sort", "core")


def available():
    return os.path.isfile(os.path.join(_CORE, "extract.py"))


class _FakeCArray(np.ndarray):
    """ndarray that can carry the ``_v_file`` attribute PyTables nodes have."""


class _FakeFile(object):
    def __init__(self, filename, mode):
        self.filename = filename
        self._arrays = []
        # the reference deletes the temp file at exit (filters.py:174-181)
        open(filename, "wb").close()

    def createCArray(self, where, name, atom, shape):
        arr = np.zeros(shape, dtype=atom.dtype).view(_FakeCArray)
        arr._v_file = self
        self._arrays.append(arr)
        return arr

    def close(self):
        self._arrays = []


class _FakeAtom(object):
    def __init__(self, dtype):
        self.dtype = dtype

    @classmethod
    def from_dtype(cls, dtype):
        return cls(np.dtype(dtype))


def _install_shims():
    if not hasattr(builtins, "basestring"):
        builtins.basestring = str
    if not hasattr(builtins, "reduce"):          # features.combine, features.py:120
        import functools
        builtins.reduce = functools.reduce
    if "tables" not in sys.modules:
        tables = types.ModuleType("tables")
        tables.Atom = _FakeAtom
        tables.openFile = _FakeFile
        tables.__spikesort_shim__ = True
        sys.modules["tables"] = tables
    if "spike_sort" not in sys.modules:
        pkg = types.ModuleType("spike_sort")
        pkg.__path__ = []
        pkg.__spikesort_shim__ = True
        stats = types.ModuleType("spike_sort.stats")
        pkg.stats = stats
        sys.modules["spike_sort"] = pkg
        sys.modules["spike_sort.stats"] = stats


_cache = {}


def load(name):
    """Return the reference module ``spike_sort.core.<name>`` executed in place."""
    if name in _cache:
        return _cache[name]
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_shims()
    path = os.path.join(_CORE, name + ".py")
    spec = importlib.util.spec_from_file_location("_ref_spike_sort_core_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod


def filters():
    """core/filters.py.  FilterFir calls signal.remez(..., Hz=FS) (filters.py:66); SciPy
    renamed that keyword to fs, so the module's `signal.remez` is wrapped to translate it -
    same routine, same arguments."""
    mod = load("filters")
    if not getattr(mod, "_remez_shimmed", False):
        import scipy.signal as _sig
        _orig = _sig.remez

        def remez(numtaps, bands, desired, *args, **kw):
            if "Hz" in kw:
                kw["fs"] = kw.pop("Hz")
            return _orig(numtaps, bands, desired, *args, **kw)

        shim = types.ModuleType("scipy_signal_shim")
        shim.__dict__.update(_sig.__dict__)
        shim.remez = remez
        mod.signal = shim
        mod._remez_shimmed = True
    return mod


def extract():
    return load("extract")


def features():
    return load("features")
