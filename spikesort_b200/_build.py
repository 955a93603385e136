"""Builds ``libspikesort_b200.so`` (the C-ABI CUDA library) in-tree with nvcc.

    python -m spikesort_b200._build [--force]

Target: sm_100a only (B200).  nvcc cross-compiles without a GPU, so this is also
the "does it build" check the driver runs through ``__graft_entry__.build()``.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libspikesort_b200.so")

SOURCES = ["lib.cu", "filtfilt.cu", "detect.cu", "align.cu", "extract.cu", "pca.cu", "pca_tc.cu", "features.cu", "resample.cu", "exchange.cu", "fir.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "scan.cuh"), os.path.join(CSRC, "internal.cuh"),
           os.path.join(INCLUDE, "spikesort_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + INCLUDE, "-I" + CSRC]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found: cannot build libspikesort_b200.so")
    return nvcc


def _stale(target, deps):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(nvcc, src, obj):
    cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
    return obj


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library."""
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs, objs = [], []
    for name in SOURCES:
        src = os.path.join(CSRC, name)
        obj = os.path.join(OBJ_DIR, name[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + HEADERS):
            jobs.append((src, obj))
    if jobs:
        if verbose:
            print("nvcc: compiling %s" % ", ".join(os.path.basename(s) for s, _ in jobs))
        with concurrent.futures.ThreadPoolExecutor(max_workers=len(jobs)) as pool:
            list(pool.map(lambda j: _compile(nvcc, *j), jobs))
    if jobs or force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
