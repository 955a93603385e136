"""The per-contact API on PAGEABLE NumPy inputs: filter_proxy of a recording and fetPCA of a
waveform block, with and without the staged upload (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from spikesort_b200 import synth, filters, features
FS, C, N = 30000, 32, 18000000
arr = synth.make_recording(C, N, FS, seed=2, device="cuda", rate_hz=10.0, noise=10.0).cpu().numpy().copy()
w = (synth.make_waveforms(25, 4000000, 4, seed=3, device="cuda") + 1.0).cpu().numpy().copy()
for mode in ("0", "1", "0", "1"):
    os.environ["SS_PAGEABLE_STAGING"] = mode
    t0 = time.perf_counter(); f = filters.fltLinearIIR({'data': arr, 'FS': FS, 'n_contacts': C}, 800., 100.); torch.cuda.synchronize(); t1 = time.perf_counter()
    t2 = time.perf_counter(); ft = features.fetPCA({'data': w}, ncomps=2); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("SS_PAGEABLE_STAGING=%s: fltLinearIIR on a 1.15 GB NumPy recording %.1f ms, fetPCA on a 1.6 GB NumPy block %.1f ms (checksum %.6f)" % (
        mode, 1e3 * (t1 - t0), 1e3 * (t3 - t2), float(np.abs(ft['data']).sum())), flush=True)
    del f
