"""sort_frontend on a PAGEABLE NumPy recording: threaded pinned staging against the runtime's own
staging of 2-D copies out of pageable memory (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import spikesort_b200 as ss
from spikesort_b200 import synth, filters
FS, C, N = 30000, 32, 18000000
x = synth.make_recording(C, N, FS, seed=2, device="cuda", rate_hz=10.0, noise=10.0)
arr = x.cpu().numpy().copy()                       # pageable
pinned = torch.empty((C, N), dtype=torch.int16).pin_memory(); pinned.copy_(x); torch.cuda.synchronize()
flt = filters.Filter(800., 100.)
kw = dict(filt=flt, sp_win=[-0.2, 0.8], edge='falling', align_type='min', thresh='auto', ncomps=2, chunksize=1000000)
ref = ss.sort_frontend({'data': pinned, 'FS': FS, 'n_contacts': C}, **kw)
from spikesort_b200 import pipeline
for mode, th in (("0", 4), ("1", 4), ("1", 8), ("1", 16), ("1", 2), ("1", 8)):
    os.environ["SS_PAGEABLE_STAGING"] = mode
    pipeline._PageableSource.THREADS = th
    out = None
    for rep in range(2):
        t0 = time.perf_counter(); out = ss.sort_frontend({'data': arr, 'FS': FS, 'n_contacts': C}, **kw)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    same = all(np.array_equal(out['spt_aligned'][c]['data'], ref['spt_aligned'][c]['data']) and
               np.array_equal(out['sp_waves'][c]['data'], ref['sp_waves'][c]['data']) and
               np.array_equal(out['features'][c]['data'], ref['features'][c]['data']) for c in range(C))
    print("pageable recording, SS_PAGEABLE_STAGING=%s, %d threads: %.1f ms, identical to the pinned run: %s" % (mode, th, 1e3 * (t1 - t0), same), flush=True)
t0 = time.perf_counter(); ss.sort_frontend({'data': pinned, 'FS': FS, 'n_contacts': C}, **kw); torch.cuda.synchronize()
print("pinned recording: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
