"""Where the host-in/host-out front end spends its time (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spikesort_b200 as ss
from spikesort_b200 import synth, filters
FS, C, N = 30000, 32, 18000000
x = synth.make_recording(C, N, FS, seed=2, device="cuda", rate_hz=10.0, noise=10.0)
host = torch.empty((C, N), dtype=torch.int16).pin_memory(); host.copy_(x); torch.cuda.synchronize()
raw = {'data': host, 'FS': FS, 'n_contacts': C}
flt = filters.Filter(800., 100.)
kw = dict(filt=flt, sp_win=[-0.2, 0.8], edge='falling', align_type='min', thresh='auto', ncomps=2, chunksize=1000000)
# raw upload speed, contiguous and 2-D
d = torch.empty_like(x)
for _ in range(2):
    t0 = time.perf_counter(); d.copy_(host, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print("contiguous H2D %.2f ms (%.1f GB/s)" % (1e3 * (t1 - t0), host.numel() * 2 / (t1 - t0) / 1e9))
from spikesort_b200 import _lib
lib = _lib.load()
for _ in range(2):
    t0 = time.perf_counter()
    for a in range(0, N, 3000000):
        lib.ss_copy_2d_async(d.data_ptr() + 2 * a, N * 2, host.data_ptr() + 2 * a, N * 2, 3000000 * 2, C, 1, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize(); t1 = time.perf_counter()
print("6 x 2-D H2D   %.2f ms (%.1f GB/s)" % (1e3 * (t1 - t0), host.numel() * 2 / (t1 - t0) / 1e9))
for slabs in (1, 'auto', 3, 9):
    for rep in range(3):
        if rep == 2:
            os.environ["SS_E2E_PROFILE"] = "1"
        t0 = time.perf_counter(); out = ss.sort_frontend(raw, slabs=slabs, **kw); torch.cuda.synchronize(); t1 = time.perf_counter()
        os.environ.pop("SS_E2E_PROFILE", None)
    print("slabs=%s: %.2f ms" % (slabs, 1e3 * (t1 - t0)), flush=True)
