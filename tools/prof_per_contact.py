"""Where the per-contact drop-in calls (what install() users run) spend their time (GPU box)."""
import cProfile, os, pstats, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spikesort_b200 import synth, filters, extract, features
FS, C, N = 30000, 32, 18000000
x = synth.make_recording(C, N, FS, seed=2, device="cuda", rate_hz=10.0, noise=10.0)
host = torch.empty((C, N), dtype=torch.int16).pin_memory(); host.copy_(x); torch.cuda.synchronize()
raw = {'data': host, 'FS': FS, 'n_contacts': C}
SP_WIN = [-0.2, 0.8]

def step(timers=None):
    def tick(name, t0):
        if timers is not None:
            torch.cuda.synchronize()
            timers[name] = timers.get(name, 0.0) + time.perf_counter() - t0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter(); f = filters.fltLinearIIR(raw, 800., 100.); tick('filter', t0)
        for c in range(C):
            t0 = time.perf_counter(); spt = extract.detect_spikes(f, thresh='auto', edge='falling', contact=c); tick('detect', t0)
            t0 = time.perf_counter(); spa = extract.align_spikes(f, spt, SP_WIN, type='min', contact=c); tick('align', t0)
            t0 = time.perf_counter(); wv = extract.extract_spikes(f, spa, SP_WIN, contacts=c); tick('extract', t0)
            t0 = time.perf_counter(); ft = features.fetPCA(wv, ncomps=2); tick('fetPCA', t0)

step(); step()
torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize()
print("per-contact step %.2f ms" % (1e3 * (time.perf_counter() - t0)))
tm = {}
step(tm)
print({k: round(1e3 * v, 2) for k, v in tm.items()})
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
