"""Does the chain run slower on a side stream than on the default stream?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spikesort_b200 import _ops, filters, pipeline, synth
dev = torch.device("cuda", 0)
FS = 30000
x = synth.make_recording(32, 18000000, FS, seed=2, device=dev, rate_hz=10.0, noise=10.0)
flt = filters.Filter(800., 100.)
orig = _ops.filtfilt_detect
evs = []
def timed(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = orig(*a, **k); e1.record(); evs.append((e0, e1)); return out
_ops.filtfilt_detect = timed
def run(tag, stream, n=12, lazy=True):
    evs.clear()
    ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream())
    with ctx:
        prev = None
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(n):
            if i == 4: t0.record()
            r = pipeline.run_chain(x, FS, filt=flt, sp_win=[-0.2, 0.8], edge='falling', align_type='min', ncomps=2, chunksize=1000000, lazy=lazy)
            if prev is not None: prev.finalize()
            prev = r
        prev.finalize(); t1.record()
    torch.cuda.synchronize()
    st = [a.elapsed_time(b) for a, b in evs[4:]]
    print(tag, "step %.3f ms  stage %.3f ms (min %.3f max %.3f)" % (t0.elapsed_time(t1) / (n - 4), sum(st) / len(st), min(st), max(st)), flush=True)
run("default stream", None)
s = torch.cuda.Stream()
run("side stream   ", s)
run("side stream   ", s)
run("default stream", None)
s2 = torch.cuda.Stream(priority=-1)
run("side hi-prio  ", s2)
run("side hi-prio  ", s2)
