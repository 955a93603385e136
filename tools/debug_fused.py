"""Debug helper: the (500, 9000, 2000) case of test_fused_filter_detect_equals_separate_passes,
one library call at a time with a device synchronisation after each."""
import os
import sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spikesort_b200
from spikesort_b200 import _ops as ops, _lib, synth

_lib.load()
dev = torch.device("cuda:0")
FS, n, cs = 500, 9000, 2000


def step(name, fn):
    try:
        r = fn()
        torch.cuda.synchronize()
        print("ok  ", name, flush=True)
        return r
    except Exception as e:                                   # noqa: BLE001
        print("FAIL", name, repr(e)[:300], flush=True)
        print("last_error:", _lib.last_error(), flush=True)
        sys.exit(1)


x = synth.make_recording(5, n, FS, seed=9, rate_hz=max(40.0, FS / 100.0), noise=7.0,
                         amp=(60.0, 160.0)).to(dev)
x[3] = 0
flt = spikesort_b200.filters.Filter(FS * 0.03, FS * 0.004)
sos, padlen = flt.sections(FS)
print("sos", sos, "padlen", padlen)
for falling in (True, False):
    y = step("filtfilt", lambda: ops.filtfilt_sos(x, sos, padlen, cs))
    thr = step("var", lambda: ops.var_threshold(y, int(10 * FS), -1.5 if falling else 1.5))
    spt, off = step("detect", lambda: ops.detect(y, thr, falling, FS))
    step("fused auto", lambda: ops.filtfilt_detect(x, sos, padlen, cs, FS, falling, thr=None,
                                                   n0=int(10 * FS), factor=-1.5 if falling else 1.5))
    step("fused thr", lambda: ops.filtfilt_detect(x, sos, padlen, cs, FS, falling, thr=thr))
    low = torch.full((5,), -2.0 if falling else 2.0, dtype=torch.float64, device=dev)
    spt4, off4 = step("detect low", lambda: ops.detect(y, low, falling, FS))
    print("n low", spt4.numel(), off4.tolist())
    y5, thr5, spt5, off5 = step("fused low", lambda: ops.filtfilt_detect(x, sos, padlen, cs, FS, falling, thr=low))
    print("equal", torch.equal(spt4, spt5), torch.equal(off4, off5), torch.equal(y, y5))
print("done")
