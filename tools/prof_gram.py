"""Profiling driver: Gram of (25, 2M, 4) waveforms, tensor-core path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SS_PCA_TC"] = sys.argv[1] if len(sys.argv) > 1 else "1"
import torch
from spikesort_b200 import _ops, synth
w = synth.make_waveforms(25, 2000000, 4, seed=5, device="cuda")
for _ in range(3):
    g = _ops.pca_gram(w)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g = _ops.pca_gram(w); e1.record(); torch.cuda.synchronize()
print("gram ms", e0.elapsed_time(e1))
