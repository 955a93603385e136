"""Traffic-shape-matched bandwidth ceiling for the filter's second sweep: read int16, write
float64 (2 B in, 8 B out per sample), nothing else - torch's cast-copy kernel and a plain
copy for comparison."""
import torch
N = 32 * 18000000
x = torch.randint(-1000, 1000, (N,), dtype=torch.int16, device="cuda")
out = torch.empty(N, dtype=torch.float64, device="cuda")
a = torch.empty(N * 5 // 8, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t = timeit(lambda: out.copy_(x))
print("int16 -> float64 cast copy: %.3f ms, %.0f GB/s (2 B read + 8 B written per sample)" % (t, N * 10 / t / 1e6))
t = timeit(lambda: b.copy_(a))
print("float64 copy of the same total bytes: %.3f ms, %.0f GB/s" % (t, a.numel() * 16 / t / 1e6))
t = timeit(lambda: out.fill_(1.0))
print("float64 fill (write only): %.3f ms, %.0f GB/s" % (t, N * 8 / t / 1e6))
