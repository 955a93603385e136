import torch, time
n = 32*18000000
h = torch.empty(n, dtype=torch.int16).pin_memory()
d = torch.empty(n, dtype=torch.int16, device='cuda')
for k in range(3):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("1 stream: %.2f ms  %.1f GB/s" % (dt*1e3, n*2/dt/1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
half = n//2
for k in range(3):
    torch.cuda.synchronize(); t=time.perf_counter()
    with torch.cuda.stream(s1): d[:half].copy_(h[:half], non_blocking=True)
    with torch.cuda.stream(s2): d[half:].copy_(h[half:], non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print("2 streams: %.2f ms  %.1f GB/s" % (dt*1e3, n*2/dt/1e9))
hp = torch.empty(n, dtype=torch.int16)
torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(hp); torch.cuda.synchronize(); dt=time.perf_counter()-t
print("pageable: %.2f ms  %.1f GB/s" % (dt*1e3, n*2/dt/1e9))
r = torch.empty(25*163770*1, dtype=torch.float32, device='cuda'); hr = torch.empty_like(r, device='cpu').pin_memory()
torch.cuda.synchronize(); t=time.perf_counter(); hr.copy_(r, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t
print("d2h 16MB: %.2f ms" % (dt*1e3))
