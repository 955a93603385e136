"""Concurrent pinned H2D bandwidth per rank, with and without binding the process to the
CPUs NVML reports as local to its GPU (run under torchrun)."""
import os, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(rank)
N = 32 * 18000000
dev = torch.empty(N, dtype=torch.int16, device="cuda")
def measure(tag):
    host = torch.empty(N, dtype=torch.int16).pin_memory()
    host.fill_(1)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter(); dev.copy_(host, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = min(best, dt)
    gbs = N * 2 / best / 1e9
    out = [None] * world
    dist.all_gather_object(out, (rank, round(gbs, 1), sorted(os.sched_getaffinity(0))[:2], len(os.sched_getaffinity(0))))
    if rank == 0:
        print(tag, out, flush=True)
    del host
measure("default affinity")
try:
    pynvml.nvmlDeviceSetCpuAffinity(h)
    measure("nvml affinity   ")
except Exception as e:
    if rank == 0: print("set affinity failed", e)
# solo bandwidth of rank 0 for reference
if rank == 0:
    host = torch.empty(N, dtype=torch.int16).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); dev.copy_(host, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); dev.copy_(host, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("rank 0 alone: %.1f GB/s" % (N * 2 / dt / 1e9), flush=True)
dist.barrier()
dist.destroy_process_group()
