"""Probe: torch symmetric memory (peer-addressable buffers over NVLink) on this box."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm
t = symm.empty((1024,), dtype=torch.float64, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs], flush=True)
t.fill_(float(rank))
hdl.barrier(channel=0)
peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float64)
val = peer[:4].clone()
torch.cuda.synchronize()
print(rank, "peer read", val.tolist(), flush=True)
# write into the peer
peer[8:12] = 100.0 + rank
hdl.barrier(channel=0)
torch.cuda.synchronize()
print(rank, "own after peer write", t[8:12].tolist(), flush=True)
# latency of barrier
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    hdl.barrier(channel=0)
e1.record(); torch.cuda.synchronize()
print(rank, "barrier us", e0.elapsed_time(e1) * 10, flush=True)
t0 = time.perf_counter()
for _ in range(100):
    hdl.barrier(channel=0)
print(rank, "barrier cpu us", (time.perf_counter() - t0) * 1e4, flush=True)
torch.cuda.synchronize()
# NCCL p2p exchange latency for comparison
a = torch.zeros(32 * 120, dtype=torch.float64, device=dev); b = torch.empty_like(a)
def xchg():
    ops = [dist.P2POp(dist.isend, a, (rank + 1) % world), dist.P2POp(dist.irecv, b, (rank - 1) % world)]
    for r in dist.batch_isend_irecv(ops): r.wait()
for _ in range(5): xchg()
torch.cuda.synchronize()
e0.record()
for _ in range(50): xchg()
e1.record(); torch.cuda.synchronize()
print(rank, "nccl sendrecv us", e0.elapsed_time(e1) * 20, flush=True)
x = torch.zeros(30000, dtype=torch.float64, device=dev)
for _ in range(5): dist.all_reduce(x)
torch.cuda.synchronize()
e0.record()
for _ in range(50): dist.all_reduce(x)
e1.record(); torch.cuda.synchronize()
print(rank, "nccl allreduce 240KB us", e0.elapsed_time(e1) * 20, flush=True)
print(rank, "p2p access", torch.cuda.can_device_access_peer(rank, (rank + 1) % world), flush=True)
dist.destroy_process_group()
