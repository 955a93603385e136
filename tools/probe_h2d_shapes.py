"""Is the e2e collapse at N >= 4 the box or the copy shape?  Under torchrun, every rank uploads
its 1.15 GB pinned shard (32 x 18,000,000 int16) three ways and reports GB/s, all ranks at once:
  contiguous : one cudaMemcpyAsync of the whole buffer (the ceiling of this box for N uploads)
  slabs-2d   : what sort_frontend does - six strided 2-D copies (ss_copy_2d_async) of the
               (contacts x time-slab) blocks of the row-major recording
  staggered  : the same slabs, rank r starting r/N of one upload time later
and rank 0 alone for reference."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spikesort_b200 import _lib                                  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
lib = _lib.load()
C, N, S = 32, 18000000, 6
host = torch.empty((C, N), dtype=torch.int16).pin_memory()
host.fill_(1)
dev = torch.empty((C, N), dtype=torch.int16, device="cuda")
st = torch.cuda.current_stream()


def contiguous():
    dev.copy_(host, non_blocking=True)


def slabs():
    per = N // S
    for k in range(S):
        a, b = k * per, (N if k == S - 1 else (k + 1) * per)
        rc = lib.ss_copy_2d_async(dev.data_ptr() + 2 * a, N * 2, host.data_ptr() + 2 * a, N * 2,
                                  (b - a) * 2, C, 1, st.cuda_stream)
        assert rc == 0


def run(tag, fn, alone=False, delay=0.0):
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        if not alone:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if delay:
            time.sleep(delay)
        t1 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t1)
        wall = time.perf_counter() - t0
    gbs = C * N * 2 / best / 1e9
    if alone:
        print("%-28s rank 0 alone: %.1f GB/s" % (tag, gbs), flush=True)
        return gbs
    out = [None] * world
    dist.all_gather_object(out, (round(gbs, 1), round(wall * 1e3, 1)))
    if rank == 0:
        print("%-28s per-rank GB/s %s  aggregate %.0f GB/s  slowest rank done after %.1f ms" % (
            tag, [o[0] for o in out], sum(o[0] for o in out), max(o[1] for o in out)), flush=True)
    return gbs


g0 = run("contiguous, all ranks", contiguous)
run("slabs-2d, all ranks", slabs)
t_one = C * N * 2 / (g0 * 1e9)
run("slabs-2d staggered", slabs, delay=rank * t_one / world)
dist.barrier()
if rank == 0:
    run("contiguous", contiguous, alone=True)
    run("slabs-2d", slabs, alone=True)
dist.barrier()
dist.destroy_process_group()
