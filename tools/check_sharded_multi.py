"""Multi-GPU check (torchrun, one process per GPU): run_chain_sharded over real NVLink peer
memory (and over NCCL) equals the unsharded chain computed locally on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_sharded_multi.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spikesort_b200 import dist as ssd, filters, pipeline, synth      # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)

FS, n_c, chunk = 30000, 8, 100000
n = 600000                                   # per shard
x = synth.make_recording(n_c, world * n, FS, seed=77, rate_hz=40.0, noise=8.0, amp=(70.0, 160.0),
                         burst=True, negative=True, device=dev)
tpl = torch.tensor([0, -40, -120, -200, -120, -40, 20, 30, 10, 0], dtype=torch.int16, device=dev)
for r in range(1, world):                    # spikes on every shard boundary
    for c, d in enumerate((-7, -3, -1, 0, 2)):
        x[c, r * n + d:r * n + d + 10] += tpl
    x[1, r * n + 1:r * n + 11] += tpl
flt = filters.Filter(800., 100.)
sp_win = [-0.2, 0.8]
whole = pipeline.run_chain(x, FS, filt=flt, sp_win=sp_win, edge='falling', align_type='min',
                           thresh='auto', ncomps=2, chunksize=chunk)
ok = True
mine_x = x[:, rank * n:(rank + 1) * n].contiguous()
# (transport, halo, lazy): the plain runs (the second 'peer' one capacity-sized), a halo too
# narrow for the spikes on the shard boundaries - every rank must learn of the miss and all of
# them redo the step with a wider halo - and the pipelined (lazy) form of the peer transport
cases = [("peer", None, False), ("nccl", None, False), ("peer", None, False), ("peer", 2, False),
         ("peer", 2, False), ("nccl", 2, False), ("peer", None, True), ("peer", None, "streams")]
side = [torch.cuda.Stream(), torch.cuda.Stream()]
for transport, halo, lazy in cases:
    if lazy == "streams":
        # two steps in flight on two streams, each with its own arena set (slot)
        for st in side:
            st.wait_stream(torch.cuda.current_stream())
        pend = []
        for rep in range(4):
            with torch.cuda.stream(side[rep % 2]):
                pend.append(ssd.run_chain_sharded(mine_x, FS, flt, sp_win, edge='falling',
                                                  align_type='min', thresh='auto', ncomps=2,
                                                  chunksize=chunk, transport=transport, lazy=True,
                                                  slot=rep % 2))
            if rep >= 1:
                with torch.cuda.stream(side[(rep - 1) % 2]):
                    res = pend[rep - 1].finalize()
        with torch.cuda.stream(side[1]):
            res = pend[3].finalize()
        torch.cuda.synchronize()
        lazy = 2
    else:
        res = ssd.run_chain_sharded(mine_x, FS, flt, sp_win, edge='falling', align_type='min',
                                    thresh='auto', ncomps=2, chunksize=chunk, transport=transport,
                                    halo=halo, lazy=lazy)
    if lazy is True:
        nxt = ssd.run_chain_sharded(mine_x, FS, flt, sp_win, edge='falling', align_type='min',
                                    thresh='auto', ncomps=2, chunksize=chunk, transport=transport,
                                    halo=halo, lazy=True)      # queued behind the first step
        res = res.finalize()
        nxt.finalize()
    torch.cuda.synchronize()
    lo_t, hi_t = rank * n * 1000.0 / FS, (rank + 1) * n * 1000.0 / FS
    good = torch.equal(res.thresh, whole.thresh)
    good &= torch.equal(res.filtered, whole.filtered[:, rank * n:(rank + 1) * n])
    n_mine = 0
    for c in range(n_c):
        w, p = whole.per_contact(c), res.per_contact(c)
        # my part of the whole list: positions given by the layout every rank holds
        before = int(res.all_counts[:rank, c].sum())
        mine = int(res.all_counts[rank, c])
        good &= int(res.all_counts[:, c].sum()) == len(w['spt'])
        good &= np.array_equal(p['spt'], w['spt'][before:before + mine])
        good &= np.array_equal(p['waves'], w['waves'][:, before:before + mine])
        sc, ws = p['scores'], w['scores'][before:before + mine]
        sgn = np.sign(np.sum(sc * ws, axis=0))
        good &= bool(np.abs(sc * sgn - ws).max() <= 1e-9 * np.abs(w['scores']).max()) if mine else True
        det_w = w['spt_detected']
        sel = det_w[(det_w >= lo_t) & (det_w < hi_t)]
        good &= np.array_equal(p['spt_detected'], sel)
        n_mine += mine
    flag = torch.tensor([int(good)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("transport %-4s halo %-4s lazy %d world %d: %s (%d spikes on rank 0)" %
              (transport, halo, lazy, world, "OK" if int(flag) else "MISMATCH", n_mine), flush=True)
    ok &= bool(int(flag))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
