"""CPU tier: the multi-GPU exchange steps of spikesort_b200/dist.py on a world of 2
gloo processes (CPU tensors; the same helpers run over NCCL on the B200 box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from spikesort_b200 import dist as ssd
        out = {}
        assert ssd.world() == (rank, world)
        # thresholds: only rank 0 knows them
        thr = torch.tensor([-10.0, -20.0, -30.0], dtype=torch.float64) if rank == 0 \
            else torch.zeros(3, dtype=torch.float64)
        thr = ssd.share_thresholds(thr)
        out["thr"] = thr.tolist()
        # neighbour halo: first sample of the right neighbour, NaN on the last rank
        first = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64) * (rank + 1) - 25.0
        nxt = ssd.exchange_first_samples(first)
        out["nxt"] = [None if np.isnan(v) else v for v in nxt.tolist()]
        last = torch.tensor([-5.0, -5.0, -50.0], dtype=torch.float64)
        mask = ssd.boundary_crossings(last, nxt, thr, True)
        out["mask"] = mask.tolist()
        spt = torch.tensor([1.0, 2.0, 5.0, 6.0, 7.0], dtype=torch.float64)
        off = torch.tensor([0, 2, 2, 5], dtype=torch.int64)
        spt2, off2 = ssd.append_boundary_spikes(spt, off, mask, 99.0)
        out["spt2"], out["off2"] = spt2.tolist(), off2.tolist()
        # Gram merge
        gram = torch.full((3, 2, 2), float(rank + 1), dtype=torch.float64)
        ssum = torch.full((3, 2), 10.0 * (rank + 1), dtype=torch.float64)
        counts = torch.tensor([rank, 4, 0], dtype=torch.int64)
        g, s, c, lay = ssd.merge_gram(gram, ssum, counts, with_layout=True)
        out["gram"], out["ssum"], out["counts"] = g[0, 0, 0].item(), s[1, 1].item(), c.tolist()
        out["layout"] = lay.tolist()
        # shared reference: lowest rank that has a spike on the contact
        ref = torch.full((3, 2), float(rank + 1), dtype=torch.float64)
        has = torch.tensor([rank == 1, True, False])
        out["ref"] = ssd.share_reference(ref, has)[:, 0].tolist()
        out["all_counts"] = ssd.gather_counts(torch.tensor([0, rank + 1, 3 + rank, 3 + rank])).tolist()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_exchange_steps_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert res[r]["thr"] == [-10.0, -20.0, -30.0]
        assert res[r]["gram"] == 3.0 and res[r]["ssum"] == 30.0 and res[r]["counts"] == [1, 8, 0]
        assert res[r]["ref"] == [2.0, 1.0, 1.0]
        assert res[r]["layout"] == [[0, 4, 0], [1, 4, 0]]
        assert res[r]["all_counts"] == [[1, 2, 0], [2, 2, 0]]
    # rank 0 sees rank 1's first samples (2,4,6)-25; falling crossing needs last > thr > next
    assert res[0]["nxt"] == [-23.0, -21.0, -19.0] and res[1]["nxt"] == [None, None, None]
    assert res[0]["mask"] == [True, True, False] and res[1]["mask"] == [False, False, False]
    assert res[0]["spt2"] == [1.0, 2.0, 99.0, 99.0, 5.0, 6.0, 7.0] and res[0]["off2"] == [0, 3, 4, 7]
    assert res[1]["spt2"] == [1.0, 2.0, 5.0, 6.0, 7.0] and res[1]["off2"] == [0, 2, 2, 5]


def test_single_process_is_identity():
    from spikesort_b200 import dist as ssd
    thr = torch.tensor([1.0, 2.0], dtype=torch.float64)
    assert ssd.share_thresholds(thr) is thr
    assert torch.isnan(ssd.exchange_first_samples(thr)).all()
    g, s, c = ssd.merge_gram(torch.ones(2, 3, 3), torch.ones(2, 3), torch.tensor([4, 5]))
    assert c.tolist() == [4, 5]
    assert ssd.gather_counts(torch.tensor([0, 2, 5])).tolist() == [[2, 3]]
