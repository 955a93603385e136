"""GPU tier, needs >= 2 GPUs (skipped otherwise): the time-sharded chain over REAL NVLink peer
memory and over NCCL, one process per GPU under torchrun, equals the unsharded chain - spike
times and waveforms bit for bit, scores to round-off - including the halo-widening rerun and
the pipelined (lazy) form.  The work is done by tools/check_sharded_multi.py (exit code 0 = every
rank agreed for every case); this wrapper makes it part of `pytest -m gpu`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_chain_over_nvlink_equals_whole(cuda):
    import torch
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("one GPU: the multi-process check needs at least two")
    world = 2 if n_gpus < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tools", "check_sharded_multi.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(res.stdout[-4000:])
    print(res.stderr[-4000:])
    assert res.returncode == 0, "sharded != whole on %d GPUs" % world
    assert res.stdout.count(": OK") >= 8 and "MISMATCH" not in res.stdout
