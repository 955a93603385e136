#!/bin/bash
# bench at N GPUs (no probes)
N=${N:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err || tail -20 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/bench_n%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("N=%s value %.4g  ms/step %.3f  filter ms %.3f share %.3f  e2e %.4g (%.1f ms)  launches %d" % (
        sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["share_of_step"],
        d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
    print("   parity", d.get("parity"))
except Exception as e:
    print("bench parse failed", e)
PY
