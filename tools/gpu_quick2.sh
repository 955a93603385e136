#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_q.log 2>&1; tail -3 gpurun_out/pytest_q.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err || tail -5 gpurun_out/bench_q.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
    print("value %.4g  ms/step %.3f  filter ms %.3f  frac %.3f  chain frac %.3f  e2e %.4g (%.1f ms) per-contact %.1f ms launches %d" % (
        d["value"], d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"],
        d["chain_roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["per_contact_api"]["ms_per_step"], d["gpu_launches"]))
    print(d["parity"], d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
SKIP=72 COUNT=48 bash tools/gpu_launchlist.sh
