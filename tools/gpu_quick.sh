#!/bin/bash
# quick regression on the GPU box: filter/detect parity tests + one bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "${1:-filtfilt or chain or detect or dropin}" > gpurun_out/pytest_quick.log 2>&1
tail -3 gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_quick.log").read().strip().splitlines()[-1])
    print("value %.4g  ms/step %.3f  filter ms %.3f  filter frac %.3f  chain frac %.3f  e2e %.4g (%.1f ms)  launches %d" % (
        d["value"], d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"],
        d["chain_roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
    print(d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
