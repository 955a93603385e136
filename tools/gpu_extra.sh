#!/bin/bash
# the extra workloads (bench.py --workload c5 / c1 / c4) on the GPU box
mkdir -p gpurun_out
for w in c5 c1 c4; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || tail -15 gpurun_out/bench_$w.err
  python - $w <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g %s  ms/step %.3f" % (d["value"], d["unit"], d["ms_per_step"]))
    print("   stages", {k: round(v, 3) for k, v in d["stage_ms"].items()})
    print("   roofline frac %.3f" % d["roofline"]["frac"], "e2e", d.get("e2e") and round(d["e2e"]["ms_per_step"], 2))
    print("   parity", d.get("parity"))
    print("   cpu", d.get("cpu_baseline"))
except Exception as e:
    print(sys.argv[1], "parse failed", e)
PY
done
