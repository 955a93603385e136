#!/bin/bash
# whole GPU tier + smoke + the secondary workloads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_full.log 2>&1; tail -5 gpurun_out/pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in c3 c2-bp8 c1 c4; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || tail -5 gpurun_out/bench_$w.err
  python - $w <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("%s: value %.4g ms/step %.3f frac %s frac_moved %s parity %s" % (sys.argv[1], d["value"], d["ms_per_step"], r.get("frac"), r.get("frac_moved"), d.get("parity")))
except Exception as e:
    print("bench parse failed", e)
PY
done
