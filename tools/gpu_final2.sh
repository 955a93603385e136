#!/bin/bash
O=gpurun_out/final2; mkdir -p $O
timeout 900 python bench.py --workload c3 --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 900 python bench.py --workload c2-bp8 --steps 10 --warmup 3 > $O/bench_c2-bp8.json 2> $O/bench_bp8.err
W=c2 SKIP=78 COUNT=52 bash tools/gpu_launchlist.sh; cp gpurun_out/launches_c2.csv $O/launches_c2.csv
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final2/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print("%-28s value %.4g %s ms/step %.3f frac %s frac_moved %s launches %s" % (f.split('/')[-1], d["value"], d.get("unit", ""), d["ms_per_step"], r.get("frac"), r.get("frac_moved"), d.get("gpu_launches")))
    except Exception as e:
        print(f, "parse failed", e)
PY
