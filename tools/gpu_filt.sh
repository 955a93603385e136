#!/bin/bash
# filter regression on the GPU box: filter parity tests, then the bench line of both configs[1] filters
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "filtfilt or filter or chain or fused or smoke" > gpurun_out/pytest_filt.log 2>&1
tail -5 gpurun_out/pytest_filt.log
for w in c2 c2-bp8; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err || tail -5 gpurun_out/bench_$w.err
  python - $w <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/bench_%s.log" % sys.argv[1]).read().strip().splitlines()[-1])
    print("%s: value %.4g  ms/step %.3f  filter ms %.3f  filter frac %.3f  chain frac %.3f  pca ms %.3f" % (
        sys.argv[1], d["value"], d["ms_per_step"], d.get("ms_filter_stage", 0), d.get("filter_stage_frac_of_hbm_peak", 0),
        0, d.get("ms_pca", 0)), d["clocks"]["reasons"])
except Exception as e:
    print("bench parse failed", e)
PY
done
if [ -n "$NCU_BP8" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bp8.csv python bench.py --workload c2-bp8 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bp8.log 2>&1
  python tools/ncu_summary.py launches gpurun_out/launches_bp8.csv | grep -i "filt_\|pca_eig"
fi
