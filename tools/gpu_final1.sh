#!/bin/bash
# round-2 evidence, one GPU: tests, bench lines of every workload, reference arm, launch lists, ncu captures
O=gpurun_out/final; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
SS_FILT_SPAN=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1_span.json 2> $O/bench_n1_span.err
for w in c5 c1 c4 c2-bp8 c3; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err || tail -5 $O/bench_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print("%-28s value %.4g %s ms/step %.3f frac %s frac_moved %s" % (f.split('/')[-1], d["value"], d.get("unit", ""), d["ms_per_step"], r.get("frac"), r.get("frac_moved")))
    except Exception as e:
        print(f, "parse failed", e)
PY
# launch lists
W=c2 SKIP=75 COUNT=50 bash tools/gpu_launchlist.sh; cp gpurun_out/launches_c2.csv $O/launches_c2.csv
SS_FILT_SPAN=1 W=c2 SKIP=66 COUNT=44 bash tools/gpu_launchlist.sh; cp gpurun_out/launches_c2.csv $O/launches_c2_span.csv
# full captures of the top kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"filt_apply|filt_summary" -s 9 -c 3 -o $O/ncu_filter_scan -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu1.log 2>&1
SS_FILT_SPAN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:filt_span -s 7 -c 1 -o $O/ncu_filter_span -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu2.log 2>&1
SS_C5_SPIKES=4000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pca_gram_tc|pca_project" -s 6 -c 2 -o $O/ncu_c5 -f python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu3.log 2>&1
ls -la $O | tail -30
