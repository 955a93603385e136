#!/bin/bash
# final numbers of round 2 (one GPU): bench lines, launch list, ncu captures of the two new kernels
O=gpurun_out/final_r2; mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err
for w in c5 c1 c4 c3 c2-bp8; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err
done
W=c2 SKIP=78 COUNT=52 bash tools/gpu_launchlist.sh; cp gpurun_out/launches_c2.csv $O/launches_c2.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"filt_apply_tma" -s 1 -c 1 -o $O/ncu_apply_tma -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/ncu_apply.log 2>&1
SS_C5_SPIKES=4000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pca_project_vec4" -s 3 -c 1 -o $O/ncu_project_vec4 -f python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_project.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final_r2/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print("%-28s value %.4g %s ms/step %.3f frac %s frac_moved %s launches %s" % (f.split('/')[-1], d["value"], d.get("unit", ""), d["ms_per_step"], r.get("frac"), r.get("frac_moved"), d.get("gpu_launches")))
    except Exception as e:
        print(f, "parse failed", e)
PY
