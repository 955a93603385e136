#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "pca or PCA or tc_ or feature or dropin or chain" > gpurun_out/pytest_pca.log 2>&1; tail -3 gpurun_out/pytest_pca.log
for n in 4000000 10000000; do
SS_C5_SPIKES=$n timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5', d['config']['n_spikes'], {k: round(v,3) for k,v in d['stage_ms'].items()}, 'gram frac %.3f project frac %.3f' % (d['roofline']['frac'], d['roofline']['project']['frac']))"
done
export SS_C5_SPIKES=4000000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pca_gram_tc|pca_project" -s 6 -c 2 -o gpurun_out/ncu_c5 -f python bench.py --workload c5 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c5.log 2>&1
tail -1 gpurun_out/ncu_c5.log | cut -c1-200
