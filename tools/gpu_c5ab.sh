#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "pca or PCA or tc_ or feature or dropin or ten_million" > gpurun_out/pytest_pca.log 2>&1; tail -3 gpurun_out/pytest_pca.log
shopt -s nullglob
for lib in default spikesort_b200/build/lib_v*.so; do
  [ "$lib" = default ] && unset SPIKESORT_B200_LIB || export SPIKESORT_B200_LIB=$PWD/$lib
  timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', d['config']['n_spikes'], {k: round(v,3) for k,v in d['stage_ms'].items()}, 'gram frac %.3f project frac %.3f' % (d['roofline']['frac'], d['roofline']['project']['frac']), (d.get('parity') or {}).get('pca_score_max_rel_err'))"
done
