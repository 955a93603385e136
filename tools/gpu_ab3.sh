#!/bin/bash
shopt -s nullglob
for lib in default spikesort_b200/build/lib_v*.so; do
  [ "$lib" = default ] && unset SPIKESORT_B200_LIB || export SPIKESORT_B200_LIB=$PWD/$lib
  SS_C5_SPIKES=10000000 timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', {k: round(v,3) for k,v in d['stage_ms'].items()}, 'gram frac %.3f' % d['roofline']['frac'])"
done
