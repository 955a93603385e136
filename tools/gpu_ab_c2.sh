#!/bin/bash
# A/B of library builds on one box (configs[1] bench line, kernel times from a launch list of the default)
mkdir -p gpurun_out
run() {
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms/step %.3f filter %.3f pca %.3f' % (d['ms_per_step'], d['ms_filter_stage'], d['ms_pca']), d['spikes_kept_rank0'])"
}
shopt -s nullglob
for rep in 1 2; do
for lib in default spikesort_b200/build/lib_v*.so; do
  [ "$lib" = default ] && unset SPIKESORT_B200_LIB || export SPIKESORT_B200_LIB=$PWD/$lib
  run $lib
done
done
