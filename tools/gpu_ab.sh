#!/bin/bash
# A/B of library builds on one box: bench lines of bench.py --workload $W for every lib in build/lib_v*.so
mkdir -p gpurun_out
W=${W:-c2-bp8}
shopt -s nullglob
for lib in default spikesort_b200/build/lib_v*.so; do
  [ "$lib" = default ] && unset SPIKESORT_B200_LIB || export SPIKESORT_B200_LIB=$PWD/$lib
  timeout 300 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'ms/step %.3f filter %.3f pca %.3f' % (d['ms_per_step'], d['ms_filter_stage'], d['ms_pca']), d['spikes_kept_rank0'])"
done
