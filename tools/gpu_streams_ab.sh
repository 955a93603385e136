#!/bin/bash
mkdir -p gpurun_out
for n in 2 3 4 2 3 4; do
  SS_BENCH_STREAMS=$n timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('streams $n', 'ms/step %.3f filter %.3f pca %.3f' % (d['ms_per_step'], d['ms_filter_stage'], d['ms_pca']), d['spikes_kept_rank0'])"
done
