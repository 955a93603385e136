#!/bin/bash
mkdir -p gpurun_out
shopt -s nullglob
run() {
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms/step %.3f filter %.3f pca %.3f' % (d['ms_per_step'], d['ms_filter_stage'], d['ms_pca']), d['spikes_kept_rank0'])"
}
for lib in default spikesort_b200/build/lib_v*.so; do
  [ "$lib" = default ] && unset SPIKESORT_B200_LIB || export SPIKESORT_B200_LIB=$PWD/$lib
  run span-$lib
done
unset SPIKESORT_B200_LIB
SS_FILT_SPAN=0 run scan-default
SPIKESORT_B200_LIB=$PWD/spikesort_b200/build/lib_vscalario.so timeout 300 python -m pytest tests -m gpu -q --timeout 300 -x -k "filtfilt or fused or sharded" 2>&1 | tail -2
