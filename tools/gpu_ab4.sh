#!/bin/bash
# two-sweep path (SS_FILT_SPAN=0) with the vectorised sweep-2 load/store: parity, then A/B of register caps
mkdir -p gpurun_out
export SS_FILT_SPAN=0
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "filtfilt or filter or chain or fused or sharded or halo or dropin" > gpurun_out/pytest_scan.log 2>&1; tail -2 gpurun_out/pytest_scan.log
shopt -s nullglob
run() {
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms/step %.3f filter %.3f pca %.3f' % (d['ms_per_step'], d['ms_filter_stage'], d['ms_pca']), d['spikes_kept_rank0'])"
}
for lib in default spikesort_b200/build/lib_v*.so; do
  [ "$lib" = default ] && unset SPIKESORT_B200_LIB || export SPIKESORT_B200_LIB=$PWD/$lib
  run scan-$lib
done
unset SPIKESORT_B200_LIB
SS_FILT_SPAN=1 run span-default
