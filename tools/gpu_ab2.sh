#!/bin/bash
# A/B of span-kernel builds on one box (c2, device-only), plus the scan path, plus parity of the default build
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "filtfilt or filter or chain or fused or sharded or halo" > gpurun_out/pytest_filt.log 2>&1; tail -2 gpurun_out/pytest_filt.log
shopt -s nullglob
run() {
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms/step %.3f filter %.3f pca %.3f' % (d['ms_per_step'], d['ms_filter_stage'], d['ms_pca']), d['spikes_kept_rank0'])"
}
for lib in default spikesort_b200/build/lib_v*.so; do
  [ "$lib" = default ] && unset SPIKESORT_B200_LIB || export SPIKESORT_B200_LIB=$PWD/$lib
  run $lib
done
unset SPIKESORT_B200_LIB
SS_FILT_SPAN=0 run scan-path
timeout 900 ncu --set full --clock-control none --import-source on -k regex:filt_span -s 4 -c 1 -o gpurun_out/ncu_span -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
