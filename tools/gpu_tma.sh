#!/bin/bash
# sweep 2 through the copy engine: its tests, the filter tests, then an A/B of the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_apply_tma.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_tma.log 2>&1
tail -5 gpurun_out/pytest_tma.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "filtfilt or chain or fused or sharded or halo" > gpurun_out/pytest_quick.log 2>&1
tail -3 gpurun_out/pytest_quick.log
run() {
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_tma.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms/step %.3f filter %.3f pca %.3f' % (d['ms_per_step'], d['ms_filter_stage'], d['ms_pca']), d['spikes_kept_rank0'])"
}
SS_APPLY_TMA=0 run element
shopt -s nullglob
for lib in default spikesort_b200/build/lib_v*.so; do
  [ "$lib" = default ] && unset SPIKESORT_B200_LIB || export SPIKESORT_B200_LIB=$PWD/$lib
  run $lib
done
unset SPIKESORT_B200_LIB
SS_APPLY_TMA=0 run element
