#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_apply_tma.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_tma.log 2>&1
tail -5 gpurun_out/pytest_tma.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "filtfilt or chain or fused or sharded or halo or dropin" > gpurun_out/pytest_quick.log 2>&1
tail -3 gpurun_out/pytest_quick.log
for w in c1 c2; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w', round(d['ms_per_step'],3), d.get('stage_ms'), d.get('ms_filter_stage'))"
done
