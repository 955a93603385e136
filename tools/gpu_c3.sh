#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "filtfilt or filter or chain or fused or sharded or halo or span or dropin" > gpurun_out/pytest_f.log 2>&1; tail -2 gpurun_out/pytest_f.log
for w in c2-bp8 c2; do
timeout 900 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d.get('roofline') or {}
print('$w', 'ms/step %.3f' % d['ms_per_step'], 'filter', d.get('ms_filter_stage') or r.get('ms_per_launch'), d.get('clocks'))"
done
