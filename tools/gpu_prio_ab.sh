#!/bin/bash
mkdir -p gpurun_out
for cfg in "2 0" "2 1" "1 1" "3 1" "2 0" "2 1"; do
  set -- $cfg
  SS_BENCH_STREAMS=$1 SS_CHAIN_TAIL_PRIO=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('streams $1 tail-prio $2', 'ms/step %.3f filter %.3f pca %.3f' % (d['ms_per_step'], d['ms_filter_stage'], d['ms_pca']), d['spikes_kept_rank0'])" || tail -3 gpurun_out/bench_ab.err
done
SS_CHAIN_TAIL_PRIO=1 timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "chain or sharded or frontend or dropin" 2>&1 | tail -2
