#!/bin/bash
mkdir -p gpurun_out
K=${K:-filt_apply_tma}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s ${SKIP:-1} -c ${COUNT:-1} -o gpurun_out/ncu_tma -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_tma.log 2>&1
tail -2 gpurun_out/ncu_tma.log | cut -c1-300
