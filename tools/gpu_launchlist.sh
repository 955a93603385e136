#!/bin/bash
# ncu launch list (durations + DRAM bytes) of the library's own kernels over two steps of bench.py
mkdir -p gpurun_out
W=${W:-c2}
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"filt_|detect_|var_|scan_|mark|emit|align|rd_|extract|pca_|first_wave|merge" -s ${SKIP:-120} -c ${COUNT:-80} --csv \
  --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
