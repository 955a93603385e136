#!/bin/bash
# launch lists (kernel durations) of the default library and of every variant in build/lib_v*.so
mkdir -p gpurun_out
shopt -s nullglob
for lib in default spikesort_b200/build/lib_v*.so; do
  [ "$lib" = default ] && unset SPIKESORT_B200_LIB || export SPIKESORT_B200_LIB=$PWD/$lib
  name=$(basename $lib .so)
  W=c2 SKIP=${SKIP:-78} COUNT=${COUNT:-52} bash tools/gpu_launchlist.sh > /dev/null
  cp gpurun_out/launches_c2.csv gpurun_out/launches_$name.csv
  echo "== $name"; python tools/launch_table.py gpurun_out/launches_$name.csv | grep -i "filt_\|kernel time" | cut -c1-112
done
