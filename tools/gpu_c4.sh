#!/bin/bash
# config 4 with the D = 256 wavefront kernel: lane-set sizes / lanes
mkdir -p gpurun_out
for gl in "7 21" "7 28" "4 16" "4 24" "8 24" "9 27" "5 20"; do
  set -- $gl
  echo "group $1 lanes $2: $(L3D_GROUP=$1 L3D_PROBE_CFG=c4 timeout 120 python tools/skip_probe.py $2 $((2*$2)) 2>&1 | tail -1)"
done
