#!/bin/bash
# config 4: lane-set sizes against the number of resident 16-CTA clusters
mkdir -p gpurun_out
for gl in "4 12" "4 16" "4 20" "8 16" "8 24" "7 21" "7 28"; do
  set -- $gl
  echo "group $1 lanes $2: $(L3D_GROUP=$1 L3D_PROBE_CFG=c4 timeout 120 python tools/skip_probe.py $2 $((2*$2)) 2>&1 | tail -1)"
done
