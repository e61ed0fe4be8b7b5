#!/bin/bash
# lane-set sizes at config 3 with the wavefront kernel (15 cluster slots: two half-size launches can be resident at once,
# so that one's fill / drain overlaps the other's steady state) and config 2 with / without the wavefront kernel
for gl in "7 28" "4 28" "4 32" "3 24" "2 16"; do
  set -- $gl
  echo "group $1 lanes $2: $(L3D_GROUP=$1 timeout 120 python tools/skip_probe.py $2 $((4*$2)) 2>&1 | tail -1)"
done
echo "c2-like (320x360 D=64 MODE_HH) wavefront: $(L3D_PROBE_CFG=c1 timeout 120 python tools/skip_probe.py 28 224 2>&1 | tail -1)"
echo "c2-like (320x360 D=64 MODE_HH) no wavefront: $(L3D_NO_VWAVE=1 L3D_PROBE_CFG=c1 timeout 120 python tools/skip_probe.py 28 224 2>&1 | tail -1)"
