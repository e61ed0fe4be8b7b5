#!/bin/bash
mkdir -p gpurun_out
L3D_PROBE_CFG=c4 timeout 300 python tools/kernel_times.py 2 2>&1 | tail -2
L3D_PROBE_CFG=c4 timeout 300 python tools/skip_probe.py 21 42 2>&1 | tail -1
for s in sgbm_vgroup sgbm_scan_hpair sgbm_cost fgs_ wls_,fgs_; do
  L3D_DEBUG_SKIP=$s L3D_PROBE_CFG=c4 timeout 300 python tools/skip_probe.py 21 42 2>&1 | tail -1
done
