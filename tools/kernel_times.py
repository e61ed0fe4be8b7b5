"""Per-kernel-group times of one SGBM matcher run at config 3, every kernel timed alone (the roofline leg of bench.py,
repeated to show the noise).  Usage: python tools/kernel_times.py [repeats]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laser_3d_reconstruction_b200 import _native as N, pipeline, synth
W, H, D, BS = {'c1': (320, 360, 64, 5), 'c3': (1280, 720, 128, 9), 'c4': (1920, 1080, 256, 11)}[os.environ.get('L3D_PROBE_CFG', 'c3')]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
K, Q = synth.camera_model(W, H)
maps = synth.warp_maps(W, H, 0) + synth.warp_maps(W, H, 1)
base = [synth.stereo_pair(W, H, D, s) for s in range(4)]
nfr = 14
L = np.stack([base[i % 4][0] for i in range(nfr)]); R = np.stack([base[i % 4][1] for i in range(nfr)])
ctx = N.Context(0)
MODE = int(os.environ.get('L3D_PROBE_MODE', '1'))
cfg = pipeline.make_pipeline_config(W, H, D, BS, MODE, Q, K, lanes=14, max_points=20000, extractor=N.STEGER_IMPROVED)
fp = pipeline.FramePipeline(cfg, maps=maps, ctx=ctx)
dL, dR = fp.upload(L), fp.upload(R)
for _ in range(2): fp.run_dev(dL, dR, nfr)
fp.set_timing(True)
for r in range(reps):
    fp.run_dev(dL, dR, nfr)
    out = []
    for g in ["sgbm_cost", "sgbm_scan_k0", "sgbm_wta", "sgbm_vgroup_down", "sgbm_vgroup_up", "sgbm_vwave_down", "sgbm_vwave_up", "wls"]:
        t, k = fp.kernel_time(g)
        out.append("%s %.4f" % (g, t / (nfr if g == "wls" else 2 * nfr)))
    print("  ".join(out), flush=True)
fp.close()
