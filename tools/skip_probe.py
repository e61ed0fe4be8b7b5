"""Steady-state frames/s of the c3 frame pipeline with some kernels skipped (L3D_DEBUG_SKIP, see common.cuh):
the marginal cost of a kernel inside the overlapped pipeline.  Results of a skipped run are garbage by design.
Usage: L3D_DEBUG_SKIP=fgs_lines python tools/skip_probe.py [lanes] [frames]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laser_3d_reconstruction_b200 import _native as N, pipeline, synth
W, H, D, BS = 1280, 720, 128, 9
if os.environ.get('L3D_PROBE_CFG') == 'c4':
    W, H, D, BS = 1920, 1080, 256, 11
if os.environ.get('L3D_PROBE_CFG') == 'c1':
    W, H, D, BS = 320, 360, 64, 5
if os.environ.get('L3D_PROBE_GEOM'):  # "W,H,D,BS"
    W, H, D, BS = (int(v) for v in os.environ['L3D_PROBE_GEOM'].split(','))
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 14
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 56
K, Q = synth.camera_model(W, H)
maps = synth.warp_maps(W, H, 0) + synth.warp_maps(W, H, 1)
base = [synth.stereo_pair(W, H, D, s) for s in range(4)]
L = np.stack([base[i % 4][0] for i in range(nfr)]); R = np.stack([base[i % 4][1] for i in range(nfr)])
ctx = N.Context(0)
MODE = int(os.environ.get('L3D_PROBE_MODE', '1'))  # cv2.StereoSGBM mode: 0 SGBM, 1 HH, 2 3WAY (the reference's default), 3 HH4
cfg = pipeline.make_pipeline_config(W, H, D, BS, MODE, Q, K, lanes=lanes, max_points=20000, extractor=N.STEGER_IMPROVED)
fp = pipeline.FramePipeline(cfg, maps=maps, ctx=ctx)
dL, dR = fp.upload(L), fp.upload(R)
for _ in range(3): fp.run_dev(dL, dR, nfr)
ms = []
for _ in range(4):
    fp.run_dev(dL, dR, nfr); ms.append(fp.last_ms)
print("skip=%-40s lanes %2d: %.3f ms/frame  (%.0f fps)" % (os.environ.get("L3D_DEBUG_SKIP", "-"), lanes, min(ms) / nfr, nfr / min(ms) * 1e3), flush=True)
fp.close()
