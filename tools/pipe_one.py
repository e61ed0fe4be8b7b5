import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laser_3d_reconstruction_b200 import _native as N, pipeline, synth
W, H, D, BS = {'c1': (320, 360, 64, 5), 'c3': (1280, 720, 128, 9), 'c4': (1920, 1080, 256, 11)}[os.environ.get('L3D_PROBE_CFG', 'c3')]
K, Q = synth.camera_model(W, H)
maps = synth.warp_maps(W, H, 0) + synth.warp_maps(W, H, 1)
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 32
base = [synth.stereo_pair(W, H, D, s) for s in range(4)]
L = np.stack([base[i % 4][0] for i in range(nfr)]); R = np.stack([base[i % 4][1] for i in range(nfr)])
ctx = N.Context(0)
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = pipeline.make_pipeline_config(W, H, D, BS, 1, Q, K, lanes=lanes, max_points=20000, extractor=N.STEGER_IMPROVED)
fp = pipeline.FramePipeline(cfg, maps=maps, ctx=ctx)
dL, dR = fp.upload(L), fp.upload(R)
for _ in range(3):
    fp.run_dev(dL, dR, nfr)
    print("run: %.3f ms/frame" % (fp.last_ms / nfr), flush=True)
