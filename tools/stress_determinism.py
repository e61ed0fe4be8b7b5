"""Run the same frames through the grouped frame pipeline many times and compare every output bit for bit with the first run:
a race in the cluster kernels' hand-offs (mbarrier phases, relaxed remote arrives, st.async slots) would show as a rare
difference.  Usage: python tools/stress_determinism.py [runs] ;  L3D_PROBE_CFG=c3|c4|c1, L3D_PROBE_GEOM=W,H,D,BS"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laser_3d_reconstruction_b200 import _native as N, pipeline, synth
W, H, D, BS = {'c1': (320, 360, 64, 5), 'c3': (1280, 720, 128, 9), 'c4': (1920, 1080, 256, 11)}[os.environ.get('L3D_PROBE_CFG', 'c3')]
if os.environ.get('L3D_PROBE_GEOM'):
    W, H, D, BS = (int(v) for v in os.environ['L3D_PROBE_GEOM'].split(','))
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
nfr = 14
K, Q = synth.camera_model(W, H)
maps = synth.warp_maps(W, H, 0) + synth.warp_maps(W, H, 1)
base = [synth.stereo_pair(W, H, D, 100 + s) for s in range(nfr)]
L = np.stack([b[0] for b in base]); R = np.stack([b[1] for b in base])
ctx = N.Context(0)
cfg = pipeline.make_pipeline_config(W, H, D, BS, 1, Q, K, lanes=14, max_points=20000, extractor=N.STEGER_IMPROVED)
fp = pipeline.FramePipeline(cfg, maps=maps, ctx=ctx)
dL, dR = fp.upload(L), fp.upload(R)
first = None
bad = 0
for r in range(runs):
    fp.run_dev(dL, dR, nfr)
    got = [fp.fetch(i) for i in range(nfr)]
    sig = [(g["disp16"].tobytes(), g["depth"].tobytes(), g["points_3d"].tobytes()) for g in got]
    if first is None:
        first = sig
    else:
        for i in range(nfr):
            if sig[i] != first[i]:
                bad += 1
                print("run %d frame %d differs from run 0" % (r, i), flush=True)
print("%dx%d D=%d: %d runs x %d frames, %d differences" % (W, H, D, runs, nfr, bad), flush=True)
fp.close(); ctx.close()
sys.exit(1 if bad else 0)
