"""Bring-up of the cluster-fused aggregation kernel (L3D_VGROUP=1) against cv2 / the oracle volumes."""
import os, sys, time
os.environ["L3D_VGROUP"] = "1"
import cv2, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from laser_3d_reconstruction_b200 import _native as N, synth
from oracle import cref

def gray_pair(W, H, D, seed):
    l, r = synth.stereo_pair(W, H, D, seed)
    return cv2.cvtColor(l, cv2.COLOR_BGR2GRAY), cv2.cvtColor(r, cv2.COLOR_BGR2GRAY)

ctx = N.Context(0)
cases = [(200, 24, 64, 5, 1), (200, 64, 64, 9, 1), (300, 56, 128, 7, 1), (300, 56, 128, 7, 0), (400, 48, 256, 11, 1), (320, 360, 64, 5, 1), (1280, 720, 128, 9, 1)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
bad = 0
for (W, H, D, bs, mode) in cases:
    lg, rg = gray_pair(W, H, D, 3)
    kw = dict(minDisparity=0, numDisparities=D, blockSize=bs, P1=24 * bs * bs, P2=96 * bs * bs, disp12MaxDiff=1000000,
              preFilterCap=63, uniquenessRatio=0, speckleWindowSize=0, speckleRange=32, mode=mode)
    p = N.SgbmParams(**kw)
    t0 = time.time()
    if W * H * D < 6e6:
        disp, raw, Cg, Sg = ctx.sgbm_compute(p, lg, rg, want_raw=True, want_volumes=True)
        _, oraw, oC, oS = cref.sgbm_compute(lg, rg, want_volumes=True, want_raw=True, **kw)
        nS = int((Sg != oS).sum())
        if nS:
            idx = np.argwhere(Sg != oS)
            print("   S mismatches", nS, "of", Sg.size, "first", idx[:3].tolist(), "rows", np.unique(idx[:, 0])[:8], "cols", np.unique(idx[:, 1])[:8])
    else:
        disp = ctx.sgbm_compute(p, lg, rg); nS = -1
    want = cv2.StereoSGBM_create(**kw).compute(lg, rg)
    nd = int((disp != want).sum())
    bad += nd > 0
    print("case", (W, H, D, bs, mode), "S diff", nS, "disp diff", nd, "%.2fs" % (time.time() - t0), flush=True)
print("BAD" if bad else "ALL OK")
