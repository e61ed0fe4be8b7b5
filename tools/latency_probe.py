"""Single-frame latency of the reference-facing calls (host buffers in, host buffers out, one frame at a time) at config 3
and config 1: what a user of SingleUSBStereoCameraManager.compute_depth / extract_centerline / reconstruct_from_depth sees."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laser_3d_reconstruction_b200 as l3d
from laser_3d_reconstruction_b200 import _native as N, pipeline, synth

for (W, H, D, BS, mode, name) in ((1280, 720, 128, 9, 1, "c3 MODE_HH"), (320, 360, 64, 5, 2, "c1 3WAY")):
    K, Q = synth.camera_model(W, H)
    maps = synth.warp_maps(W, H, 0) + synth.warp_maps(W, H, 1)
    l, r = synth.stereo_pair(W, H, D, 3)
    ctx = N.Context(0)
    ctx.set_rectify_maps(0, maps[0], maps[1]); ctx.set_rectify_maps(1, maps[2], maps[3])
    cfg = pipeline.depth_config(D, BS, mode, Q)
    ex = l3d.ImprovedStegerExtractor(sigma=3.0)
    rec = l3d.Reconstructor(K, np.array([0.3, 0, -1, 0.4]), False)
    t = {"compute_depth": [], "extract_centerline": [], "reconstruct_from_depth": []}
    for i in range(12):
        t0 = time.perf_counter(); rect, depth = ctx.compute_depth(cfg, l, r)
        t1 = time.perf_counter(); pts = ex.extract_centerline(rect)
        t2 = time.perf_counter(); p3 = rec.reconstruct_from_depth(pts, depth)
        t3 = time.perf_counter()
        if i >= 2:
            t["compute_depth"].append(t1 - t0); t["extract_centerline"].append(t2 - t1); t["reconstruct_from_depth"].append(t3 - t2)
    print(name, "  ".join("%s %.2f ms" % (k, 1e3 * float(np.median(v))) for k, v in t.items()), "points", len(pts), flush=True)
    ctx.close()

# cv2.StereoBM look-alike (readme.md:392-397), host arrays in and out
import cv2
from laser_3d_reconstruction_b200 import stereo
for (W, H, D, bs) in ((320, 360, 64, 15), (1280, 720, 128, 15)):
    l, r = synth.stereo_pair(W, H, D, 3)
    lg, rg = cv2.cvtColor(l, cv2.COLOR_BGR2GRAY), cv2.cvtColor(r, cv2.COLOR_BGR2GRAY)
    m = stereo.StereoBM_create(numDisparities=D, blockSize=bs)
    ref = cv2.StereoBM_create(numDisparities=D, blockSize=bs)
    tg, tc = [], []
    for i in range(8):
        t0 = time.perf_counter(); a = m.compute(lg, rg); t1 = time.perf_counter(); b = ref.compute(lg, rg); t2 = time.perf_counter()
        if i >= 2:
            tg.append(t1 - t0); tc.append(t2 - t1)
    print("StereoBM %dx%d D=%d bs=%d: GPU %.2f ms, cv2 (host, %d threads) %.2f ms, equal %s" % (
        W, H, D, bs, 1e3 * float(np.median(tg)), cv2.getNumThreads(), 1e3 * float(np.median(tc)), bool(np.array_equal(a, b))), flush=True)
