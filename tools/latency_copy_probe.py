"""Single-frame latency of compute_depth / extract_centerline / reconstruct_from_depth at config 3 against the number of
copy threads of the staged host <-> device path (L3D_COPY_THREADS; 0 = the driver's pageable copies), plus the raw copy
times of one view up and one depth map down."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laser_3d_reconstruction_b200 as l3d
from laser_3d_reconstruction_b200 import _native as N, pipeline, synth

W, H, D, BS, mode = 1280, 720, 128, 9, 1
K, Q = synth.camera_model(W, H)
maps = synth.warp_maps(W, H, 0) + synth.warp_maps(W, H, 1)
l, r = synth.stereo_pair(W, H, D, 3)
for threads in sys.argv[1:] or ["0", "1", "2", "4", "8"]:
    os.environ["L3D_COPY_THREADS"] = threads
    ctx = N.Context(0)
    ctx.set_rectify_maps(0, maps[0], maps[1]); ctx.set_rectify_maps(1, maps[2], maps[3])
    cfg = pipeline.depth_config(D, BS, mode, Q)
    for old in (getattr(N._default, "ctx", None) or {}).values():  # the classes below pick up a fresh default context
        old.close()
    N._default.__dict__.pop("ctx", None)
    ex = l3d.ImprovedStegerExtractor(sigma=3.0)
    rec = l3d.Reconstructor(K, np.array([0.3, 0, -1, 0.4]), False)
    t = {"compute_depth": [], "extract_centerline": [], "reconstruct_from_depth": [], "bgr2gray": []}
    for i in range(14):
        t0 = time.perf_counter(); rect, depth = ctx.compute_depth(cfg, l, r)
        t1 = time.perf_counter(); pts = ex.extract_centerline(rect)
        t2 = time.perf_counter(); p3 = rec.reconstruct_from_depth(pts, depth)
        t3 = time.perf_counter(); g = ctx.bgr2gray(l)
        t4 = time.perf_counter()
        if i >= 4:
            for k, v in zip(t, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                t[k].append(v)
    print("copy threads %s: " % threads + "  ".join("%s %.2f ms" % (k, 1e3 * float(np.median(v))) for k, v in t.items()),
          "points", len(pts), flush=True)
    ctx.close()

# where the time of extract_centerline / reconstruct_from_depth goes (default copy threads): library call alone against the
# reference-facing method (argument checks + list <-> array conversions on top)
os.environ.pop("L3D_COPY_THREADS", None)
for old in (getattr(N._default, "ctx", None) or {}).values():
    old.close()
N._default.__dict__.pop("ctx", None)
from laser_3d_reconstruction_b200.core.laser_extractor import _steger_params
ctx = N.default_context(0)
ctx.set_rectify_maps(0, maps[0], maps[1]); ctx.set_rectify_maps(1, maps[2], maps[3])
cfg = pipeline.depth_config(D, BS, mode, Q)
rect, depth = ctx.compute_depth(cfg, l, r)
ex = l3d.ImprovedStegerExtractor(sigma=3.0, verbose=False)
rec = l3d.Reconstructor(K, np.array([0.3, 0, -1, 0.4]), False)
sp = _steger_params(N.STEGER_IMPROVED, 3.0, 200, 0.5)
rp = rec._params(N.RECON_DEPTH)
def best(f, n=20):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); out = f(); ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts)), out
t_lib, xy = best(lambda: ctx.steger_extract(sp, rect))
t_cls, pts = best(lambda: ex.extract_centerline(rect))
t_l2a, arr = best(lambda: N.points_to_array(pts))
t_rlib, _ = best(lambda: ctx.reconstruct(rp, arr, depth))
t_rcls, _ = best(lambda: rec.reconstruct_from_depth(pts, depth))
t_gray, _ = best(lambda: ctx.bgr2gray(rect))
print("extract: library call %.2f ms, class method %.2f ms | reconstruct: list->array %.2f ms, library call %.2f ms, class method "
      "%.2f ms | bgr2gray (2.8 MB up, 0.9 MB down) %.2f ms | %d points" % (t_lib, t_cls, t_l2a, t_rlib, t_rcls, t_gray, len(pts)), flush=True)
