"""GPU bring-up probe: every stage against the oracle / cv2 with verbose mismatch diagnostics, then
a first timing of the c3 pipeline.  Run on the GPU box:  python tests/bringup/gpu_probe.py [quick]
(not a pytest; tests/ holds the real parity suite)."""
import ctypes as C
import os
import sys
import time
import traceback

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from laser_3d_reconstruction_b200 import _native as N  # noqa: E402
from laser_3d_reconstruction_b200 import synth  # noqa: E402
from oracle import cref, ref_ops  # noqa: E402

QUICK = "quick" in sys.argv
FAILS = []


def report(name, a, b, tol=None):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        print("  [FAIL] %-40s shape %s vs %s" % (name, a.shape, b.shape)); FAILS.append(name); return False
    if tol is None:
        bad = a != b
    else:
        bad = np.abs(a.astype(np.float64) - b.astype(np.float64)) > tol
    nb = int(bad.sum())
    if nb == 0:
        print("  [ ok ] %-40s %s" % (name, a.shape)); return True
    idx = np.argwhere(bad)[:4]
    vals = [(tuple(int(v) for v in i), a[tuple(i)].item(), b[tuple(i)].item()) for i in idx]
    print("  [FAIL] %-40s %d/%d differ; first (idx, got, want): %s" % (name, nb, a.size, vals))
    FAILS.append(name)
    return False


def section(title):
    print("\n== %s" % title, flush=True)


def sgbm_case(ctx, l, r, kw, name, volumes=True):
    p = N.SgbmParams(**kw)
    want = cv2.StereoSGBM_create(**kw).compute(l, r)
    if volumes and kw["mode"] != 2:
        disp, raw, Cg, Sg = ctx.sgbm_compute(p, l, r, want_raw=True, want_volumes=True)
        od, oraw, oC, oS = cref.sgbm_compute(l, r, want_volumes=True, want_raw=True, **kw)
        report(name + " C", Cg, oC)
        report(name + " S", Sg, oS)
        report(name + " raw", raw, oraw)
    elif volumes:
        disp, raw = ctx.sgbm_compute(p, l, r, want_raw=True)
        od, oraw = cref.sgbm_compute(l, r, want_raw=True, **kw)
        report(name + " raw", raw, oraw)
    else:
        disp = ctx.sgbm_compute(p, l, r)
    return report(name + " disp==cv2", disp, want)


def gray_pair(W, H, D, seed, quant=0):
    l, r = synth.stereo_pair(W, H, D, seed)
    lg, rg = cv2.cvtColor(l, cv2.COLOR_BGR2GRAY), cv2.cvtColor(r, cv2.COLOR_BGR2GRAY)
    if quant:
        lg = (lg // quant * quant).astype(np.uint8); rg = (rg // quant * quant).astype(np.uint8)
    return lg, rg


def guarded(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()
        FAILS.append(fn.__name__)


def main():
    ctx = N.Context(0)
    print("lib:", ctx.lib.l3d_version().decode())
    rng = np.random.default_rng(0)

    def t_remap():
        section("remap + gray")
        for t in range(3):
            H, W = (53, 97) if t == 0 else (240, 320)
            src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
            mx = (rng.random((H, W)) * (W + 8) - 4).astype(np.float32)
            my = (rng.random((H, W)) * (H + 8) - 4).astype(np.float32)
            if t == 1:
                mx = (np.round(mx * 64) / 64).astype(np.float32); my = (np.round(my * 64) / 64).astype(np.float32)
            ctx.set_rectify_maps(0, mx, my)
            rect, gray = ctx.remap_gray(0, src, (H, W))
            want = cv2.remap(src, mx, my, cv2.INTER_LINEAR)
            report("remap %d" % t, rect, want)
            report("gray %d" % t, gray, cv2.cvtColor(want, cv2.COLOR_BGR2GRAY))
        big = rng.integers(0, 256, (60, 200, 3), dtype=np.uint8)
        view = big[:, 100:]  # non-contiguous split view like _split_frame
        mx, my = synth.warp_maps(100, 60)
        ctx.set_rectify_maps(1, mx, my)
        rect, gray = ctx.remap_gray(1, view, (60, 100))
        report("remap view", rect, cv2.remap(view, mx, my, cv2.INTER_LINEAR))
    guarded(t_remap)

    def t_sgbm_small():
        section("SGBM small, all modes, volumes vs oracle")
        for mode in (0, 1, 2):
            for (H, W, D, bs) in [(48, 96, 16, 3), (50, 130, 32, 5), (64, 200, 64, 9), (56, 300, 128, 7), (48, 400, 256, 11),
                                  (52, 180, 96, 5), (50, 120, 48, 3)]:
                for minD, swap in ((0, False), (-(D - 1), True), (3, False)):
                    for (uq, d12, sw) in [(10, 1, 100), (0, 1000000, 0)]:
                        lg, rg = gray_pair(W, H, D, 3 + mode, quant=16 if bs == 5 else 0)
                        if swap:
                            lg, rg = rg, lg
                        kw = dict(minDisparity=minD, numDisparities=D, blockSize=bs, P1=24 * bs * bs, P2=96 * bs * bs,
                                  disp12MaxDiff=d12, preFilterCap=63, uniquenessRatio=uq, speckleWindowSize=sw,
                                  speckleRange=32, mode=mode)
                        ok = sgbm_case(ctx, lg, rg, kw, "m%d %dx%d D%d bs%d minD%d u%d" % (mode, W, H, D, bs, minD, uq))
                        if not ok and QUICK:
                            return
    guarded(t_sgbm_small)

    def t_sgbm_big():
        section("SGBM c1 / c3 sizes vs cv2")
        for mode in (2, 0, 1):
            lg, rg = gray_pair(320, 360, 64, 7)
            base, mut, right = ref_ops.sgbm_param_sets(64, 5, mode)
            sgbm_case(ctx, lg, rg, base, "c1 m%d as-constructed" % mode, volumes=False)
            sgbm_case(ctx, lg, rg, mut, "c1 m%d wls-mutated" % mode, volumes=False)
            sgbm_case(ctx, rg, lg, right, "c1 m%d right-matcher" % mode, volumes=False)
        lg, rg = gray_pair(1280, 720, 128, 11)
        for mode in (1, 0, 2):
            base, mut, right = ref_ops.sgbm_param_sets(128, 9, mode)
            sgbm_case(ctx, lg, rg, mut, "c3 m%d wls-mutated" % mode, volumes=False)
            if mode == 1:
                sgbm_case(ctx, rg, lg, right, "c3 m1 right-matcher", volumes=False)
                sgbm_case(ctx, lg, rg, base, "c3 m1 as-constructed", volumes=False)
        if not QUICK:
            lg, rg = gray_pair(1920, 1080, 256, 5)
            base, mut, right = ref_ops.sgbm_param_sets(256, 11, 1)
            sgbm_case(ctx, lg, rg, mut, "c4 m1 wls-mutated", volumes=False)
        ln = rng.integers(0, 256, (96, 400)).astype(np.uint8); rn = rng.integers(0, 256, (96, 400)).astype(np.uint8)
        for mode in (0, 1, 2):
            kw = dict(minDisparity=0, numDisparities=256, blockSize=11, P1=2904, P2=11616, disp12MaxDiff=1000000,
                      preFilterCap=63, uniquenessRatio=0, speckleWindowSize=0, speckleRange=32, mode=mode)
            sgbm_case(ctx, ln, rn, kw, "noise D256 bs11 m%d" % mode, volumes=False)
    guarded(t_sgbm_big)

    def t_post():
        section("median / speckles")
        for t in range(4):
            H, W = (37, 61) if t < 2 else (360, 320)
            d = (rng.integers(-1, 40, (H, W)) * 16 + rng.integers(0, 16, (H, W))).astype(np.int16)
            d[rng.random((H, W)) < 0.3] = -16
            report("median %d" % t, ctx.median3_s16(d), cv2.medianBlur(d, 3))
            e = d.copy(); cv2.filterSpeckles(e, -16, 25, 32)
            report("speckles %d" % t, ctx.filter_speckles(d, -16, 25, 32), e)
    guarded(t_post)

    def t_wls():
        section("WLS vs oracle restatement (parity unpinned vs ximgproc)")
        for (W, H, D, bs, mode) in [(320, 360, 64, 5, 2), (1280, 720, 128, 9, 1)]:
            lg, rg = gray_pair(W, H, D, 9)
            base, mut, right = ref_ops.sgbm_param_sets(D, bs, mode)
            dl = cv2.StereoSGBM_create(**mut).compute(lg, rg)
            dr = cv2.StereoSGBM_create(**right).compute(rg, lg)
            r = int(np.ceil(0.5 * bs))
            want, wconf = cref.wls_filter(dl, dr, lg, 0, D, r, 8000.0, 1.5, want_conf=True)
            p = N.WlsParams(8000.0, 1.5, 0, D, r, 24)
            got, conf = ctx.wls_filter(p, dl, dr, lg, want_conf=True)
            report("wls conf %dx%d" % (W, H), conf, wconf)
            report("wls out  %dx%d" % (W, H), got, want)
            print("     max |diff| = %d LSB, frac>1LSB = %.5f" % (np.abs(got.astype(int) - want).max(),
                                                               (np.abs(got.astype(int) - want) > 1).mean()))
            Q = synth.camera_model(W, H)[1]
            report("depth Q  %dx%d" % (W, H), ctx.disp_to_depth(want, Q), ref_ops.depth_from_disparity(want, Q))
            report("depth def %dx%d" % (W, H), ctx.disp_to_depth(want, None), ref_ops.depth_from_disparity(want, None))
    guarded(t_wls)

    def t_laser():
        section("laser extractors vs reference restatement")
        cfg = dict(hsv_lower=(50, 100, 180), hsv_upper=(70, 255, 255), brightness_threshold=200, min_area=50)
        for (W, H, D) in [(320, 360, 64), (1280, 720, 128)]:
            left, _ = synth.stereo_pair(W, H, D, 1)
            wpts, wm1, wm2 = ref_ops.simple_extract(left, want_masks=True, **cfg)
            pts, m1, m2 = ctx.simple_extract(left, cfg["hsv_lower"], cfg["hsv_upper"], 200, 50, want_masks=True)
            report("simple mask_morph %d" % W, m1, wm1)
            report("simple mask_final %d" % W, m2, wm2)
            report("simple points %d" % W, pts, np.array(wpts).reshape(-1, 2))
            for variant, fn, name in ((0, ref_ops.fast_steger_extract, "fast"), (1, ref_ops.improved_steger_extract, "improved"),
                                      (2, ref_ops.improved_steger_extract_optimized, "optimized"), (3, ref_ops.hybrid_extract, "hybrid")):
                sp = N.StegerParams(variant, 2.0 if variant == 3 else 3.0, 200, 0.5, (C.c_int * 4)(0, 0, 0, 0),
                                    (C.c_int * 3)(50, 100, 180), (C.c_int * 3)(70, 255, 255))
                got = ctx.steger_extract(sp, left)
                want = np.array(fn(left), np.float64).reshape(-1, 2)
                if got.shape == want.shape:
                    err = np.abs(got - want).max() if len(want) else 0.0
                    print("  [%s] steger %-10s %d: %d pts, max err %.2e px" % ("ok" if err <= 0.01 else "FAIL", name, W, len(want), err))
                    if err > 0.01:
                        FAILS.append("steger " + name)
                else:
                    # point-set agreement
                    from scipy.spatial import cKDTree
                    d = cKDTree(want).query(got)[0] if len(want) and len(got) else np.array([1.0])
                    print("  [FAIL?] steger %-10s %d: got %d pts want %d; matched(<=0.01) %.4f" % (name, W, len(got), len(want), (d <= 0.01).mean()))
                    FAILS.append("steger count " + name)
    guarded(t_laser)

    def t_recon():
        section("reconstruction vs reference restatement")
        W, H, D = 320, 360, 64
        left, right = synth.stereo_pair(W, H, D, 2)
        K, Q = synth.camera_model(W, H)
        lrect, depth, aux = ref_ops.depth_path(left, right, None, D, 5, 2, Q, use_wls=False, want_all=True)
        pts = ref_ops.improved_steger_extract(left)
        for refr in (0, 1):
            rp = N.ReconParams(); rp.kind = N.RECON_PLANE
            rp.K[:] = list(K.reshape(9)); rp.plane[:] = list(synth.LASER_PLANE); rp.use_refraction = refr; rp.n_water = 1.33
            got = ctx.reconstruct(rp, pts)
            want = ref_ops.ReconstructorRef(K, synth.LASER_PLANE, bool(refr)).reconstruct_laser_line(pts).reshape(-1, 3)
            ok = got.shape == want.shape and np.allclose(got, want, rtol=1e-9, atol=1e-12)
            print("  [%s] plane refr=%d: %s vs %s, max rel %.2e" % ("ok" if ok else "FAIL", refr, got.shape, want.shape,
                  (np.abs(got - want) / np.maximum(np.abs(want), 1e-12)).max() if got.shape == want.shape and len(want) else -1))
            if not ok:
                FAILS.append("recon plane")
        rp = N.ReconParams(); rp.kind = N.RECON_DEPTH; rp.K[:] = list(K.reshape(9))
        got = ctx.reconstruct(rp, pts, depth)
        want = ref_ops.ReconstructorRef(K, synth.LASER_PLANE, False).reconstruct_from_depth(pts, depth).reshape(-1, 3)
        report("recon from_depth", got, want)
        disp = aux["df"].astype(np.float32) / 16.0
        ir = ref_ops.ImprovedLaserReconstructorRef(Q)
        for kind, fn in ((N.RECON_DISPARITY, lambda: ir.reconstruct_from_disparity(pts, disp)),
                         (N.RECON_DISPARITY_MEDIAN, lambda: ir.reconstruct_with_interpolation(pts, disp, 3, 1.0))):
            rp = N.ReconParams(); rp.kind = kind; rp.fx = ir.fx; rp.baseline = ir.baseline; rp.cx = ir.cx; rp.cy = ir.cy
            rp.min_disparity = 1.0; rp.window = 3
            got = ctx.reconstruct(rp, pts, disp).astype(np.float32)
            report("recon kind %d" % kind, got, fn().reshape(-1, 3))
    guarded(t_recon)

    def t_pipeline():
        section("c3 pipeline timing")
        W, H, D, bs = 1280, 720, 128, 9
        from laser_3d_reconstruction_b200.pipeline import make_pipeline_config
        K, Q = synth.camera_model(W, H)
        nf = 8
        frames = [synth.stereo_pair(W, H, D, s) for s in range(2)]
        L = np.stack([frames[i % 2][0] for i in range(nf)]); R = np.stack([frames[i % 2][1] for i in range(nf)])
        lib = ctx.lib
        dl = lib.l3d_dev_alloc(ctx.h, L.nbytes); dr = lib.l3d_dev_alloc(ctx.h, R.nbytes)
        lib.l3d_memcpy_h2d(ctx.h, dl, L.ctypes.data, L.nbytes); lib.l3d_memcpy_h2d(ctx.h, dr, R.ctypes.data, R.nbytes)
        mlx, mly = synth.warp_maps(W, H, 0); mrx, mry = synth.warp_maps(W, H, 1)
        for lanes in (1, 2, 4):
            cfg = make_pipeline_config(W, H, D, bs, 1, Q, K, extractor=N.STEGER_IMPROVED, lanes=lanes, max_points=20000)
            ph = C.c_void_p()
            ctx.check(lib.l3d_pipeline_create(ctx.h, C.byref(cfg), C.byref(ph)), "pipeline_create")
            ctx.check(lib.l3d_pipeline_set_maps(ph, 0, mlx.ctypes.data_as(C.c_void_p), mly.ctypes.data_as(C.c_void_p)), "set_maps")
            ctx.check(lib.l3d_pipeline_set_maps(ph, 1, mrx.ctypes.data_as(C.c_void_p), mry.ctypes.data_as(C.c_void_p)), "set_maps")
            counts = (C.c_int * nf)()
            lib.l3d_pipeline_set_timing(ph, 1)
            for it in range(3):
                t0 = time.time()
                ctx.check(lib.l3d_pipeline_run_dev(ph, C.c_void_p(dl), C.c_void_p(dr), nf, counts), "run_dev")
                wall = time.time() - t0
                ms = lib.l3d_pipeline_last_ms(ph)
                print("  lanes=%d it=%d: %.2f ms / %d frames = %.3f ms/frame (wall %.1f ms), pts %s" % (lanes, it, ms, nf, ms / nf, wall * 1e3, list(counts)[:3]))
            for name in ("sgbm_cost", "sgbm_scan", "sgbm_wta", "wls"):
                t = C.c_float(); k = C.c_int()
                lib.l3d_pipeline_kernel_time(ph, name.encode(), C.byref(t), C.byref(k))
                print("     %-10s %8.3f ms over %d timed launches -> %.3f ms each" % (name, t.value, k.value, t.value / max(k.value, 1)))
            lib.l3d_pipeline_set_timing(ph, 0)
            ctx.check(lib.l3d_pipeline_run_dev(ph, C.c_void_p(dl), C.c_void_p(dr), nf, counts), "run_dev")
            print("  lanes=%d untimed-events: %.3f ms/frame" % (lanes, lib.l3d_pipeline_last_ms(ph) / nf))
            if lanes == 1:
                # parity of the fused pipeline against the stage-wise reference path for frame 0
                depth = np.empty((H, W), np.float32); rect = np.empty((H, W, 3), np.uint8); disp = np.empty((H, W), np.int16)
                xyz = np.empty((20000, 3)); nx = C.c_int(); nz = C.c_int()
                ctx.check(lib.l3d_pipeline_fetch(ph, 0, rect.ctypes.data_as(C.c_void_p), depth.ctypes.data_as(C.c_void_p),
                                                 disp.ctypes.data_as(C.c_void_p), None, xyz.ctypes.data_as(C.c_void_p),
                                                 C.byref(nx), C.byref(nz)), "fetch")
                wrect, wdepth, aux = ref_ops.depth_path(frames[0][0], frames[0][1], (mlx, mly, mrx, mry), D, bs, 1, Q, want_all=True)
                report("pipeline rect", rect, wrect)
                report("pipeline disp", disp, aux["df"])
                report("pipeline depth", depth, wdepth)
                wp = ref_ops.improved_steger_extract(wrect)
                wxyz = ref_ops.ReconstructorRef(K, synth.LASER_PLANE, False).reconstruct_from_depth(wp, wdepth).reshape(-1, 3)
                print("     points: got %d/%d want %d/%d" % (nx.value, nz.value, len(wp), len(wxyz)))
                if nz.value == len(wxyz) and len(wxyz):
                    print("     xyz max rel err %.2e" % (np.abs(xyz[:nz.value] - wxyz) / np.maximum(np.abs(wxyz), 1e-9)).max())
            lib.l3d_pipeline_destroy(ph)
        lib.l3d_dev_free(ctx.h, dl); lib.l3d_dev_free(ctx.h, dr)
    guarded(t_pipeline)

    print("\nFAILED: %d %s" % (len(FAILS), FAILS[:30]))
    print("launches:", ctx.launches)
    return 1 if FAILS else 0


if __name__ == "__main__":
    sys.exit(main())
