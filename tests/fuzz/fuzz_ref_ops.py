"""Differential fuzzing of oracle/ref_ops.py (the restatement of the reference's Python classes) against the REAL reference
classes imported from /root/reference (build container only): random frames with blobs, stripes, saturated patches and noise
through every extractor, random points / depth / disparity maps through both reconstructors.

    python tests/fuzz/fuzz_ref_ops.py [seed] [iterations]
"""
import contextlib, io, os, sys
import cv2
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, "/root/reference")
from oracle import ref_ops


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def pts(a):
    return np.array(a, np.float64).reshape(-1, 2)


def frame(rng, W, H):
    img = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), rng.uniform(0.8, 3.0))
    img = np.clip((img.astype(np.float32) - 128) * rng.uniform(1.0, 3.0) + rng.uniform(60, 160), 0, 255).astype(np.uint8)
    for _ in range(int(rng.integers(1, 5))):   # green-ish stripes / blobs, some saturated
        col = (int(rng.integers(0, 200)), int(rng.integers(180, 256)), int(rng.integers(0, 200)))
        if rng.random() < 0.6:
            x0, x1 = rng.integers(0, W, 2)
            cv2.line(img, (int(x0), 0), (int(x1), H - 1), col, int(rng.integers(1, 9)))
        else:
            cv2.circle(img, (int(rng.integers(0, W)), int(rng.integers(0, H))), int(rng.integers(2, 25)), col, -1)
    if rng.random() < 0.5:
        img = cv2.GaussianBlur(img, (0, 0), rng.uniform(0.5, 2.5))
    return img


def main():
    from core.laser_extractor import FastStegerExtractor, SimpleLaserExtractor
    from core.reconstruction import Reconstructor
    from improved_reconstruction import ImprovedLaserReconstructor
    from improved_steger import HybridLaserExtractor, ImprovedStegerExtractor
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    rng = np.random.default_rng(seed)
    bad = 0

    def check(name, got, want, tol=0.0, ctx=""):
        nonlocal bad
        got, want = np.asarray(got), np.asarray(want)
        ok = got.shape == want.shape and (np.array_equal(got, want) if tol == 0 else
                                          (got.size == 0 or np.nanmax(np.abs(got.astype(np.float64) - want.astype(np.float64))) <= tol))
        if not ok:
            bad += 1
            print("MISMATCH", name, ctx, got.shape, want.shape,
                  "" if got.shape != want.shape or got.size == 0 else float(np.nanmax(np.abs(got.astype(np.float64) - want))))

    for it in range(n):
        W, H = int(rng.integers(40, 200)), int(rng.integers(30, 140))
        img = frame(rng, W, H)
        thr = int(rng.choice([100, 150, 200, 230]))
        lo = [int(rng.integers(30, 60)), int(rng.integers(20, 120)), int(rng.integers(80, 200))]
        hi = [int(rng.integers(60, 95)), 255, 255]
        area = int(rng.choice([0, 5, 50, 200]))
        ctx = "it %d %dx%d thr %d" % (it, W, H, thr)
        s = quiet(SimpleLaserExtractor, hsv_lower=np.array(lo), hsv_upper=np.array(hi), brightness_threshold=thr, min_area=area)
        check("simple", pts(ref_ops.simple_extract(img, lo, hi, thr, area)), pts(quiet(s.extract_centerline, img)), 0.0, ctx)
        sigma = float(rng.choice([1.0, 2.0, 3.0]))
        f = quiet(FastStegerExtractor, sigma=sigma, brightness_threshold=thr)
        roi = None if rng.random() < 0.6 else (int(rng.integers(0, W // 2)), int(rng.integers(0, H // 2)), int(rng.integers(5, W)), int(rng.integers(5, H)))
        src = img if rng.random() < 0.7 else cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        check("fast", pts(ref_ops.fast_steger_extract(src, sigma, thr, roi)), pts(quiet(f.extract_centerline, src, roi)), 1e-3, ctx)
        resp = float(rng.choice([0.5, 0.3, 1.0]))
        im = quiet(ImprovedStegerExtractor, sigma=sigma, brightness_threshold=thr, response_threshold=resp)
        check("improved", pts(ref_ops.improved_steger_extract(src, sigma, thr, resp)), pts(quiet(im.extract_centerline, src)), 1e-3, ctx)
        check("optimized", pts(ref_ops.improved_steger_extract_optimized(src, sigma, thr, resp)), pts(quiet(im.extract_centerline_optimized, src)), 1e-3, ctx)
        hy = quiet(HybridLaserExtractor, hsv_lower=np.array(lo), hsv_upper=np.array(hi), brightness_threshold=thr, sigma=sigma)
        check("hybrid", pts(ref_ops.hybrid_extract(img, lo, hi, thr, sigma)), pts(quiet(hy.extract_centerline, img)), 1e-3, ctx)
        # reconstructors
        K = np.array([[rng.uniform(0.5, 1.5) * W, 0, W / 2 + rng.uniform(-5, 5)], [0, rng.uniform(0.5, 1.5) * W, H / 2 + rng.uniform(-5, 5)], [0, 0, 1]])
        plane = np.array([rng.uniform(-1, 1), rng.uniform(-0.3, 0.3), rng.uniform(-1.5, 1.0), rng.uniform(-0.6, 0.6)])
        p2 = [(float(rng.uniform(-5, W + 5)), float(rng.uniform(-5, H + 5))) for _ in range(60)]
        depth = (rng.random((H, W)) * 3000).astype(np.float32)
        depth[rng.random((H, W)) < 0.3] = 0
        for refr in (False, True):
            a = quiet(Reconstructor, K, plane, refr); b = ref_ops.ReconstructorRef(K, plane, refr)
            with np.errstate(all="ignore"):
                check("laser_line refr=%s" % refr, b.reconstruct_laser_line(p2), quiet(a.reconstruct_laser_line, p2), 1e-9, ctx)
                check("from_depth", b.reconstruct_from_depth(p2, depth), quiet(a.reconstruct_from_depth, p2, depth), 0.0, ctx)
        Q = np.array([[1, 0, 0, -W / 2], [0, 1, 0, -H / 2], [0, 0, 0, 0.8 * W], [0, 0, 1 / 0.06, 0]])
        disp = (rng.random((H, W)) * 64).astype(np.float32)
        disp[rng.random((H, W)) < 0.2] = 0
        disp[rng.random((H, W)) < 0.02] = np.nan
        disp[rng.random((H, W)) < 0.02] = np.inf
        a = quiet(ImprovedLaserReconstructor, Q); b = ref_ops.ImprovedLaserReconstructorRef(Q)
        md = float(rng.choice([1.0, 0.5, 8.0]))
        with np.errstate(all="ignore"):
            check("from_disparity", b.reconstruct_from_disparity(p2, disp, md), quiet(a.reconstruct_from_disparity, p2, disp, md), 0.0, ctx)
            check("interp", b.reconstruct_with_interpolation(p2, disp, 3, md), quiet(a.reconstruct_with_interpolation, p2, disp, 3, md), 0.0, ctx)
    print("done bad =", bad)


if __name__ == "__main__":
    main()
