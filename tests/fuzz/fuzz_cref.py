"""Differential fuzzing of the C restatement (oracle/csrc via oracle/cref.py) against the installed cv2 binary over random
sizes and parameters: remap, median / speckle filters, reprojectImageTo3D depth, Simple-extractor masks (inRange + CLOSE +
OPEN + contour fill), Gaussian / Sobel float filters.  SGBM: tests/fuzz/fuzz_oracle.py; stereoRectify: tests/fuzz/fuzz_rectify.py.

    python tests/fuzz/fuzz_cref.py [seed] [iterations]
"""
import os, sys
import cv2
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cref

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rng = np.random.default_rng(seed)
bad = 0


def check(name, got, want, ctx="", tol=None):
    global bad
    ok = np.array_equal(got, want) if tol is None else float(np.abs(got.astype(np.float64) - want).max()) <= tol
    if not ok:
        bad += 1
        d = np.abs(got.astype(np.float64) - want.astype(np.float64))
        print("MISMATCH", name, ctx, "n", int((d > (tol or 0)).sum()), "max", float(d.max()))


K3 = np.ones((3, 3), np.uint8)
for it in range(n):
    W, H = int(rng.integers(3, 220)), int(rng.integers(3, 160))
    ctx = "it %d %dx%d" % (it, W, H)
    # remap: random maps reaching outside, exact fixed-point half-way values, constant maps
    src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    kind = int(rng.integers(0, 4))
    mx = (rng.random((H, W)) * (W + 10) - 5).astype(np.float32)
    my = (rng.random((H, W)) * (H + 10) - 5).astype(np.float32)
    if kind == 1:
        mx, my = (np.round(mx * 64) / 64).astype(np.float32), (np.round(my * 64) / 64).astype(np.float32)
    if kind == 2:
        mx, my = np.round(mx).astype(np.float32), np.round(my).astype(np.float32)
    if kind == 3:
        mx = (np.arange(W, dtype=np.float32)[None, :] + np.float32(rng.uniform(-2, 2))).repeat(H, 0)
        my = (np.arange(H, dtype=np.float32)[:, None] + np.float32(rng.uniform(-2, 2))).repeat(W, 1)
    check("remap", cref.remap_bilinear(src, mx, my), cv2.remap(src, mx, my, cv2.INTER_LINEAR), ctx + " kind %d" % kind)
    # median3 + speckles on int16 disparity-like images with plateaus and spikes
    d16 = (rng.integers(-16, 2048, (H, W)) // int(rng.choice([1, 16, 64])) * int(rng.choice([1, 16]))).astype(np.int16)
    if W >= 3 and H >= 3:
        check("median3", cref.median3_s16(d16), cv2.medianBlur(d16, 3), ctx)
    ms, md, nv = int(rng.choice([0, 1, 5, 50, 400])), int(rng.choice([0, 1, 16, 32, 512])), int(rng.choice([-16, 0, -1]))
    want = d16.copy()
    cv2.filterSpeckles(want, nv, ms, md)
    check("speckles", cref.filter_speckles(d16, nv, ms, md), want, ctx + " ms %d md %d" % (ms, md))
    # depth through reprojectImageTo3D with a random Q and the reference's three masks
    Q = np.array([[1, 0, 0, -rng.uniform(0, W)], [0, 1, 0, -rng.uniform(0, H)], [0, 0, 0, rng.uniform(50, 900)],
                  [0, 0, rng.uniform(2, 40) * (1 if rng.random() < 0.8 else -1), rng.uniform(-3, 3)]])
    disp = d16.astype(np.float32) / 16.0
    p3 = cv2.reprojectImageTo3D(disp, Q)[:, :, 2].copy()
    p3[p3 < 0] = 0; p3[p3 > 10] = 0; p3[disp <= 0] = 0
    check("depth_q", cref.disp_to_depth_q(d16, Q), p3, ctx)
    # Simple-extractor masks
    img = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), rng.uniform(0.6, 4.0))
    img = np.clip((img.astype(np.float32) - 128) * rng.uniform(1.5, 5.0) + 128, 0, 255).astype(np.uint8)
    lo = [int(rng.integers(0, 90)), int(rng.integers(0, 150)), int(rng.integers(0, 200))]
    hi = [int(rng.integers(lo[0], 181)), int(rng.integers(lo[1], 256)), int(rng.integers(lo[2], 256))]
    thr, area = int(rng.integers(0, 220)), float(rng.choice([0, 1.5, 10, 50]))
    hsv = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    m = cv2.bitwise_and(cv2.inRange(hsv, np.array(lo), np.array(hi)), ((cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) > thr) * 255).astype(np.uint8))
    m1 = cv2.morphologyEx(cv2.morphologyEx(m, cv2.MORPH_CLOSE, K3), cv2.MORPH_OPEN, K3)
    cs, _ = cv2.findContours(m1, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    m2 = np.zeros_like(m1)
    for c in cs:
        if cv2.contourArea(c) > area:
            cv2.drawContours(m2, [c], -1, 255, -1)
    g1, g2 = cref.simple_masks(img, lo, hi, thr, area)
    check("simple m1", g1, m1, ctx)
    check("simple m2", g2, m2, ctx + " area %g" % area)
    # float filters of the Steger chain
    if W >= 8 and H >= 8:
        gimg = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY).astype(np.float32)
        sg = float(rng.choice([1.0, 1.5, 2.0, 3.0]))
        gb = cv2.GaussianBlur(gimg, (0, 0), sg)
        check("gauss", cref.gaussian_blur_f32(gimg, sg), gb, ctx + " sigma %g" % sg, tol=2e-4)
        for dx, dy in ((1, 0), (0, 1)):
            check("sobel", cref.sobel3_f32(gb, dx, dy), cv2.Sobel(gb, cv2.CV_32F, dx, dy, ksize=3), ctx, tol=1e-3)
print("done bad =", bad)
