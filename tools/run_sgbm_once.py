"""One c3-sized StereoSGBM run (MODE_HH, WLS-mutated parameters) through the C ABI, twice: the
command profiled by ncu for the SGBM kernels (cost, scan, WTA)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from laser_3d_reconstruction_b200 import _native as N  # noqa: E402

W, H, D, bs = 1280, 720, 128, 9
if len(sys.argv) > 1 and sys.argv[1] == "c1":
    W, H, D, bs = 320, 360, 64, 5
rng = np.random.default_rng(0)
base = rng.integers(0, 256, (H, W + D), dtype=np.uint8)
import cv2  # noqa: E402
base = cv2.GaussianBlur(base, (0, 0), 1.5)
left, right = np.ascontiguousarray(base[:, D // 2:D // 2 + W]), np.ascontiguousarray(base[:, D // 2 + 20:D // 2 + 20 + W])
ctx = N.Context(0)
p = N.SgbmParams(0, D, bs, 24 * bs * bs, 96 * bs * bs, 1000000, 63, 0, 0, 32, 1)
for _ in range(2):
    d = ctx.sgbm_compute(p, left, right)
print("valid fraction", float((d >= 0).mean()), "launches", ctx.launches)
