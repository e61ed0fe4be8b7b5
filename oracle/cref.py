"""ctypes front-end of the oracle's C restatement (oracle/csrc/*.c -> oracle/_ref/liborc.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package (laser_3d_reconstruction_b200) never imports this module.
"""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


class SgbmParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff",
        "preFilterCap", "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")]


class BmParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "minDisparity", "numDisparities", "blockSize", "preFilterCap", "textureThreshold", "uniquenessRatio",
        "speckleWindowSize", "speckleRange", "disp12MaxDiff")]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.orc_sgbm_compute.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def bgr2gray(bgr):
    bgr = _c(bgr, np.uint8)
    out = np.empty(bgr.shape[:2], np.uint8)
    lib().orc_bgr2gray(_p(bgr), C.c_long(out.size), _p(out))
    return out


def bgr2hsv(bgr):
    bgr = _c(bgr, np.uint8)
    out = np.empty_like(bgr)
    lib().orc_bgr2hsv(_p(bgr), C.c_long(out.size // 3), _p(out))
    return out


def remap_bilinear(src, mapx, mapy):
    src = _c(src, np.uint8)
    mapx, mapy = _c(mapx, np.float32), _c(mapy, np.float32)
    cn = 1 if src.ndim == 2 else src.shape[2]
    dh, dw = mapx.shape
    out = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, cn), np.uint8)
    lib().orc_remap_bilinear(_p(src), src.shape[1], src.shape[0], cn, _p(mapx), _p(mapy), dw, dh, _p(out))
    return out


def sgbm_compute(left, right, want_volumes=False, want_raw=False, **kw):
    """kw: cv2.StereoSGBM_create keyword names.  Returns disp16 (and optionally raw, C, S)."""
    left, right = _c(left, np.uint8), _c(right, np.uint8)
    H, W = left.shape
    p = SgbmParams(**kw)
    disp = np.empty((H, W), np.int16)
    raw = np.empty((H, W), np.int16) if want_raw else None
    Cv = Sv = None
    if want_volumes:
        minD, D = p.minDisparity, p.numDisparities
        width1 = (W + min(minD, 0)) - max(minD + D, 0)
        Cv = np.zeros((H, max(width1, 0), D), np.int16)
        Sv = np.zeros_like(Cv)
    rc = lib().orc_sgbm_compute(_p(left), _p(right), W, H, C.byref(p), _p(disp), _p(raw), _p(Cv), _p(Sv))
    if rc != 0:
        raise ValueError("orc_sgbm_compute failed rc=%d" % rc)
    res = [disp]
    if want_raw:
        res.append(raw)
    if want_volumes:
        res += [Cv, Sv]
    return res[0] if len(res) == 1 else tuple(res)


def bm_compute(left, right, numDisparities=64, blockSize=15, minDisparity=0, preFilterCap=31, textureThreshold=10,
               uniquenessRatio=15, speckleWindowSize=0, speckleRange=0, disp12MaxDiff=-1):
    """cv2.StereoBM_create(numDisparities, blockSize).compute(left, right) with the setters' parameters."""
    left, right = _c(left, np.uint8), _c(right, np.uint8)
    H, W = left.shape
    p = BmParams(minDisparity, numDisparities, blockSize, preFilterCap, textureThreshold, uniquenessRatio,
                 speckleWindowSize, speckleRange, disp12MaxDiff)
    disp = np.empty((H, W), np.int16)
    rc = lib().orc_bm_compute(_p(left), _p(right), W, H, C.byref(p), _p(disp))
    if rc != 0:
        raise ValueError("orc_bm_compute: unsupported parameters")
    return disp


def bm_prefilter(img, cap=31):
    img = _c(img, np.uint8)
    out = np.empty_like(img)
    lib().orc_bm_prefilter_xsobel(_p(img), img.shape[1], img.shape[0], int(cap), _p(out))
    return out


def median3_s16(a):
    a = _c(a, np.int16)
    out = np.empty_like(a)
    lib().orc_median3_s16(_p(a), a.shape[1], a.shape[0], _p(out))
    return out


def filter_speckles(a, new_val, max_size, max_diff):
    a = _c(a, np.int16).copy()
    lib().orc_filter_speckles(_p(a), a.shape[1], a.shape[0], int(new_val), int(max_size), int(max_diff))
    return a


def wls_filter(dl, dr, guide, min_disp, num_disp, dd_radius, lam=8000.0, sigma_color=1.5,
               lrc_thresh=24, want_conf=False, variant=0):
    dl, dr, guide = _c(dl, np.int16), _c(dr, np.int16), _c(guide, np.uint8)
    H, W = dl.shape
    out = np.empty((H, W), np.int16)
    conf = np.empty((H, W), np.float32) if want_conf else None
    lib().orc_wls_filter_v(_p(dl), _p(dr), _p(guide), W, H, int(min_disp), int(num_disp), int(dd_radius),
                           C.c_double(lam), C.c_double(sigma_color), int(lrc_thresh), int(variant), _p(out), _p(conf))
    return (out, conf) if want_conf else out


def disp_to_depth_q(disp16, Q):
    disp16 = _c(disp16, np.int16)
    Q = _c(Q, np.float64)
    out = np.empty(disp16.shape, np.float32)
    lib().orc_disp_to_depth_q(_p(disp16), disp16.shape[1], disp16.shape[0], _p(Q), _p(out))
    return out


def disp_to_depth_default(disp16):
    disp16 = _c(disp16, np.int16)
    out = np.empty(disp16.shape, np.float32)
    lib().orc_disp_to_depth_default(_p(disp16), disp16.shape[1], disp16.shape[0], _p(out))
    return out


def simple_masks(bgr, hsv_lo, hsv_hi, bright_thr, min_area):
    bgr = _c(bgr, np.uint8)
    H, W = bgr.shape[:2]
    lo = (C.c_int * 3)(*[int(v) for v in hsv_lo])
    hi = (C.c_int * 3)(*[int(v) for v in hsv_hi])
    m1 = np.empty((H, W), np.uint8)
    m2 = np.empty((H, W), np.uint8)
    lib().orc_simple_masks(_p(bgr), W, H, lo, hi, int(bright_thr), C.c_double(min_area), _p(m1), _p(m2))
    return m1, m2


def close_open3(mask):
    mask = _c(mask, np.uint8)
    out = np.empty_like(mask)
    lib().orc_close_open3(_p(mask), mask.shape[1], mask.shape[0], _p(out))
    return out


def gaussian_blur_f32(img, sigma):
    img = _c(img, np.float32)
    out = np.empty_like(img)
    lib().orc_gaussian_blur_f32(_p(img), img.shape[1], img.shape[0], C.c_double(sigma), _p(out))
    return out


def sobel3_f32(img, dx, dy):
    img = _c(img, np.float32)
    out = np.empty_like(img)
    lib().orc_sobel3_f32(_p(img), img.shape[1], img.shape[0], int(dx), int(dy), _p(out))
    return out
