"""cv2.StereoSGBM / cv2.ximgproc look-alikes backed by libl3d.so.

The reference touches these objects directly (``camera.stereo_matcher.compute(lgray, rgray)``,
test_improved_laser.py:151, test_depth.py:68; ``wls_filter.filter(...)``,
camera/single_usb_stereo_camera.py:328-332), so the camera class exposes the same objects with the
same methods.  All arithmetic runs in the sm_100a kernels; nothing here calls cv2.
"""
import math

import numpy as np

from . import _native as N

STEREO_SGBM_MODE_SGBM = 0
STEREO_SGBM_MODE_HH = 1
STEREO_SGBM_MODE_SGBM_3WAY = 2
STEREO_SGBM_MODE_HH4 = 3

_FIELDS = ("minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff", "preFilterCap",
           "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")


class StereoSGBM:
    """Same keyword arguments, defaults, getters/setters and ``compute`` as cv2.StereoSGBM_create
    (camera/single_usb_stereo_camera.py:252-274)."""

    def __init__(self, minDisparity=0, numDisparities=16, blockSize=3, P1=0, P2=0, disp12MaxDiff=0,
                 preFilterCap=0, uniquenessRatio=0, speckleWindowSize=0, speckleRange=0,
                 mode=STEREO_SGBM_MODE_SGBM, device=0):
        self._p = dict(minDisparity=int(minDisparity), numDisparities=int(numDisparities), blockSize=int(blockSize),
                       P1=int(P1), P2=int(P2), disp12MaxDiff=int(disp12MaxDiff), preFilterCap=int(preFilterCap),
                       uniquenessRatio=int(uniquenessRatio), speckleWindowSize=int(speckleWindowSize),
                       speckleRange=int(speckleRange), mode=int(mode))
        self.device = device

    @classmethod
    def create(cls, **kw):
        return cls(**kw)

    def params(self):
        return N.SgbmParams(**self._p)

    def compute(self, left, right):
        """left/right: single-channel uint8 HxW -> int16 HxW disparity x16, invalid = (minD-1)*16."""
        left, right = np.asarray(left), np.asarray(right)
        if left.ndim == 3 and left.shape[2] == 1:
            left, right = left[:, :, 0], right[:, :, 0]
        if left.dtype != np.uint8 or right.dtype != np.uint8:
            raise TypeError("StereoSGBM.compute expects uint8 images")
        return N.default_context(self.device).sgbm_compute(self.params(), left, right)


def _add_accessors():
    for f in _FIELDS:
        cap = f[0].upper() + f[1:]
        setattr(StereoSGBM, "get" + cap, (lambda f: lambda self: self._p[f])(f))
        setattr(StereoSGBM, "set" + cap, (lambda f: lambda self, v: self._p.__setitem__(f, int(v)))(f))


_add_accessors()


def StereoSGBM_create(**kw):
    return StereoSGBM(**kw)


_BM_FIELDS = ("minDisparity", "numDisparities", "blockSize", "preFilterCap", "textureThreshold", "uniquenessRatio",
              "speckleWindowSize", "speckleRange", "disp12MaxDiff")


class StereoBM:
    """cv2.StereoBM_create(numDisparities, blockSize) look-alike (the matcher readme.md:392-397 suggests for speed):
    same defaults (preFilterCap 31, textureThreshold 10, uniquenessRatio 15, no speckle filter, disp12MaxDiff -1),
    the same getters / setters and ``compute``.  PREFILTER_XSOBEL only; minDisparity <= 0 only."""

    def __init__(self, numDisparities=0, blockSize=21, device=0):
        self._p = dict(minDisparity=0, numDisparities=int(numDisparities) if numDisparities else 64, blockSize=int(blockSize),
                       preFilterCap=31, textureThreshold=10, uniquenessRatio=15, speckleWindowSize=0, speckleRange=0,
                       disp12MaxDiff=-1)
        self.device = device

    def params(self):
        return N.BmParams(**self._p)

    # cv2.StereoBM's remaining accessors.  What the kernels do not implement is refused when it is asked for, not ignored.
    _pre_filter_size, _smaller_block_size, _roi1, _roi2 = 9, 0, (0, 0, 0, 0), (0, 0, 0, 0)

    def getPreFilterType(self):
        return 1  # cv2.StereoBM_PREFILTER_XSOBEL, OpenCV's default

    def setPreFilterType(self, v):
        if int(v) != 1:
            raise ValueError("StereoBM: only PREFILTER_XSOBEL (1) is implemented")

    def getPreFilterSize(self):
        return self._pre_filter_size

    def setPreFilterSize(self, v):   # the window of PREFILTER_NORMALIZED_RESPONSE; XSOBEL does not use it
        self._pre_filter_size = int(v)

    def getSmallerBlockSize(self):
        return self._smaller_block_size

    def setSmallerBlockSize(self, v):   # stored by OpenCV, read by nothing
        self._smaller_block_size = int(v)

    def getROI1(self):
        return self._roi1

    def setROI1(self, r):
        self._roi1 = tuple(int(v) for v in r)

    def getROI2(self):
        return self._roi2

    def setROI2(self, r):
        self._roi2 = tuple(int(v) for v in r)

    def compute(self, left, right):
        """left/right: single-channel uint8 HxW -> int16 HxW disparity x16, invalid = (minD-1)*16."""
        left, right = np.asarray(left), np.asarray(right)
        if left.dtype != np.uint8 or right.dtype != np.uint8:
            raise TypeError("StereoBM.compute expects uint8 images")
        if (self._roi1[2] > 0 and self._roi1[3] > 0) or (self._roi2[2] > 0 and self._roi2[3] > 0):
            raise ValueError("StereoBM: valid-disparity rectangles from ROI1 / ROI2 are not implemented (leave them empty)")
        return N.default_context(self.device).bm_compute(self.params(), left, right)


def _add_bm_accessors():
    for f in _BM_FIELDS:
        cap = f[0].upper() + f[1:]
        setattr(StereoBM, "get" + cap, (lambda f: lambda self: self._p[f])(f))
        setattr(StereoBM, "set" + cap, (lambda f: lambda self, v: self._p.__setitem__(f, int(v)))(f))


_add_bm_accessors()


def StereoBM_create(numDisparities=0, blockSize=21, device=0):
    return StereoBM(numDisparities, blockSize, device)


def createRightMatcher(matcher_left):
    """cv2.ximgproc.createRightMatcher for a StereoSGBM (camera/single_usb_stereo_camera.py:277):
    minDisparity = -(minD + numD) + 1, uniqueness 0, disp12MaxDiff 1e6, speckle off."""
    p = matcher_left._p
    return StereoSGBM(minDisparity=-(p["minDisparity"] + p["numDisparities"]) + 1, numDisparities=p["numDisparities"],
                      blockSize=p["blockSize"], P1=p["P1"], P2=p["P2"], disp12MaxDiff=1000000,
                      preFilterCap=p["preFilterCap"], uniquenessRatio=0, speckleWindowSize=0, speckleRange=0,
                      mode=p["mode"], device=matcher_left.device)


class DisparityWLSFilter:
    """cv2.ximgproc.DisparityWLSFilter (lambda 8000, sigma_color 1.0 by default; the reference sets
    8000 / 1.5, camera/single_usb_stereo_camera.py:280-282).  PARITY UNPINNED: cv2.ximgproc is not
    available in the build image; the arithmetic follows oracle/csrc/orc_wls.c."""

    def __init__(self, matcher_left):
        p = matcher_left._p
        self.device = matcher_left.device
        self._min_disp = p["minDisparity"]
        self._num_disp = p["numDisparities"]
        self._dd_radius = int(math.ceil(0.5 * p["blockSize"]))
        self._lambda = 8000.0
        self._sigma = 1.0
        self._lrc = 24
        self._conf = None
        self._roi = (0, 0, 0, 0)

    def setLambda(self, v):
        self._lambda = float(v)

    def getLambda(self):
        return self._lambda

    def setSigmaColor(self, v):
        self._sigma = float(v)

    def getSigmaColor(self):
        return self._sigma

    def setLRCthresh(self, v):
        self._lrc = int(v)

    def getLRCthresh(self):
        return self._lrc

    def setDepthDiscontinuityRadius(self, v):
        self._dd_radius = int(v)

    def getDepthDiscontinuityRadius(self):
        return self._dd_radius

    def getConfidenceMap(self):
        return self._conf

    def getROI(self):
        """(x, y, w, h) of the region the last `filter` call processed: the columns from minDisparity + numDisparities on
        (the restatement's reading of ximgproc's valid-disparity ROI for a StereoSGBM, oracle/csrc/orc_wls.c:145); (0, 0, 0, 0)
        before the first call."""
        return self._roi

    def params(self):
        return N.WlsParams(self._lambda, self._sigma, self._min_disp, self._num_disp, self._dd_radius, self._lrc)

    def filter(self, disparity_map_left, left_view, filtered_disparity_map=None, disparity_map_right=None, ROI=None,
               right_view=None):
        if disparity_map_right is None:
            raise ValueError("DisparityWLSFilter.filter: disparity_map_right is required (confidence mode)")
        guide = np.asarray(left_view)
        if guide.ndim == 3:
            raise ValueError("DisparityWLSFilter.filter: the reference passes a single-channel guide (left_gray)")
        out, conf = N.default_context(self.device).wls_filter(self.params(), disparity_map_left, disparity_map_right,
                                                              guide, want_conf=True)
        self._conf = conf
        H, W = np.asarray(disparity_map_left).shape[:2]
        x0 = min(max(self._min_disp + self._num_disp, 0), W)
        self._roi = (x0, 0, W - x0, H)
        return out


def createDisparityWLSFilter(matcher_left):
    """cv2.ximgproc.createDisparityWLSFilter: MUTATES the left matcher (disp12MaxDiff = 1e6,
    speckleWindowSize = 0, uniquenessRatio = 0) exactly as OpenCV does."""
    f = DisparityWLSFilter(matcher_left)
    matcher_left._p["disp12MaxDiff"] = 1000000
    matcher_left._p["speckleWindowSize"] = 0
    matcher_left._p["uniquenessRatio"] = 0
    return f
