"""SimpleLaserExtractor / FastStegerExtractor with the reference's API (core/laser_extractor.py);
the per-pixel work runs in libl3d.so (csrc/laser.cu).  No CPU fallback."""
import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from .. import _native as N


def _steger_params(variant, sigma, bright_thr, resp_thr=0.5, roi=None, hsv_lo=(0, 0, 0), hsv_hi=(0, 0, 0)):
    r = (0, 0, 0, 0) if roi is None else tuple(int(v) for v in roi)
    return N.StegerParams(variant, float(sigma), int(bright_thr), float(resp_thr), (C.c_int * 4)(*r),
                          (C.c_int * 3)(*[int(v) for v in hsv_lo]), (C.c_int * 3)(*[int(v) for v in hsv_hi]))


def _check_image(image, allow_gray):
    image = np.asarray(image)
    if image.dtype != np.uint8:
        raise TypeError("laser extractors expect uint8 images")
    if image.ndim == 2 and allow_gray:
        return image
    if image.ndim == 3 and image.shape[2] == 3:
        return image
    raise ValueError("expected an HxWx3 BGR image" + (" or an HxW gray image" if allow_gray else ""))


class SimpleLaserExtractor:
    """core/laser_extractor.py:14-100: HSV + brightness mask -> CLOSE/OPEN -> contour-area filter ->
    brightness-weighted centroid per row."""

    def __init__(self, hsv_lower=None, hsv_upper=None, brightness_threshold=100, min_area=50, device=0, verbose=True):
        self.hsv_lower = hsv_lower if hsv_lower is not None else np.array([40, 50, 100])
        self.hsv_upper = hsv_upper if hsv_upper is not None else np.array([80, 255, 255])
        self.brightness_threshold = brightness_threshold
        self.min_area = min_area
        self.device = device
        if verbose:
            print(f"SimpleLaserExtractor 初始化: HSV范围 {self.hsv_lower} ~ {self.hsv_upper}, "
                  f"亮度阈值 {self.brightness_threshold}, 最小面积 {self.min_area}")

    def extract_centerline(self, image: np.ndarray) -> List[Tuple[float, float]]:
        image = _check_image(image, allow_gray=False)
        pts = N.default_context(self.device).simple_extract(image, self.hsv_lower, self.hsv_upper,
                                                            self.brightness_threshold, self.min_area)
        return N.points_to_list(pts)  # python floats, built at C speed

    def extract_masks(self, image: np.ndarray):
        """(mask after morphology :69, final contour mask :81-82, points) -- for parity tests."""
        image = _check_image(image, allow_gray=False)
        pts, m1, m2 = N.default_context(self.device).simple_extract(
            image, self.hsv_lower, self.hsv_upper, self.brightness_threshold, self.min_area, want_masks=True)
        return m1, m2, N.points_to_list(pts)  # python floats, built at C speed


class FastStegerExtractor:
    """core/laser_extractor.py:103-285: Gaussian sigma -> 2-/3-tap differences -> 2x2 Hessian
    eigen-analysis on every pixel brighter than the threshold -> sub-pixel offset."""

    def __init__(self, sigma: float = 3.0, brightness_threshold: int = 200, use_lut: bool = True, device=0,
                 verbose=True):
        self.sigma = sigma
        self.brightness_threshold = brightness_threshold
        self.use_lut = use_lut
        self.lut = None  # the reference builds a LUT it never reads (:138-144)
        self.kernel_size = int(2 * np.ceil(3 * sigma) + 1)
        self.device = device
        if verbose:
            print(f"FastStegerExtractor 初始化: Sigma {sigma}, 亮度阈值 {brightness_threshold}, 核大小 {self.kernel_size}")

    def extract_centerline(self, image: np.ndarray, roi: Optional[Tuple[int, int, int, int]] = None) -> List[
            Tuple[float, float]]:
        image = _check_image(image, allow_gray=True)
        if roi is not None:
            x, y, w, h = [int(v) for v in roi]
            H, W = image.shape[:2]
            # numpy slicing clamps; an empty ROI yields no points
            x0, y0 = max(x, 0), max(y, 0)
            if x < 0 or y < 0:
                raise ValueError("roi origin must be non-negative")
            w, h = min(w, W - x0), min(h, H - y0)
            if w <= 0 or h <= 0:
                return []
            roi = (x0, y0, w, h)
        p = _steger_params(N.STEGER_FAST, self.sigma, self.brightness_threshold, roi=roi)
        pts = N.default_context(self.device).steger_extract(p, image)
        return N.points_to_list_f32(pts)  # np.float32 pairs, like the reference

    def extract_batch(self, images: List[np.ndarray]) -> List[List[Tuple[float, float]]]:
        """core/laser_extractor.py:279-285."""
        return [self.extract_centerline(img) for img in images]
