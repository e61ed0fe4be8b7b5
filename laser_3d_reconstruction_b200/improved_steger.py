"""ImprovedStegerExtractor / HybridLaserExtractor with the reference's API (improved_steger.py);
Gaussian + Sobel chain + ridge test run in libl3d.so (csrc/laser.cu).  No CPU fallback."""
from typing import List, Tuple

import numpy as np

from . import _native as N
from .core.laser_extractor import _check_image, _steger_params


class ImprovedStegerExtractor:
    """improved_steger.py:12-223."""

    def __init__(self, sigma: float = 3.0, brightness_threshold: int = 200, response_threshold: float = 0.5,
                 device=0, verbose=True):
        self.sigma = sigma
        self.brightness_threshold = brightness_threshold
        self.response_threshold = response_threshold
        self.device = device
        if verbose:
            print(f"ImprovedStegerExtractor 初始化: Sigma {sigma}, 亮度阈值 {brightness_threshold}, "
                  f"响应阈值 {response_threshold}")

    def _run(self, variant, image):
        image = _check_image(image, allow_gray=True)
        p = _steger_params(variant, self.sigma, self.brightness_threshold, self.response_threshold)
        pts = N.default_context(self.device).steger_extract(p, image)
        return N.points_to_list(pts)  # python floats, built at C speed

    def extract_centerline(self, image: np.ndarray) -> List[Tuple[float, float]]:
        """improved_steger.py:39-126: every ridge pixel, raster order."""
        return self._run(N.STEGER_IMPROVED, image)

    def extract_centerline_optimized(self, image: np.ndarray) -> List[Tuple[float, float]]:
        """improved_steger.py:128-223: the strongest ridge pixel of each row."""
        return self._run(N.STEGER_OPTIMIZED, image)


StegerLaserExtractor = ImprovedStegerExtractor  # the name BASELINE.json's north_star uses


class HybridLaserExtractor:
    """improved_steger.py:226-344: HSV mask AND Steger (sigma 2), strongest ridge pixel per row."""

    def __init__(self, hsv_lower=None, hsv_upper=None, brightness_threshold: int = 200, sigma: float = 2.0,
                 device=0, verbose=True):
        self.hsv_lower = np.array(hsv_lower if hsv_lower is not None else [50, 100, 180])
        self.hsv_upper = np.array(hsv_upper if hsv_upper is not None else [70, 255, 255])
        self.brightness_threshold = brightness_threshold
        self.sigma = sigma
        self.device = device
        if verbose:
            print(f"HybridLaserExtractor 初始化: HSV {self.hsv_lower} ~ {self.hsv_upper}, Sigma {sigma}")

    def extract_centerline(self, image: np.ndarray) -> List[Tuple[float, float]]:
        image = _check_image(image, allow_gray=False)
        p = _steger_params(N.STEGER_HYBRID, self.sigma, self.brightness_threshold, 0.5, None, self.hsv_lower,
                           self.hsv_upper)
        pts = N.default_context(self.device).steger_extract(p, image)
        return N.points_to_list(pts)  # python floats, built at C speed
