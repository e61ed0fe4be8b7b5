"""ImprovedLaserReconstructor with the reference's API (improved_reconstruction.py): laser pixels +
disparity map -> 3D, per-point work in libl3d.so (csrc/recon.cu).  No CPU fallback."""
from typing import List, Tuple

import numpy as np

from . import _native as N


class ImprovedLaserReconstructor:
    def __init__(self, Q_matrix: np.ndarray, device=0, verbose=True):
        """:17-35: fx = Q[2,3], baseline = 1/Q[3,2], cx = -Q[0,3], cy = -Q[1,3]."""
        self.Q = Q_matrix
        self.fx = Q_matrix[2, 3]
        self.baseline = 1.0 / Q_matrix[3, 2]
        self.cx = -Q_matrix[0, 3]
        self.cy = -Q_matrix[1, 3]
        self.device = device
        if verbose:
            print(f"ImprovedLaserReconstructor 初始化: fx {self.fx:.2f}, 基线 {self.baseline:.4f}m, "
                  f"主点 ({self.cx:.2f}, {self.cy:.2f})")

    def _params(self, kind, min_disparity, window=3):
        p = N.ReconParams()
        p.kind = kind
        p.fx, p.baseline, p.cx, p.cy = float(self.fx), float(self.baseline), float(self.cx), float(self.cy)
        p.min_disparity = float(min_disparity)
        p.window = int(window)
        p.n_water = 1.33
        return p

    def _run(self, kind, laser_points, disparity_map, min_disparity, window=3):
        if len(laser_points) == 0:
            return np.array([])
        disp = np.asarray(disparity_map)
        if disp.ndim != 2:
            raise ValueError("disparity_map must be HxW")
        xy = np.asarray(laser_points, np.float64).reshape(-1, 2)
        out = N.default_context(self.device).reconstruct(self._params(kind, min_disparity, window), xy,
                                                         disp.astype(np.float32, copy=False))
        return out.astype(np.float32) if len(out) else np.array([])

    def reconstruct_from_disparity(self, laser_points: List[Tuple[float, float]], disparity_map: np.ndarray,
                                   min_disparity: float = 1.0) -> np.ndarray:
        """:37-86."""
        return self._run(N.RECON_DISPARITY, laser_points, disparity_map, min_disparity)

    def reconstruct_with_interpolation(self, laser_points: List[Tuple[float, float]], disparity_map: np.ndarray,
                                       window_size: int = 3, min_disparity: float = 1.0) -> np.ndarray:
        """:88-152: median of the valid disparities in the window around each laser pixel."""
        if window_size < 1 or window_size > 9 or not (window_size & 1):
            raise ValueError("window_size must be odd and <= 9")
        return self._run(N.RECON_DISPARITY_MEDIAN, laser_points, disparity_map, min_disparity, window_size)

    def create_laser_depth_map(self, laser_points: List[Tuple[float, float]], disparity_map: np.ndarray,
                               image_shape: Tuple[int, int]) -> np.ndarray:
        """:154-186: depth only at the (rounded) laser pixels, `disparity > 1.0`.  Built from the
        same GPU reconstruction (Z of reconstruct_from_disparity with min_disparity just above 1)."""
        h, w = image_shape
        out = np.zeros((h, w), np.float32)
        if len(laser_points) == 0:
            return out
        disp = np.asarray(disparity_map, np.float32)
        xy = np.asarray(laser_points, np.float64).reshape(-1, 2)
        px = np.rint(xy[:, 0]).astype(np.int64)
        py = np.rint(xy[:, 1]).astype(np.int64)
        ok = (px >= 0) & (px < w) & (py >= 0) & (py < h)
        if not ok.any():
            return out
        # points that survive the reference's tests, in order; Z from the device kernel
        md = float(np.nextafter(np.float32(1.0), np.float32(2.0)))  # "> 1.0" on float32 values
        pts = self._run(N.RECON_DISPARITY, [tuple(p) for p in xy[ok]], disp[:h, :w], md)
        d = disp[py[ok].clip(0, disp.shape[0] - 1), px[ok].clip(0, disp.shape[1] - 1)]
        with np.errstate(all="ignore"):
            z = (float(self.fx) * float(self.baseline)) / d.astype(np.float64)
        keep = (d > 1.0) & ~np.isnan(d) & ~np.isinf(d) & (z > 0) & (z < 10.0)
        if len(pts) == int(keep.sum()):
            out[py[ok][keep], px[ok][keep]] = pts[:, 2]
        else:  # shapes disagree only if disparity_map is smaller than image_shape: follow the reference's bounds
            raise ValueError("disparity_map smaller than image_shape")
        return out


def fix_roi_alignment(left_rect, right_rect, roi_left, roi_right):
    """:189-227 -- pure slicing (host)."""
    x1, y1 = max(roi_left[0], roi_right[0]), max(roi_left[1], roi_right[1])
    x2 = min(roi_left[0] + roi_left[2], roi_right[0] + roi_right[2])
    y2 = min(roi_left[1] + roi_left[3], roi_right[1] + roi_right[3])
    if x2 - x1 > 0 and y2 - y1 > 0:
        return left_rect[y1:y2, x1:x2], right_rect[y1:y2, x1:x2]
    return left_rect, right_rect
