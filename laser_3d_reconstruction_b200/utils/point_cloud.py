"""PointCloudProcessor with the reference's API (utils/point_cloud.py:12-181).

The reference uses Open3D when it is installed and otherwise its own numpy / scipy code; this class follows the
Open3D-free code (`_simple_voxel_downsample` :54-78, `_simple_outlier_removal` :108-131) bit for bit, on the GPU
(csrc/pointcloud.cu).  The PLY writer is host I/O and produces the same ASCII file (:134-181).  No CPU fallback.
"""
import os
from typing import Optional

import numpy as np

from .. import _native as N


class PointCloudProcessor:
    def __init__(self, device: int = 0, verbose: bool = True):
        self.try_open3d = False  # the reference's attribute: Open3D is never used here
        self.device = device
        if verbose:
            print("点云处理器 (B200): voxel_downsample / statistical_outlier_removal on the GPU")

    def voxel_downsample(self, points: np.ndarray, voxel_size: float = 0.002) -> np.ndarray:
        if len(points) == 0:
            return points
        return N.default_context(self.device).voxel_downsample(points, voxel_size)

    def statistical_outlier_removal(self, points: np.ndarray, nb_neighbors: int = 20, std_ratio: float = 2.0) -> np.ndarray:
        if len(points) < nb_neighbors:
            return points
        out = N.default_context(self.device).statistical_outlier_removal(points, nb_neighbors, std_ratio)
        return out if len(out) else np.array([])  # np.array([]) of an empty list, like the reference

    def save_ply(self, points: np.ndarray, filename: str, colors: Optional[np.ndarray] = None):
        if len(points) == 0:
            print("空点云，无法保存")
            return
        if not filename.endswith('.ply'):
            filename += '.ply'
        os.makedirs(os.path.dirname(filename) if os.path.dirname(filename) else '.', exist_ok=True)
        with_colors = colors is not None and len(colors) == len(points)
        with open(filename, 'w') as f:
            f.write("ply\nformat ascii 1.0\n")
            f.write(f"element vertex {len(points)}\n")
            f.write("property float x\nproperty float y\nproperty float z\n")
            if with_colors:
                f.write("property uchar red\nproperty uchar green\nproperty uchar blue\n")
            f.write("end_header\n")
            for i in range(len(points)):
                f.write(f"{points[i, 0]:.6f} {points[i, 1]:.6f} {points[i, 2]:.6f}")
                if with_colors:
                    c = colors[i]
                    f.write(f" {int(c[0])} {int(c[1])} {int(c[2])}")
                f.write("\n")
        print(f"点云已保存到: {filename} ({len(points)} 点)")

    def save_pcd(self, points: np.ndarray, filename: str, colors: Optional[np.ndarray] = None):
        """The reference needs Open3D for PCD and otherwise writes PLY (:183-196); so does this."""
        print("需要Open3D才能保存PCD格式，改为保存PLY格式")
        self.save_ply(points, filename.replace('.pcd', '.ply'), colors)

    # ---- the rest of the reference's public surface: host-side helpers around the two GPU calls ------------------
    def estimate_normals(self, points: np.ndarray, radius: float = 0.01) -> np.ndarray:
        """:216-237: an Open3D-only feature; without Open3D the reference says so and returns zeros."""
        print("需要Open3D来计算法向量")
        return np.zeros_like(points)

    def compute_point_cloud_metrics(self, points: np.ndarray) -> dict:
        """:239-278: count, centre, spread, bounding box, volume and the mean nearest-neighbour spacing (host numpy: a
        report, not part of the per-frame path)."""
        if len(points) == 0:
            return {'num_points': 0, 'bbox': None, 'center': None, 'dimensions': None}
        lo, hi = np.min(points, axis=0), np.max(points, axis=0)
        metrics = {'num_points': len(points), 'center': np.mean(points, axis=0), 'std': np.std(points, axis=0),
                   'min': lo, 'max': hi}
        metrics['dimensions'] = hi - lo
        metrics['volume'] = np.prod(metrics['dimensions'])
        if len(points) > 1:
            metrics['avg_point_spacing'] = _mean_nearest_neighbour_distance(np.asarray(points))
        return metrics

    def visualize(self, points: np.ndarray, colors: Optional[np.ndarray] = None, window_name: str = "3D Point Cloud"):
        """:280-321: an Open3D viewer; without Open3D the reference prints the point count and returns."""
        print("需要Open3D进行可视化")
        print(f"点云包含 {len(points)} 个点")

    def merge_and_clean(self, point_clouds: list, voxel_size: float = 0.002) -> np.ndarray:
        """:323-349: stack the non-empty clouds, voxel down-sample, remove statistical outliers."""
        if not point_clouds:
            return np.array([])
        merged = np.vstack([pc for pc in point_clouds if len(pc) > 0])
        if len(merged) == 0:
            return merged
        return self.statistical_outlier_removal(self.voxel_downsample(merged, voxel_size))


def _mean_nearest_neighbour_distance(points: np.ndarray) -> float:
    """mean over the points of the distance to their nearest other point (the reference asks scipy's cKDTree for k = 2)"""
    try:
        from scipy.spatial import cKDTree
        d, _ = cKDTree(points).query(points, k=2)
        return np.mean(d[:, 1])
    except ImportError:
        p = np.asarray(points, np.float64)
        best = np.empty(len(p))
        for i0 in range(0, len(p), 1024):   # blocked brute force
            d2 = ((p[i0:i0 + 1024, None, :] - p[None, :, :]) ** 2).sum(-1)
            d2[np.arange(d2.shape[0]), np.arange(i0, i0 + d2.shape[0])] = np.inf
            best[i0:i0 + 1024] = np.sqrt(d2.min(axis=1))
        return np.mean(best)
