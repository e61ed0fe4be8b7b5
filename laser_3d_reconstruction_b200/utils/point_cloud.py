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
