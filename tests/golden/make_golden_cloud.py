"""Golden vectors for the point-cloud sink, produced by the REAL reference class
(/root/reference/utils/point_cloud.py, which runs its Open3D-free code here because Open3D is not installed).

    python tests/golden/make_golden_cloud.py      # writes tests/golden/cloud.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from utils.point_cloud import PointCloudProcessor  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(7)
# a laser-scan-like cloud: a wavy sheet, 2 mm voxels, plus exact duplicates and a regular lattice (equal neighbour distances)
u, v = rng.uniform(-0.05, 0.05, 3000), rng.uniform(-0.03, 0.03, 3000)
sheet = np.stack([u, v, 0.4 + 0.02 * np.sin(40 * u) + rng.normal(0, 0.0003, 3000)], 1)
dups = np.repeat(sheet[:30], 25, axis=0)
g = np.arange(6) * 0.004
lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + np.array([0.2, 0.2, 0.2])
cloud = np.concatenate([sheet, dups, lattice])
cloud = cloud[rng.permutation(len(cloud))]
p = PointCloudProcessor()
assert not p.try_open3d
c32 = cloud.astype(np.float32)  # main.py:208 hands the accumulated cloud over as float32
v32 = p.voxel_downsample(c32, 0.002)
np.savez_compressed(os.path.join(HERE, "cloud.npz"), cloud=cloud, voxel_2mm_f32=v32, sor_20_2_f32=p.statistical_outlier_removal(v32, 20, 2.0),
                    voxel_2mm=p.voxel_downsample(cloud, 0.002), voxel_5mm=p.voxel_downsample(cloud, 0.005),
                    sor_20_2=p.statistical_outlier_removal(cloud, 20, 2.0), sor_8_0=p.statistical_outlier_removal(cloud, 8, 0.0),
                    sor_20_tiny=p.statistical_outlier_removal(cloud, 20, 1e-20))
print({k: v.shape for k, v in np.load(os.path.join(HERE, "cloud.npz")).items()})
