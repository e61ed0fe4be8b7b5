from .point_cloud import PointCloudProcessor  # noqa: F401
