"""Camera module (reference camera/__init__.py:6)."""
from .single_usb_stereo_camera import SingleUSBStereoCameraManager, visualize_depth

__all__ = ["SingleUSBStereoCameraManager", "visualize_depth"]
