"""SingleUSBStereoCameraManager with the reference's API (camera/single_usb_stereo_camera.py) and
the per-frame depth path (:311-359: remap -> gray -> SGBM left/right -> WLS -> depth) running in the
sm_100a kernels of libl3d.so.

What stays in cv2 because it is hardware I/O: the UVC capture (``cv2.VideoCapture``) and the MJPG decode behind it.  The
init-time calibration maths (:162-206) needs no cv2: ``cv2.stereoRectify`` is restated in ``rectify.py`` (host,
float64) and ``cv2.initUndistortRectifyMap`` is a kernel of the library with ``gpu_maps=True`` (the default keeps
cv2's function for the maps; both give the same f32 maps).  The side-by-side split (:143-150) costs no copy: the two
views go to the library as strided rows.  There is no CPU fallback for the per-frame work: a missing library or GPU
raises.

New, keyword-only: ``num_disparities``, ``block_size``, ``sgbm_mode``, ``use_wls``, ``device``,
``verbose`` (defaults reproduce the reference: 64/5 below 400 px per eye else 96/7, MODE_SGBM_3WAY,
WLS on) and ``compute_depth(left, right)`` -- the README batch entry point (readme.md:357-374).
"""
import json
import os
from typing import Dict, Optional, Tuple

import cv2
import numpy as np

from .. import _native as N
from .. import stereo
from .rectify import CALIB_ZERO_DISPARITY, stereo_rectify


class SingleUSBStereoCameraManager:
    def __init__(self, camera_id: int = 0, width: int = 640, height: int = 240, fps: int = 30,
                 split_mode: str = 'horizontal', calibration_file: str = 'stereo_calibration.json', *,
                 num_disparities: Optional[int] = None, block_size: Optional[int] = None,
                 sgbm_mode: Optional[int] = None, use_wls: bool = True, device: int = 0, verbose: bool = True,
                 gpu_maps: bool = False):
        self.camera_id = camera_id
        self.width = width
        self.height = height
        self.fps = fps
        self.split_mode = split_mode
        self.calibration_file = calibration_file
        self.cap = None
        self.camera_matrix_left = None
        self.dist_coeffs_left = None
        self.camera_matrix_right = None
        self.dist_coeffs_right = None
        self.R = None
        self.T = None
        self.R1 = self.R2 = self.P1 = self.P2 = self.Q = None
        self.roi_left = None   # never assigned by the reference either
        self.roi_right = None
        self.map_left_x = self.map_left_y = self.map_right_x = self.map_right_y = None
        self.stereo_matcher = None
        self.right_matcher = None
        self.wls_filter = None
        if split_mode == 'horizontal':
            self.single_width, self.single_height = width // 2, height
        else:
            self.single_width, self.single_height = width, height // 2
        self._num_disparities = num_disparities
        self._block_size = block_size
        self._sgbm_mode = sgbm_mode
        self._use_wls = bool(use_wls)
        self._gpu_maps = bool(gpu_maps)  # rectification maps by l3d_init_undistort_rectify_map instead of cv2 (same bits)
        self.device = device
        self.verbose = verbose
        self._ctx = None
        self._maps_key = None
        self._say("单USB双目相机管理器配置 (B200):")
        self._say(f"  相机ID: {camera_id}  总分辨率: {width}x{height}  单目分辨率: {self.single_width}x{self.single_height}"
                  f"  分割模式: {split_mode}  帧率: {fps} FPS")

    def _say(self, *a):
        if self.verbose:
            print(*a)

    # ---- initialisation (reference :84-141) -----------------------------------------------
    def initialize(self) -> bool:
        self.cap = cv2.VideoCapture(self.camera_id)
        if not self.cap.isOpened():
            print(f"❌ 无法打开相机 {self.camera_id}")
            return False
        self.cap.set(cv2.CAP_PROP_FRAME_WIDTH, self.width)
        self.cap.set(cv2.CAP_PROP_FRAME_HEIGHT, self.height)
        self.cap.set(cv2.CAP_PROP_FPS, self.fps)
        ret, frame = self.cap.read()
        if not ret or frame is None:
            print("❌ 无法读取图像")
            return False
        self._split_frame(frame)
        if not self._load_calibration():
            self._say("⚠️  标定文件加载失败，使用默认参数")
            self._use_default_calibration()
        self._initialize_stereo_matcher_optimized()
        for _ in range(10):
            self.cap.read()
        return True

    def initialize_offline(self) -> bool:
        """initialize() without the capture device: calibration + matcher only (batch use,
        readme.md:357-374)."""
        if not self._load_calibration():
            self._use_default_calibration()
        self._initialize_stereo_matcher_optimized()
        return True

    def _split_frame(self, frame: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """reference :143-150 -- non-contiguous views, like the reference."""
        if self.split_mode == 'horizontal':
            mid = frame.shape[1] // 2
            return frame[:, :mid], frame[:, mid:]
        mid = frame.shape[0] // 2
        return frame[:mid, :], frame[mid:, :]

    def _load_calibration(self) -> bool:
        """reference :152-213: calibration file -> rectification matrices (rectify.py) -> maps (GPU kernel with gpu_maps)."""
        if not os.path.exists(self.calibration_file):
            self._say(f"  标定文件不存在: {self.calibration_file}")
            return False
        try:
            with open(self.calibration_file, 'r', encoding='utf-8') as f:
                calib = json.load(f)
            self.camera_matrix_left = np.array(calib['camera_matrix_left'])
            self.dist_coeffs_left = np.array(calib['dist_coeffs_left'])
            self.camera_matrix_right = np.array(calib['camera_matrix_right'])
            self.dist_coeffs_right = np.array(calib['dist_coeffs_right'])
            self.R = np.array(calib['R'])
            self.T = np.array(calib['T'])
            size = (self.single_width, self.single_height)
            # cv2.stereoRectify(..., flags=cv2.CALIB_ZERO_DISPARITY, alpha=0) restated in rectify.py (P1, P2, Q bit-identical to
            # cv2 4.x, R1 / R2 to the last bit or two, the maps built from them identical): no cv2 call at init time
            self.R1, self.R2, self.P1, self.P2, self.Q, _roi_l, _roi_r = stereo_rectify(
                self.camera_matrix_left, self.dist_coeffs_left, self.camera_matrix_right, self.dist_coeffs_right,
                size, self.R, self.T, flags=CALIB_ZERO_DISPARITY, alpha=0)
            if self._gpu_maps:
                ctx = self._context()
                self.map_left_x, self.map_left_y = ctx.init_undistort_rectify_map(
                    self.camera_matrix_left, self.dist_coeffs_left, self.R1, self.P1, size)
                self.map_right_x, self.map_right_y = ctx.init_undistort_rectify_map(
                    self.camera_matrix_right, self.dist_coeffs_right, self.R2, self.P2, size)
            else:
                self.map_left_x, self.map_left_y = cv2.initUndistortRectifyMap(
                    self.camera_matrix_left, self.dist_coeffs_left, self.R1, self.P1, size, cv2.CV_32FC1)
                self.map_right_x, self.map_right_y = cv2.initUndistortRectifyMap(
                    self.camera_matrix_right, self.dist_coeffs_right, self.R2, self.P2, size, cv2.CV_32FC1)
            self._say(f"✓ 从 {self.calibration_file} 加载标定参数, 基线距离: {np.linalg.norm(self.T):.3f}m")
            return True
        except Exception as e:  # the reference swallows and reports, :211-213
            print(f"  加载标定参数失败: {e}")
            return False

    def _use_default_calibration(self):
        """reference :215-231."""
        fx = fy = 350.0
        cx, cy = self.single_width / 2.0, self.single_height / 2.0
        self.camera_matrix_left = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
        self.camera_matrix_right = self.camera_matrix_left.copy()
        self.dist_coeffs_left = np.zeros((1, 5))
        self.dist_coeffs_right = np.zeros((1, 5))

    def _initialize_stereo_matcher_optimized(self):
        """reference :233-292: matcher parameters, right matcher, WLS(8000, 1.5).  The objects are
        libl3d-backed look-alikes of the cv2 ones (stereo.py); creating the WLS filter mutates the
        left matcher exactly as cv2.ximgproc does."""
        if self.single_width < 400:
            nd, bs = 64, 5
        else:
            nd, bs = 96, 7
        if self._num_disparities is not None:
            nd = int(self._num_disparities)
        if self._block_size is not None:
            bs = int(self._block_size)
        mode = stereo.STEREO_SGBM_MODE_SGBM_3WAY if self._sgbm_mode is None else int(self._sgbm_mode)
        self.stereo_matcher = stereo.StereoSGBM(
            minDisparity=0, numDisparities=nd, blockSize=bs, P1=8 * 3 * bs ** 2, P2=32 * 3 * bs ** 2,
            disp12MaxDiff=1, uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, preFilterCap=63,
            mode=mode, device=self.device)
        if self._use_wls:
            self.right_matcher = stereo.createRightMatcher(self.stereo_matcher)
            self.wls_filter = stereo.createDisparityWLSFilter(self.stereo_matcher)
            self.wls_filter.setLambda(8000.0)
            self.wls_filter.setSigmaColor(1.5)
        self._say(f"  立体匹配参数: 视差范围 0-{nd}, 匹配块大小 {bs}, WLS {'是' if self._use_wls else '否'}, mode {mode}")

    # ---- the hot path -----------------------------------------------------------------------
    def _context(self):
        if self._ctx is None:
            self._ctx = N.Context(self.device)  # raises without library / GPU: no CPU fallback
        return self._ctx

    def _depth_config(self):
        if self.stereo_matcher is None:
            raise RuntimeError("stereo matcher not initialised: call initialize() or initialize_offline()")
        cfg = N.DepthConfig()
        cfg.left = self.stereo_matcher.params()
        use_wls = self.wls_filter is not None and self.right_matcher is not None
        if use_wls:
            cfg.right = self.right_matcher.params()
            cfg.wls = self.wls_filter.params()
        cfg.use_wls = int(use_wls)
        cfg.use_maps = int(self.map_left_x is not None)
        cfg.use_Q = int(self.Q is not None)
        if self.Q is not None:
            cfg.Q[:] = [float(v) for v in np.asarray(self.Q, np.float64).reshape(16)]
        return cfg

    def _upload_maps(self, ctx):
        key = tuple(id(m) for m in (self.map_left_x, self.map_left_y, self.map_right_x, self.map_right_y))
        if key != self._maps_key:
            ctx.set_rectify_maps(0, self.map_left_x, self.map_left_y)
            ctx.set_rectify_maps(1, self.map_right_x, self.map_right_y)
            self._maps_key = key

    def rectify_and_depth(self, left_img: np.ndarray, right_img: np.ndarray):
        """reference :311-359 for an already captured/split pair -> (left_rectified, depth_map)."""
        ctx = self._context()
        cfg = self._depth_config()
        if cfg.use_maps:
            self._upload_maps(ctx)
        rect, depth = ctx.compute_depth(cfg, left_img, right_img)
        return rect, depth

    def compute_depth(self, left_img: np.ndarray, right_img: np.ndarray) -> np.ndarray:
        """readme.md:367 -- depth map (f32 HxW, metres, 0 = invalid) of one stereo pair."""
        return self.rectify_and_depth(left_img, right_img)[1]

    def get_frames(self) -> Tuple[Optional[np.ndarray], Optional[np.ndarray]]:
        """reference :294-359."""
        if self.cap is None or not self.cap.isOpened():
            return None, None
        ret, frame = self.cap.read()
        if not ret:
            return None, None
        left_img, right_img = self._split_frame(frame)
        return self.rectify_and_depth(left_img, right_img)

    def get_camera_intrinsics(self) -> Optional[Dict]:
        """reference :361-382."""
        if self.camera_matrix_left is None:
            return None
        K = self.P1[:3, :3] if self.P1 is not None else self.camera_matrix_left
        baseline = np.linalg.norm(self.T) if self.T is not None else 0.06
        return {'width': self.single_width, 'height': self.single_height, 'fx': float(K[0, 0]), 'fy': float(K[1, 1]),
                'cx': float(K[0, 2]), 'cy': float(K[1, 2]), 'baseline': float(baseline)}

    def stop(self):
        """reference :384-388."""
        if self.cap is not None:
            self.cap.release()
            self._say("✓ 相机已释放")
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None
            self._maps_key = None


def visualize_depth(depth_map: np.ndarray, max_depth: float = 2.0, colormap=cv2.COLORMAP_JET) -> np.ndarray:
    """reference :392-416 (display helper, host-side)."""
    norm = np.clip(depth_map / max_depth * 255, 0, 255).astype(np.uint8)
    col = cv2.applyColorMap(norm, colormap)
    col[depth_map <= 0] = [0, 0, 0]
    return col
