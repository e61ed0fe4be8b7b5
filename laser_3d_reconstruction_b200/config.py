"""Config constants the hot path consumes (reference config.py:9-99; dead keys kept for drop-in
compatibility -- SURVEY fact 9)."""
import numpy as np


class Config:
    SINGLE_USB_CAMERA_ID = 1
    CAMERA_WIDTH = 640
    CAMERA_HEIGHT = 240
    CAMERA_FPS = 30
    SPLIT_MODE = 'horizontal'
    STEREO_CALIBRATION_FILE = "stereo_calibration.json"
    STEREO_NUM_DISPARITIES = 64
    STEREO_BLOCK_SIZE = 5
    STEREO_USE_WLS_FILTER = True
    STEREO_LAMBDA = 8000.0
    STEREO_SIGMA = 1.5
    LASER_EXTRACTOR_TYPE = 'simple'
    SIMPLE_LASER_HSV_LOWER = np.array([50, 100, 180])
    SIMPLE_LASER_HSV_UPPER = np.array([70, 255, 255])
    SIMPLE_LASER_BRIGHTNESS_THRESHOLD = 200
    SIMPLE_LASER_MIN_AREA = 50
    STEGER_SIGMA = 3.0
    STEGER_BRIGHTNESS_THRESHOLD = 200
    STEGER_USE_LUT = True
    LASER_PLANE_COEFFICIENTS = np.array([0, 0, 1, 0], dtype=np.float64)
    USE_REFRACTION_CORRECTION = False
    WATER_REFRACTION_INDEX = 1.33
    VOXEL_SIZE = 0.002
    OUTLIER_REMOVAL_NEIGHBORS = 20
    OUTLIER_REMOVAL_STD_RATIO = 2.0
    SAVE_FORMAT = 'ply'
    DEBUG_MODE = False
    OUTPUT_DIR = 'output'
    AUTO_SAVE_INTERVAL = 60
    MIN_POINT_CLOUD_SIZE = 100
    JETSON_OPTIMIZED = False
    USE_CUDA = True  # this build is CUDA-only
    NUM_THREADS = 4

    @classmethod
    def to_dict(cls):
        out = {}
        for key in dir(cls):
            if not key.startswith('_') and key.isupper():
                v = getattr(cls, key)
                out[key] = v.tolist() if isinstance(v, np.ndarray) else v
        return out

    @classmethod
    def print_config(cls):
        """reference config.py:114-147: the settings a run uses, by group, on stdout."""
        rule = "=" * 60
        groups = (
            ("camera", (("id", cls.SINGLE_USB_CAMERA_ID), ("resolution", f"{cls.CAMERA_WIDTH}x{cls.CAMERA_HEIGHT}"),
                        ("fps", cls.CAMERA_FPS), ("split mode", cls.SPLIT_MODE), ("calibration", cls.STEREO_CALIBRATION_FILE))),
            ("stereo matching", (("disparities", cls.STEREO_NUM_DISPARITIES), ("block size", cls.STEREO_BLOCK_SIZE),
                                 ("WLS filter", "on" if cls.STEREO_USE_WLS_FILTER else "off"))),
            ("laser extraction", (("extractor", cls.LASER_EXTRACTOR_TYPE),) + (
                (("HSV range", f"{cls.SIMPLE_LASER_HSV_LOWER} ~ {cls.SIMPLE_LASER_HSV_UPPER}"),
                 ("brightness threshold", cls.SIMPLE_LASER_BRIGHTNESS_THRESHOLD)) if cls.LASER_EXTRACTOR_TYPE == 'simple' else
                (("sigma", cls.STEGER_SIGMA), ("brightness threshold", cls.STEGER_BRIGHTNESS_THRESHOLD)))),
            ("reconstruction", (("refraction correction", "on" if cls.USE_REFRACTION_CORRECTION else "off"),
                                ("voxel size", cls.VOXEL_SIZE), ("output dir", cls.OUTPUT_DIR))),
        )
        print("\n" + rule + "\nlaser3d-b200 configuration (single USB stereo camera)\n" + rule)
        for title, rows in groups:
            print(f"\n{title}:")
            for name, value in rows:
                print(f"  {name}: {value}")
        print(rule + "\n")
