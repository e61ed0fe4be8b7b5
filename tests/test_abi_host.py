"""CPU-side checks: the C-ABI library loads and exports every symbol include/l3d.h declares (no
compute calls), ctypes struct layouts match the header, the host logic of the reference-API classes,
and the loud failure without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from laser_3d_reconstruction_b200 import _native as N
from laser_3d_reconstruction_b200 import pipeline, sharding, stereo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "l3d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(l3d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = N.load()
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), "libl3d.so does not export %s" % s
    assert set(N.EXPORTS) == set(syms)
    assert b"sm_100a" in lib.l3d_version()


def test_struct_layouts_match_header():
    # sizes as the C compiler lays them out (ints 4, doubles 8, natural alignment)
    assert C.sizeof(N.SgbmParams) == 11 * 4
    assert C.sizeof(N.WlsParams) == 2 * 8 + 6 * 4  # + solver, variant
    assert C.sizeof(N.DepthConfig) == 44 + 44 + 40 + 12 + 4 + 16 * 8
    assert C.sizeof(N.StegerParams) == 8 + 8 + 8 + 8 + 16 + 12 + 12
    assert N.ReconParams.K.offset == 8 and N.ReconParams.window.offset == 8 + 72 + 32 + 8 + 8 + 5 * 8


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(N.L3DError, match="no CPU fallback|no CUDA"):
        N.Context(0)
    from laser_3d_reconstruction_b200 import SimpleLaserExtractor
    with pytest.raises(N.L3DError):
        SimpleLaserExtractor(verbose=False).extract_centerline(np.zeros((8, 8, 3), np.uint8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "laser_3d_reconstruction_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "liborc" not in txt, f
    # ... and nothing under tools/ (measurement helpers) does either: whatever checks against the oracle lives in tests/
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith((".py", ".sh")):
            txt = open(os.path.join(ROOT, "tools", f), encoding="utf-8").read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M) and "liborc" not in txt, f


def test_wls_creation_mutates_left_matcher():
    m = stereo.StereoSGBM(minDisparity=0, numDisparities=64, blockSize=5, P1=600, P2=2400, disp12MaxDiff=1,
                          uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, preFilterCap=63, mode=2)
    r = stereo.createRightMatcher(m)
    assert (r.getMinDisparity(), r.getNumDisparities(), r.getUniquenessRatio(), r.getDisp12MaxDiff(),
            r.getSpeckleWindowSize()) == (-63, 64, 0, 1000000, 0)
    w = stereo.createDisparityWLSFilter(m)
    assert (m.getUniquenessRatio(), m.getDisp12MaxDiff(), m.getSpeckleWindowSize()) == (0, 1000000, 0)
    assert (w.getLambda(), w.getSigmaColor(), w.getLRCthresh(), w.getDepthDiscontinuityRadius()) == (8000.0, 1.0, 24, 3)
    # pipeline.depth_config builds the same parameter sets
    cfg = pipeline.depth_config(64, 5, 2, np.eye(4))
    for f, _ in N.SgbmParams._fields_:
        assert getattr(cfg.left, f) == getattr(m.params(), f), f
        if f != "speckleRange":
            assert getattr(cfg.right, f) == getattr(r.params(), f), f


def test_camera_class_host_side(golden_real, tmp_path):
    """Calibration loading / matcher selection of the camera class without touching the GPU."""
    import json
    from laser_3d_reconstruction_b200 import SingleUSBStereoCameraManager
    g = golden_real
    calib = {k: g[k].tolist() for k in ("R", "T")}
    calib.update(camera_matrix_left=g["K_left"].tolist(), dist_coeffs_left=g["dist_left"].tolist(),
                 camera_matrix_right=g["K_right"].tolist(), dist_coeffs_right=g["dist_right"].tolist())
    path = tmp_path / "stereo_calibration.json"
    path.write_text(json.dumps(calib))
    cam = SingleUSBStereoCameraManager(camera_id=0, width=640, height=240, calibration_file=str(path), verbose=False)
    assert (cam.single_width, cam.single_height) == (320, 240)
    assert cam.initialize_offline()
    assert np.array_equal(cam.Q, g["Q"])
    assert np.array_equal(cam.map_left_x[::8, ::8], g["map_left_x"])
    assert cam.roi_left is None and cam.roi_right is None
    intr = cam.get_camera_intrinsics()
    assert np.allclose([intr[k] for k in ("fx", "fy", "cx", "cy", "baseline")], g["intr"])
    assert cam.stereo_matcher.getNumDisparities() == 64 and cam.stereo_matcher.getBlockSize() == 5
    assert cam.stereo_matcher.getMode() == stereo.STEREO_SGBM_MODE_SGBM_3WAY
    assert cam.stereo_matcher.getUniquenessRatio() == 0  # mutated by the WLS filter creation
    assert cam.right_matcher.getMinDisparity() == -63
    l, r = cam._split_frame(g["frame_a"])
    assert l.shape == (240, 320, 3) and not r.flags["C_CONTIGUOUS"]
    assert cam.get_frames() == (None, None)  # no capture device opened
    big = SingleUSBStereoCameraManager(width=2560, height=720, calibration_file="/nonexistent", num_disparities=128,
                                       block_size=9, sgbm_mode=1, verbose=False)
    big.initialize_offline()
    assert big.Q is None and big.map_left_x is None
    assert (big.stereo_matcher.getNumDisparities(), big.stereo_matcher.getBlockSize(), big.stereo_matcher.getMode(),
            big.stereo_matcher.getP2()) == (128, 9, 1, 7776)
    assert SingleUSBStereoCameraManager(width=1280, height=720, verbose=False, calibration_file="/x").single_width == 640


def test_shard_and_pack_roundtrip():
    for world in (1, 2, 3, 8):
        seen = sorted(sum((sharding.shard_frames(37, r, world) for r in range(world)), []))
        assert seen == list(range(37))
    with pytest.raises(ValueError):
        sharding.shard_frames(4, 2, 2)
    rng = np.random.default_rng(0)
    clouds = [rng.random((n, 3)) for n in (0, 5, 1, 0, 7)]
    tabs = [sharding.pack_clouds(sharding.shard_frames(5, r, 2), [clouds[f] for f in sharding.shard_frames(5, r, 2)])
            for r in range(2)]
    back = sharding.unpack_clouds(np.concatenate(tabs), 5)
    assert all(np.array_equal(a, b) for a, b in zip(back, clouds))


def test_point_list_conversions():
    """The point lists the reference's API carries ([(x, y), ...]) <-> arrays: the fast conversions give what the plain
    numpy ones give, for every input form a caller may hand over."""
    from laser_3d_reconstruction_b200 import _native as N
    rng = np.random.default_rng(3)
    a32 = rng.random((500, 2), np.float32) * 1000
    a64 = a32.astype(np.float64)
    lst = N.points_to_list(a32)
    assert lst == list(map(tuple, a64.tolist())) and type(lst[0]) is tuple and type(lst[0][0]) is float
    assert N.points_to_list(np.empty((0, 2), np.float32)) == []
    l32 = N.points_to_list_f32(a32)
    assert l32 == [(x, y) for x, y in a32] and type(l32[0]) is tuple and type(l32[0][0]) is np.float32
    assert N.points_to_list_f32(np.empty((0, 2), np.float32)) == [] and N.points_to_list_f32(a32[:1]) == [(a32[0, 0], a32[0, 1])]
    forms = [lst, [list(p) for p in lst], tuple(lst), [(x, y) for x, y in a32], a32, a64, a64.tolist(), a64[::2]]
    for f in forms:
        got = N.points_to_array(f)
        want = np.asarray(f, np.float64).reshape(-1, 2)
        assert got.dtype == np.float64 and got.shape == want.shape and np.array_equal(got, want)
    assert N.points_to_array([]).shape == (0, 2)
    assert np.array_equal(N.points_to_array([(1, 2)]), [[1.0, 2.0]])
    mixed = [(1.0, 2.0), [3, 4.5], (np.float32(5.5), 6)]
    assert np.array_equal(N.points_to_array(mixed), np.asarray(mixed, np.float64))
    with pytest.raises((TypeError, ValueError)):
        N.points_to_array([(1.0, 2.0), (3.0,), (4.0, 5.0)])   # ragged middle entry: numpy's own error, nothing silent
    with pytest.raises((TypeError, ValueError)):
        N.points_to_array([(1.0, 2.0), (3.0, 4.0, 5.0), (6.0,), (7.0, 8.0)])   # 2n values in all, but not n pairs
    with pytest.raises((TypeError, ValueError)):
        N.points_to_array([(1.0, "a"), (2.0, 3.0)])


REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree exists in the build container only")
def test_public_api_surface_matches_the_reference():
    """Every public method of the reference's hot-path classes exists here with the same leading parameter names, and the
    host-side helpers (Config, point-cloud metrics, visualize_laser_depth) give the reference's own results."""
    import contextlib
    import importlib
    import inspect
    import io
    import sys
    import laser_3d_reconstruction_b200 as l3d
    from laser_3d_reconstruction_b200.camera import SingleUSBStereoCameraManager
    from laser_3d_reconstruction_b200.utils import PointCloudProcessor
    sys.path.insert(0, REFERENCE)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("config", "core", "camera", "utils")}
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            pairs = [("core.laser_extractor", "SimpleLaserExtractor", l3d.SimpleLaserExtractor),
                     ("core.laser_extractor", "FastStegerExtractor", l3d.FastStegerExtractor),
                     ("core.reconstruction", "Reconstructor", l3d.Reconstructor),
                     ("improved_steger", "ImprovedStegerExtractor", l3d.ImprovedStegerExtractor),
                     ("improved_steger", "HybridLaserExtractor", l3d.HybridLaserExtractor),
                     ("improved_reconstruction", "ImprovedLaserReconstructor", l3d.ImprovedLaserReconstructor),
                     ("camera.single_usb_stereo_camera", "SingleUSBStereoCameraManager", SingleUSBStereoCameraManager),
                     ("utils.point_cloud", "PointCloudProcessor", PointCloudProcessor),
                     ("config", "Config", l3d.Config), ("main", "LaserReconstructionSystem", l3d.LaserReconstructionSystem)]
            for mod, name, mine in pairs:
                ref = getattr(importlib.import_module(mod), name)
                for k, v in inspect.getmembers(ref, predicate=lambda f: inspect.isfunction(f) or inspect.ismethod(f)):
                    if k.startswith("_") and k != "__init__":
                        continue
                    assert hasattr(mine, k), "%s.%s missing" % (name, k)
                    want = list(inspect.signature(v).parameters)
                    got = list(inspect.signature(getattr(mine, k)).parameters)
                    assert got[:len(want)] == want, (name, k, want, got)
            ref_ir = importlib.import_module("improved_reconstruction")
            for fn in ("fix_roi_alignment", "visualize_laser_depth"):
                assert hasattr(ref_ir, fn) and hasattr(l3d, fn)
            # Config: same constants (USE_CUDA is the one deliberate difference)
            RefConfig = importlib.import_module("config").Config
            for k in dir(RefConfig):
                if k.isupper() and k != "USE_CUDA":
                    assert np.array_equal(getattr(RefConfig, k), getattr(l3d.Config, k)), k
            # point-cloud metrics and the depth visualisation against the reference's own code
            rng = np.random.default_rng(5)
            cloud = rng.normal(size=(400, 3)) * [0.2, 0.1, 0.05] + [0, 0, 1]
            want = importlib.import_module("utils.point_cloud").PointCloudProcessor().compute_point_cloud_metrics(cloud)
            got = PointCloudProcessor(verbose=False).compute_point_cloud_metrics(cloud)
            assert set(got) == set(want)
            for k in want:
                assert np.allclose(got[k], want[k], rtol=1e-12, atol=0), k
            empty = PointCloudProcessor(verbose=False).compute_point_cloud_metrics(np.empty((0, 3)))
            assert empty == {'num_points': 0, 'bbox': None, 'center': None, 'dimensions': None}
            img = rng.integers(0, 256, (60, 80, 3), dtype=np.uint8)
            depth = (rng.random((60, 80)) * 6).astype(np.float32)
            depth[rng.random((60, 80)) < 0.3] = 0
            pts = [(float(rng.uniform(-3, 83)), float(rng.uniform(-3, 63))) for _ in range(150)]
            a = l3d.visualize_laser_depth(img, pts, depth, 5.0)
            b = ref_ir.visualize_laser_depth(img, pts, depth, 5.0)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    finally:
        sys.path.remove(REFERENCE)
        for k in list(sys.modules):
            if k.split(".")[0] in ("config", "core", "camera", "utils", "improved_steger", "improved_reconstruction", "main"):
                sys.modules.pop(k, None)
        sys.modules.update(saved)


def test_system_run_realtime_headless(tmp_path):
    """LaserReconstructionSystem.run_realtime (main.py:235-343 without the windows): the loop, the dropped frames, the
    auto-save and the final save, with stand-ins for the camera and the GPU stages (host logic only)."""
    from laser_3d_reconstruction_b200.system import LaserReconstructionSystem

    class Cam:
        def __init__(self):
            self.n, self.stopped = 0, False

        def get_frames(self):
            self.n += 1
            if self.n % 3 == 0:
                return None, None   # a dropped capture
            return np.zeros((4, 6, 3), np.uint8), np.ones((4, 6), np.float32)

        def stop(self):
            self.stopped = True

    class Ext:
        def extract_centerline(self, img):
            return [(1.0, 1.0), (2.0, 2.0)]

    class Rec:
        def reconstruct_from_depth(self, pts, depth):
            return np.array([[0.0, 0.0, 1.0], [np.nan, 0.0, 1.0], [0.1, 0.0, 1.0]])

    saved = []

    class Proc:
        def voxel_downsample(self, p, voxel_size):
            return p

        def statistical_outlier_removal(self, p, nb_neighbors, std_ratio):
            return p

        def save_ply(self, p, path):
            saved.append((len(p), os.path.basename(path)))

    class Cfg(l3d_Config()):
        OUTPUT_DIR = str(tmp_path / "out")
        MIN_POINT_CLOUD_SIZE = 4
        AUTO_SAVE_INTERVAL = 0.0

    s = LaserReconstructionSystem(Cfg(), verbose=False)
    s.camera, s.laser_extractor, s.reconstructor, s.point_cloud_processor = Cam(), Ext(), Rec(), Proc()
    assert s.run_realtime(max_frames=5) == 5
    assert s.camera.stopped and s.camera.n == 7          # 5 processed + 2 dropped captures
    assert len(s.point_cloud) == 10                       # two finite points per frame, the NaN row filtered (main.py:181-184)
    assert saved and saved[-1][1].startswith("final_") and saved[-1][0] == 10
    assert any(not name.startswith("final_") for _, name in saved)   # the auto-save ran inside the loop
    s2 = LaserReconstructionSystem(Cfg(), verbose=False)
    s2.camera, s2.laser_extractor, s2.reconstructor, s2.point_cloud_processor = Cam(), Ext(), Rec(), Proc()
    assert s2.run_realtime(duration=0.05) > 0 and s2.camera.stopped


def l3d_Config():
    from laser_3d_reconstruction_b200.config import Config
    return Config


def test_matcher_accessors_match_cv2_defaults():
    """Every getter cv2.StereoSGBM / cv2.StereoBM exposes exists on the look-alikes and starts at OpenCV's default; setters
    round-trip; what the kernels do not implement is refused."""
    import cv2
    for ref, mine in ((cv2.StereoSGBM_create(), stereo.StereoSGBM_create()), (cv2.StereoBM_create(), stereo.StereoBM_create()),
                      (cv2.StereoSGBM_create(minDisparity=3, numDisparities=96, blockSize=7, P1=11, P2=99, mode=1),
                       stereo.StereoSGBM_create(minDisparity=3, numDisparities=96, blockSize=7, P1=11, P2=99, mode=1)),
                      (cv2.StereoBM_create(numDisparities=48, blockSize=9), stereo.StereoBM_create(numDisparities=48, blockSize=9))):
        for name in dir(ref):
            if name.startswith("get") and name != "getDefaultName":
                assert hasattr(mine, name), name
                want, got = getattr(ref, name)(), getattr(mine, name)()
                assert tuple(np.ravel(want)) == tuple(np.ravel(got)), (type(ref).__name__, name, want, got)
                setter = "set" + name[3:]
                assert hasattr(mine, setter), setter
    bm = stereo.StereoBM_create()
    bm.setSmallerBlockSize(5); bm.setPreFilterSize(11); bm.setPreFilterType(1)
    assert (bm.getSmallerBlockSize(), bm.getPreFilterSize(), bm.getPreFilterType()) == (5, 11, 1)
    with pytest.raises(ValueError):
        bm.setPreFilterType(0)
    bm.setROI1((0, 0, 10, 10))
    with pytest.raises(ValueError):
        bm.compute(np.zeros((8, 8), np.uint8), np.zeros((8, 8), np.uint8))


@pytest.mark.parametrize("sanitize", [False, True])
def test_copy_pool_host_only(tmp_path, sanitize):
    """csrc/hostcopy.cuh's thread pool on the host alone (tests/csrc/copypool_test.cpp): exact copies for every size class
    and thread count, bursts and idle gaps, clean shutdown, a forked child that copies and tears the pool down without
    hanging; once more under ThreadSanitizer where g++ has it (no data race reports)."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = "/usr/local/cuda/include"
    exe = str(tmp_path / "copypool_test")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-pthread", "-I", cuda_inc, "-I", os.path.join(root, "laser_3d_reconstruction_b200", "csrc"),
           os.path.join(root, "tests", "csrc", "copypool_test.cpp"), "-o", exe]
    if sanitize:
        cmd.insert(1, "-fsanitize=thread")
    b = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    if b.returncode != 0:
        if sanitize or not os.path.isdir(cuda_inc):
            pytest.skip("toolchain cannot build this variant: " + b.stderr[-200:])
        raise AssertionError(b.stderr)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "copypool ok" in r.stdout, (r.returncode, r.stdout[-300:], r.stderr[-600:])
    assert "ThreadSanitizer" not in r.stderr
