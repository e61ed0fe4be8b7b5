"""oracle/ref_ops.py (restatement of the reference's Python classes) against the golden vectors
generated from the real reference by tests/golden/make_golden.py."""
import numpy as np
import pytest

from oracle import ref_ops

CFG = dict(hsv_lower=(50, 100, 180), hsv_upper=(70, 255, 255), brightness_threshold=200, min_area=50)


def as_pts(p, dtype=np.float64):
    return np.array(p, dtype=dtype).reshape(-1, 2)


def test_simple(golden_synth):
    g = golden_synth
    assert np.array_equal(as_pts(ref_ops.simple_extract(g["left"], **CFG)), g["simple_cfg"])
    assert np.array_equal(as_pts(ref_ops.simple_extract(g["left"])), g["simple_def"])
    assert len(g["simple_cfg"]) == g["left"].shape[0]  # the synthetic stripe is found on every row


@pytest.mark.parametrize("loop", [False, True])
def test_fast_steger(golden_synth, loop):
    g = golden_synth
    import cv2
    assert np.array_equal(as_pts(ref_ops.fast_steger_extract(g["left"], loop=loop), np.float32), g["fast"])
    if not loop:
        assert np.array_equal(as_pts(ref_ops.fast_steger_extract(g["left"], roi=(100, 40, 150, 200)), np.float32), g["fast_roi"])
        gray = cv2.cvtColor(g["left"], cv2.COLOR_BGR2GRAY)
        assert np.array_equal(as_pts(ref_ops.fast_steger_extract(gray), np.float32), g["fast_gray"])


@pytest.mark.parametrize("loop", [False, True])
def test_improved_variants(golden_synth, loop):
    g = golden_synth
    assert np.array_equal(as_pts(ref_ops.improved_steger_extract(g["left"], loop=loop)), g["improved"])
    assert np.array_equal(as_pts(ref_ops.improved_steger_extract_optimized(g["left"], loop=loop)), g["optimized"])
    assert np.array_equal(as_pts(ref_ops.hybrid_extract(g["left"], loop=loop)), g["hybrid"])


def test_reconstructors(golden_synth):
    from laser_3d_reconstruction_b200 import synth
    g = golden_synth
    sp = [tuple(p) for p in g["simple_cfg"]]
    for name, refr in (("rec_air", False), ("rec_water", True)):
        rec = ref_ops.ReconstructorRef(g["K"], synth.LASER_PLANE, refr)
        assert np.array_equal(rec.reconstruct_from_depth(sp, g["depth"]).reshape(-1, 3), g[name + "_depth"])
        assert np.array_equal(rec.reconstruct_laser_line(sp).reshape(-1, 3), g[name + "_line"])
    rec = ref_ops.ReconstructorRef(g["K"], synth.LASER_PLANE, True)
    fp = [(np.float32(x), np.float32(y)) for x, y in g["fast"]]
    assert np.array_equal(rec.reconstruct_from_depth(fp, g["depth"]).reshape(-1, 3), g["rec_fast_depth"])
    assert np.array_equal(rec.reconstruct_laser_line(fp).reshape(-1, 3), g["rec_fast_line"])
    disp = g["disp16_3way"].astype(np.float32) / 16.0
    ir = ref_ops.ImprovedLaserReconstructorRef(g["Q"])
    ip = [tuple(p) for p in g["optimized"]]
    assert np.array_equal(ir.reconstruct_from_disparity(ip, disp).reshape(-1, 3), g["irec_disp"])
    assert np.array_equal(ir.reconstruct_with_interpolation(ip, disp, 3, 1.0).reshape(-1, 3), g["irec_interp"])


def test_depth_path_real_pair(golden_real):
    """The camera class's rectification + as-constructed matcher on a real 320x240 pair."""
    import cv2
    g = golden_real
    size = (320, 240)
    R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(g["K_left"], g["dist_left"], g["K_right"], g["dist_right"], size,
                                                g["R"], g["T"], flags=cv2.CALIB_ZERO_DISPARITY, alpha=0)
    assert np.array_equal(Q, g["Q"])
    mlx, mly = cv2.initUndistortRectifyMap(g["K_left"], g["dist_left"], R1, P1, size, cv2.CV_32FC1)
    mrx, mry = cv2.initUndistortRectifyMap(g["K_right"], g["dist_right"], R2, P2, size, cv2.CV_32FC1)
    assert np.array_equal(mlx[::8, ::8], g["map_left_x"])
    for tag in ("a", "b"):
        frame = g["frame_" + tag]
        l, r = frame[:, :320], frame[:, 320:]
        lrect, depth, aux = ref_ops.depth_path(l, r, (mlx, mly, mrx, mry), 64, 5, cv2.STEREO_SGBM_MODE_SGBM_3WAY, Q,
                                               use_wls=False, want_all=True)
        assert np.array_equal(lrect, g["lrect_" + tag])
        assert np.array_equal(aux["dl"], g["disp16_" + tag])
        assert np.array_equal(depth, g["depth_" + tag])


def test_point_cloud_sink_restatement_matches_golden():
    """oracle restatement of the reference's Open3D-free PointCloudProcessor code against vectors the real
    class produced (tests/golden/make_golden_cloud.py)."""
    import os
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cloud.npz")))
    cloud = g["cloud"]
    assert np.array_equal(ref_ops.simple_voxel_downsample(cloud, 0.002), g["voxel_2mm"])
    assert np.array_equal(ref_ops.simple_voxel_downsample(cloud, 0.005), g["voxel_5mm"])
    assert np.array_equal(ref_ops.simple_outlier_removal(cloud, 20, 2.0), g["sor_20_2"])
    assert ref_ops.simple_outlier_removal(cloud, 8, 0.0).shape == g["sor_8_0"].shape == (0,)
    assert ref_ops.simple_outlier_removal(cloud, 20, 1e-20).shape == (0,)
    v32 = ref_ops.simple_voxel_downsample(cloud.astype(np.float32), 0.002)  # main.py:208: float32 cloud, float32 arithmetic
    assert v32.dtype == np.float32 and np.array_equal(v32, g["voxel_2mm_f32"])
    assert np.array_equal(ref_ops.simple_outlier_removal(v32, 20, 2.0), g["sor_20_2_f32"])
