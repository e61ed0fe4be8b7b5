"""cv2.stereoRectify without cv2 (reference camera/single_usb_stereo_camera.py:176-187): the init-time
3x3 algebra that turns the calibration (K1, d1, K2, d2, R, T) into R1, R2, P1, P2, Q and the two valid-pixel rectangles.

It follows OpenCV 4.x's algorithm step by step in float64 (with its float32 corner-point arrays where OpenCV has
them): average rotation by Rodrigues, alignment of the baseline with the x (or y) axis, new focal length = mean of the two
fy (fx) values, new principal points from the four undistorted image corners, CALIB_ZERO_DISPARITY averaging, and the
`alpha` zoom from the inner / outer rectangles of a 9 x 9 grid of undistorted points.

One step cannot be repeated bit for bit: OpenCV first re-orthogonalises R with its own Jacobi SVD before taking the
rotation vector.  For a calibration file's R (orthonormal to ~1e-16) that changes the last bit or two of R1 / R2; the f32
rectification maps built from the result are unchanged.  tests/test_oracle_cv2.py pins matrices (<= 1e-13), rectangles
(==) and maps (==) against cv2 on the shipped calibration and on synthetic rigs.

Differential fuzzing against cv2.stereoRectify over random rigs (4 / 5 / 8 / 12 / 14 distortion coefficients, horizontal and
vertical baselines, every alpha; tests/fuzz/fuzz_rectify.py, 2 400 rigs) pinned three more points: the OUTER rectangle spans the
border points of the 9 x 9 grid only; points the inverse distortion model turns into NaN are skipped by OpenCV's MIN / MAX
macros (and a rig that is NaN at the image corners yields NaN matrices and empty ROIs, as in cv2); coefficients 13 / 14 are
the tilted-sensor model.  What cannot be pinned: at alpha = 0 one ROI edge lies exactly on the image border, so
`ceil(0 +- 1e-14)` decides between (0, 0, w, h) and (0, 1, w, h - 1) on the last bits of R1 / R2 (about 1.5 % of random
rigs differ from cv2 by that one pixel; the shipped calibration does not).
"""
import numpy as np

CALIB_ZERO_DISPARITY = 1024  # cv2.CALIB_ZERO_DISPARITY


def rodrigues_to_vector(R):
    """rotation matrix -> rotation vector (cv2.Rodrigues, matrix input)"""
    R = np.asarray(R, np.float64).reshape(3, 3)
    U, _, Vt = np.linalg.svd(R)
    R = U @ Vt
    r = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = np.sqrt((r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) * 0.25)
    c = (R[0, 0] + R[1, 1] + R[2, 2] - 1) * 0.5
    c = min(max(c, -1.0), 1.0)
    theta = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t = (R[0, 0] + 1) * 0.5
        rx = np.sqrt(max(t, 0.0))
        t = (R[1, 1] + 1) * 0.5
        ry = np.sqrt(max(t, 0.0)) * (-1.0 if R[0, 1] < 0 else 1.0)
        t = (R[2, 2] + 1) * 0.5
        rz = np.sqrt(max(t, 0.0)) * (-1.0 if R[0, 2] < 0 else 1.0)
        if abs(rx) < abs(ry) and abs(rx) < abs(rz) and (R[1, 2] > 0) != (ry * rz > 0):
            rz = -rz
        v = np.array([rx, ry, rz])
        return v * (theta / np.linalg.norm(v))
    vth = 1.0 / (2.0 * s)
    vth *= theta
    return r * vth


def rodrigues_to_matrix(om):
    """rotation vector -> rotation matrix (cv2.Rodrigues, vector input)"""
    om = np.asarray(om, np.float64).reshape(3)
    theta = np.sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2])
    if theta < np.finfo(np.float64).eps:
        return np.eye(3)
    c, s = np.cos(theta), np.sin(theta)
    c1 = 1.0 - c
    itheta = 1.0 / theta
    r = om * itheta
    rrt = np.outer(r, r)
    rx = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    return c * np.eye(3) + c1 * rrt + s * rx


def _dist14(d):
    k = np.zeros(14)
    d = np.asarray(d, np.float64).reshape(-1)
    k[:len(d)] = d
    return k


def undistort_points(pts, K, dist, R=None, P=None, dtype=np.float32):
    """cv2.undistortPoints: 5 fixed-point iterations of the inverse distortion model in f64, then the optional rotation R
    and projection P; input and result have `dtype` (OpenCV's CV_32FC2 corner arrays / CV_64FC2 grid arrays)."""
    K = np.asarray(K, np.float64)
    k = _dist14(dist)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    ifx, ify = 1.0 / fx, 1.0 / fy
    out = np.empty((len(pts), 2), dtype)
    RR = np.eye(3) if R is None else np.asarray(R, np.float64).reshape(3, 3)
    if P is not None:
        P = np.asarray(P, np.float64)
        RR = P[:3, :3] @ RR
    pts64 = np.asarray(pts, dtype).astype(np.float64)
    inv_tilt = _inv_tilt_matrix(k[12], k[13]) if (k[12] != 0 or k[13] != 0) else None
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):   # a folding model may overflow: NaN out, like cv2
        _undistort_loop(pts64, out, k, cx, cy, ifx, ify, RR, inv_tilt)
    return out


def _inv_tilt_matrix(tau_x, tau_y):
    """inverse of OpenCV's tilted-sensor projection (computeTiltProjectionMatrix; distortion coefficients 13 and 14)"""
    cx_, sx_, cy_, sy_ = np.cos(tau_x), np.sin(tau_x), np.cos(tau_y), np.sin(tau_y)
    rot_x = np.array([[1, 0, 0], [0, cx_, sx_], [0, -sx_, cx_]])
    rot_y = np.array([[cy_, 0, -sy_], [0, 1, 0], [sy_, 0, cy_]])
    rot_xy = rot_y @ rot_x
    inv = 1.0 / rot_xy[2, 2]
    inv_proj_z = np.array([[inv, 0, inv * rot_xy[0, 2]], [0, inv, inv * rot_xy[1, 2]], [0, 0, 1]])
    return rot_xy.T @ inv_proj_z


def _undistort_loop(pts64, out, k, cx, cy, ifx, ify, RR, inv_tilt=None):
    for i, (u, v) in enumerate(pts64):
        x = (u - cx) * ifx
        y = (v - cy) * ify
        if inv_tilt is not None:   # compensate the sensor tilt first
            vx = inv_tilt[0, 0] * x + inv_tilt[0, 1] * y + inv_tilt[0, 2]
            vy = inv_tilt[1, 0] * x + inv_tilt[1, 1] * y + inv_tilt[1, 2]
            vz = inv_tilt[2, 0] * x + inv_tilt[2, 1] * y + inv_tilt[2, 2]
            ip = 1.0 / vz if vz else 1.0
            x, y = ip * vx, ip * vy
        x0, y0 = x, y
        for _ in range(5):
            r2 = x * x + y * y
            icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
            if icdist < 0:
                x, y = (u - cx) * ifx, (v - cy) * ify
                break
            dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2
            dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2
            x = (x0 - dx) * icdist
            y = (y0 - dy) * icdist
        xx = RR[0, 0] * x + RR[0, 1] * y + RR[0, 2]
        yy = RR[1, 0] * x + RR[1, 1] * y + RR[1, 2]
        ww = 1.0 / (RR[2, 0] * x + RR[2, 1] * y + RR[2, 2])
        out[i, 0] = xx * ww
        out[i, 1] = yy * ww


def _rectangles(K, dist, R, P, size):
    """getRectangles (OpenCV 4.x, double precision): inner and outer bounding rectangles (x, y, w, h) of a 9 x 9 grid of
    undistorted points"""
    N = 9
    w, h = size
    pts = np.array([[float(x) * (w - 1) / (N - 1), float(y) * (h - 1) / (N - 1)] for y in range(N) for x in range(N)], np.float64)
    u = undistort_points(pts, K, dist, R, P, dtype=np.float64).reshape(N, N, 2)
    # OpenCV folds the points in with MIN / MAX macros whose comparisons are false for a NaN operand, so a point the
    # inverse distortion model cannot handle (it yields NaN there) is simply skipped
    fmax, fmin = np.fmax.reduce, np.fmin.reduce
    inner_x0, inner_x1 = fmax(u[:, 0, 0]), fmin(u[:, N - 1, 0])
    inner_y0, inner_y1 = fmax(u[0, :, 1]), fmin(u[N - 1, :, 1])
    # the outer rectangle spans the grid's BORDER points only (cv2 4.13; found by fuzzing against cv2.stereoRectify: with a
    # distortion model that folds inside the image, interior grid points land far outside and would blow the rectangle up)
    b = np.concatenate([u[0], u[N - 1], u[:, 0], u[:, N - 1]])
    outer_x0, outer_x1 = fmin(b[:, 0]), fmax(b[:, 0])
    outer_y0, outer_y1 = fmin(b[:, 1]), fmax(b[:, 1])
    inner = (inner_x0, inner_y0, inner_x1 - inner_x0, inner_y1 - inner_y0)
    outer = (outer_x0, outer_y0, outer_x1 - outer_x0, outer_y1 - outer_y0)
    return inner, outer


def stereo_rectify(K1, d1, K2, d2, size, R, T, flags=CALIB_ZERO_DISPARITY, alpha=-1.0):
    """-> R1, R2, P1, P2, Q, roi1, roi2 as cv2.stereoRectify(K1, d1, K2, d2, size, R, T, flags=flags, alpha=alpha)."""
    K1, K2 = np.asarray(K1, np.float64).reshape(3, 3), np.asarray(K2, np.float64).reshape(3, 3)
    T = np.asarray(T, np.float64).reshape(3)
    nx, ny = int(size[0]), int(size[1])
    R = np.asarray(R, np.float64)
    om = rodrigues_to_vector(R) if R.size == 9 else R.reshape(3).copy()
    om = om * -0.5
    r_r = rodrigues_to_matrix(om)
    t = r_r @ T
    idx = 0 if abs(t[0]) > abs(t[1]) else 1
    c = t[idx]
    nt = np.sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2])
    if not nt > 0.0:
        raise ValueError("stereo_rectify: zero baseline")
    uu = np.zeros(3)
    uu[idx] = 1.0 if c > 0 else -1.0
    ww = np.cross(t, uu)
    nw = np.sqrt(ww[0] * ww[0] + ww[1] * ww[1] + ww[2] * ww[2])
    if nw > 0.0:
        ww = ww * (np.arccos(abs(c) / nt) / nw)
    wR = rodrigues_to_matrix(ww)
    R1 = wR @ r_r.T
    R2 = wR @ r_r
    t = R2 @ T
    ratio = 0.5  # newImageSize == imageSize
    fc_new = (K1[idx ^ 1, idx ^ 1] + K2[idx ^ 1, idx ^ 1]) * ratio
    cc_new = np.zeros((2, 2))
    corners = np.array([[(i % 2) * (nx - 1), (0 if i < 2 else 1) * (ny - 1)] for i in range(4)], np.float32)
    for k, (A, Dk, Rk) in enumerate(((K1, d1, R1), (K2, d2, R2))):
        und = undistort_points(corners, A, Dk)  # f32 normalised points
        # cvProjectPoints2 with rotation Rk, zero translation, camera matrix diag(fc_new, fc_new, 1), no distortion -> f32
        proj = np.empty((4, 2), np.float32)
        with np.errstate(invalid="ignore", over="ignore"):   # NaN corners (a model that cannot be inverted there) pass through
            for i in range(4):
                X = np.array([np.float64(und[i, 0]), np.float64(und[i, 1]), 1.0])
                Y = Rk @ X
                z = 1.0 / Y[2] if Y[2] != 0 else 1.0
                proj[i, 0] = np.float32(Y[0] * z * fc_new)
                proj[i, 1] = np.float32(Y[1] * z * fc_new)
        avg = proj.astype(np.float64).sum(axis=0) / 4.0
        cc_new[k, 0] = (nx - 1) / 2 - avg[0]
        cc_new[k, 1] = (ny - 1) / 2 - avg[1]
    if flags & CALIB_ZERO_DISPARITY:
        cc_new[0, 0] = cc_new[1, 0] = (cc_new[0, 0] + cc_new[1, 0]) * 0.5
        cc_new[0, 1] = cc_new[1, 1] = (cc_new[0, 1] + cc_new[1, 1]) * 0.5
    elif idx == 0:
        cc_new[0, 1] = cc_new[1, 1] = (cc_new[0, 1] + cc_new[1, 1]) * 0.5
    else:
        cc_new[0, 0] = cc_new[1, 0] = (cc_new[0, 0] + cc_new[1, 0]) * 0.5
    P1 = np.zeros((3, 4))
    P1[0, 0] = P1[1, 1] = fc_new
    P1[0, 2], P1[1, 2], P1[2, 2] = cc_new[0, 0], cc_new[0, 1], 1.0
    P2 = P1.copy()
    P2[0, 2], P2[1, 2] = cc_new[1, 0], cc_new[1, 1]
    P2[idx, 3] = t[idx] * fc_new
    alpha = min(alpha, 1.0)
    inner1, outer1 = _rectangles(K1, d1, R1, P1, (nx, ny))
    inner2, outer2 = _rectangles(K2, d2, R2, P2, (nx, ny))
    cx1_0, cy1_0, cx2_0, cy2_0 = cc_new[0, 0], cc_new[0, 1], cc_new[1, 0], cc_new[1, 1]
    cx1, cy1, cx2, cy2 = nx * cx1_0 / nx, ny * cy1_0 / ny, nx * cx2_0 / nx, ny * cy2_0 / ny
    s = 1.0
    if alpha >= 0:
        def rect_scale(cx, cy, cx0, cy0, r, pick):
            x, y, w, h = r
            return pick(pick(pick(cx / (cx0 - x), cy / (cy0 - y)), (nx - 1 - cx) / (x + w - cx0)), (ny - 1 - cy) / (y + h - cy0))
        s0 = max(rect_scale(cx1, cy1, cx1_0, cy1_0, inner1, max), rect_scale(cx2, cy2, cx2_0, cy2_0, inner2, max))
        s1 = min(rect_scale(cx1, cy1, cx1_0, cy1_0, outer1, min), rect_scale(cx2, cy2, cx2_0, cy2_0, outer2, min))
        s = s0 * (1 - alpha) + s1 * alpha
    fc_new *= s
    cc_new = np.array([[cx1, cy1], [cx2, cy2]])
    P1[0, 0] = P1[1, 1] = fc_new
    P1[0, 2], P1[1, 2] = cx1, cy1
    P2[0, 0] = P2[1, 1] = fc_new
    P2[0, 2], P2[1, 2] = cx2, cy2
    P2[idx, 3] = s * P2[idx, 3]

    def roi(inner, cx0, cy0, cx, cy):
        x, y, w, h = inner
        if not np.all(np.isfinite([x, y, w, h, s, cx, cy, cx0, cy0])):
            return (0, 0, 0, 0)   # a rig whose model cannot be inverted at the image corners: cv2 returns NaN matrices and empty ROIs
        x0, y0 = int(np.ceil((x - cx0) * s + cx)), int(np.ceil((y - cy0) * s + cy))
        w0, h0 = int(np.floor(w * s)), int(np.floor(h * s))
        xa, ya = max(x0, 0), max(y0, 0)
        xb, yb = min(x0 + w0, nx), min(y0 + h0, ny)
        if xb <= xa or yb <= ya:
            return (0, 0, 0, 0)
        return (xa, ya, xb - xa, yb - ya)

    roi1 = roi(inner1, cx1_0, cy1_0, cx1, cy1)
    roi2 = roi(inner2, cx2_0, cy2_0, cx2, cy2)
    Q = np.array([[1, 0, 0, -cc_new[0, 0]],
                  [0, 1, 0, -cc_new[0, 1]],
                  [0, 0, 0, fc_new],
                  [0, 0, -1.0 / t[idx], (cc_new[0, 0] - cc_new[1, 0] if idx == 0 else cc_new[0, 1] - cc_new[1, 1]) / t[idx]]])
    return R1, R2, P1, P2, Q, roi1, roi2
