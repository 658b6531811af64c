"""TEST INFRASTRUCTURE ONLY -- CPU oracle for PnPSolver::solvePnP.

Reference: /root/reference/src/pnp_solver.cpp:7-52 -- image points {left.bottom, left.top,
right.top, right.bottom} (:41-44), always the small-armor object points (:47-48, 23-28), then
cv::solvePnP(obj, img, K, D, rvec, tvec, false, SOLVEPNP_IPPE) (:49-51) with K/D from
/root/reference/config/camera_info.yaml:4-12.

The arithmetic lives in OpenCV calib3d (system dependency, version unpinned by package.xml:13);
the binary oracle in this image is cv2 4.13.0 (`solve_cv2`).  `solve_ippe` restates the
published algorithm (Collins & Bartoli, "Infinitesimal Plane-based Pose Estimation", IJCV 2014,
as implemented in OpenCV modules/calib3d/src/ippe.cpp and undistort.dispatch.cpp):
  1. undistortPoints: normalise with K, then exactly 5 fixed-point iterations of the
     plumb-bob inverse (TermCriteria default count 5), result rounded to FP32 (the output
     array takes the Point2f input depth -- measured: without that rounding the restatement is
     5e-6 off cv2, with it 6e-11);
  2. canonical object frame (centred, planar, z = 0) -- constant per armor size;
  3. exact homography canonical plane -> normalised image from the 4 correspondences;
  4. IPPE: Jacobian of the homography at the origin -> two rotations; translation by linear LSQ;
  5. back to the model frame, order the two poses by reprojection RMSE in normalised
     coordinates, return the lower;
  6. rotation matrix -> rvec with IPPE's rot2vec (axis from the skew part, angle = acos).
It is pinned against cv2.solvePnP on seeded quads in tests/test_oracle_cpu.py (<=1e-9 relative)
and on the survey's known-answer vector.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

# /root/reference/config/camera_info.yaml:4-12
K_DEFAULT = np.array([957.669211, 0.0, 345.943891, 0.0, 969.127115, 284.057302, 0.0, 0.0, 1.0])
D_DEFAULT = np.array([-0.405274, 0.126058, -0.026939, -0.006503, 0.0])

# /root/reference/include/irmv_detection/pnp_solver.hpp:30-33 (mm)
SMALL_W, SMALL_H, LARGE_W, LARGE_H = 135.0, 55.0, 225.0, 55.0


def object_points(large: bool = False) -> np.ndarray:
    """/root/reference/src/pnp_solver.cpp:17-33: LB, LT, RT, RB; x forward, y left, z up (m)."""
    hy = (LARGE_W if large else SMALL_W) / 2.0 / 1000.0
    hz = (LARGE_H if large else SMALL_H) / 2.0 / 1000.0
    return np.array([[0, hy, -hz], [0, hy, hz], [0, -hy, hz], [0, -hy, -hz]], np.float64)


def undistort_points(pts: np.ndarray, K=K_DEFAULT, D=D_DEFAULT, iters: int = 5) -> np.ndarray:
    """pts f64[...,2] pixel -> normalised undistorted coordinates (cv::undistortPoints, no R/P)."""
    fx, fy, cx, cy = K[0], K[4], K[2], K[5]
    k1, k2, p1, p2, k3 = D[:5]
    x0 = (pts[..., 0] - cx) / fx
    y0 = (pts[..., 1] - cy) / fy
    x, y = x0.copy(), y0.copy()
    for _ in range(iters):
        r2 = x * x + y * y
        icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2)
        # OpenCV: a negative icdist aborts the iteration with the initial guess
        bad = icdist < 0
        dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
        dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y
        xn = (x0 - dx) * icdist
        yn = (y0 - dy) * icdist
        x = np.where(bad, x0, xn)
        y = np.where(bad, y0, yn)
    return np.stack((x, y), -1)


def _homography_rect_to_quad(a: float, b: float, q: np.ndarray) -> np.ndarray:
    """Exact homography taking the canonical rectangle corners to the 4 image points.

    Canonical corner order follows object_points(): (+a,-b), (+a,+b), (-a,+b), (-a,-b) with
    canonical X = model y, canonical Y = model z.  Square->quad closed form (Heckbert 1989)
    composed with the affine rectangle->unit-square map.  q f64[n,4,2].  Returns H f64[n,3,3],
    normalised so H[2,2] = 1.
    """
    x0, y0 = q[:, 0, 0], q[:, 0, 1]
    x1, y1 = q[:, 1, 0], q[:, 1, 1]
    x2, y2 = q[:, 2, 0], q[:, 2, 1]
    x3, y3 = q[:, 3, 0], q[:, 3, 1]
    dx1, dx2, sx = x1 - x2, x3 - x2, x0 - x1 + x2 - x3
    dy1, dy2, sy = y1 - y2, y3 - y2, y0 - y1 + y2 - y3
    det = dx1 * dy2 - dx2 * dy1
    g = (sx * dy2 - dx2 * sy) / det
    h = (dx1 * sy - sx * dy1) / det
    n = q.shape[0]
    Hs = np.empty((n, 3, 3))
    Hs[:, 0, 0] = x1 - x0 + g * x1
    Hs[:, 0, 1] = x3 - x0 + h * x3
    Hs[:, 0, 2] = x0
    Hs[:, 1, 0] = y1 - y0 + g * y1
    Hs[:, 1, 1] = y3 - y0 + h * y3
    Hs[:, 1, 2] = y0
    Hs[:, 2, 0] = g
    Hs[:, 2, 1] = h
    Hs[:, 2, 2] = 1.0
    # unit square (u,v): corner0=(0,0), corner1=(1,0), corner2=(1,1), corner3=(0,1)
    # canonical (X,Y): corner0=(+a,-b), corner1=(+a,+b), corner2=(-a,+b), corner3=(-a,-b)
    #   => u = (Y + b)/(2b), v = (a - X)/(2a)
    A = np.array([[0.0, 1.0 / (2 * b), 0.5], [-1.0 / (2 * a), 0.0, 0.5], [0.0, 0.0, 1.0]])
    H = Hs @ A
    return H / H[:, 2:3, 2:3]


def _rot_z_to(v: np.ndarray) -> np.ndarray:
    """Rotation Rv with Rv @ e_z = v/|v| (Rodrigues about e_z x v). v f64[n,3]."""
    a = v / np.linalg.norm(v, axis=1, keepdims=True)
    ax, ay, az = a[:, 0], a[:, 1], a[:, 2]
    n = v.shape[0]
    R = np.empty((n, 3, 3))
    k = 1.0 / (1.0 + az)
    R[:, 0, 0] = 1.0 - ax * ax * k
    R[:, 0, 1] = -ax * ay * k
    R[:, 0, 2] = ax
    R[:, 1, 0] = -ax * ay * k
    R[:, 1, 1] = 1.0 - ay * ay * k
    R[:, 1, 2] = ay
    R[:, 2, 0] = -ax
    R[:, 2, 1] = -ay
    R[:, 2, 2] = az
    return R


def _ippe_rotations(H: np.ndarray):
    j00 = H[:, 0, 0] - H[:, 2, 0] * H[:, 0, 2]
    j01 = H[:, 0, 1] - H[:, 2, 1] * H[:, 0, 2]
    j10 = H[:, 1, 0] - H[:, 2, 0] * H[:, 1, 2]
    j11 = H[:, 1, 1] - H[:, 2, 1] * H[:, 1, 2]
    p, q = H[:, 0, 2], H[:, 1, 2]
    n = H.shape[0]
    Rv = _rot_z_to(np.stack((p, q, np.ones(n)), 1))
    b00 = Rv[:, 0, 0] - p * Rv[:, 2, 0]
    b01 = Rv[:, 0, 1] - p * Rv[:, 2, 1]
    b10 = Rv[:, 1, 0] - q * Rv[:, 2, 0]
    b11 = Rv[:, 1, 1] - q * Rv[:, 2, 1]
    dtinv = 1.0 / (b00 * b11 - b01 * b10)
    bi00, bi01, bi10, bi11 = dtinv * b11, -dtinv * b01, -dtinv * b10, dtinv * b00
    a00 = bi00 * j00 + bi01 * j10
    a01 = bi00 * j01 + bi01 * j11
    a10 = bi10 * j00 + bi11 * j10
    a11 = bi10 * j01 + bi11 * j11
    ata00 = a00 * a00 + a01 * a01
    ata01 = a00 * a10 + a01 * a11
    ata11 = a10 * a10 + a11 * a11
    gamma = np.sqrt(0.5 * (ata00 + ata11 + np.sqrt((ata00 - ata11) ** 2 + 4.0 * ata01 * ata01)))
    r00, r01, r10, r11 = a00 / gamma, a01 / gamma, a10 / gamma, a11 / gamma
    b0 = np.sqrt(np.maximum(1.0 - r00 * r00 - r10 * r10, 0.0))
    b1 = np.sqrt(np.maximum(1.0 - r01 * r01 - r11 * r11, 0.0))
    sp = -(r00 * r01 + r10 * r11)
    b1 = np.where(sp < 0, -b1, b1)
    out = []
    for sgn in (1.0, -1.0):
        c1 = np.stack((r00, r10, sgn * b0), 1)
        c2 = np.stack((r01, r11, sgn * b1), 1)
        c3 = np.cross(c1, c2)
        Rt = np.stack((c1, c2, c3), 2)
        out.append(Rv @ Rt)
    return out


def _ippe_translation(X: np.ndarray, q: np.ndarray, R: np.ndarray) -> np.ndarray:
    """Linear least squares for t given R; X f64[4,2] canonical, q f64[n,4,2], R f64[n,3,3]."""
    Px = R[:, 0:1, 0] * X[None, :, 0] + R[:, 0:1, 1] * X[None, :, 1]
    Py = R[:, 1:2, 0] * X[None, :, 0] + R[:, 1:2, 1] * X[None, :, 1]
    Pz = R[:, 2:3, 0] * X[None, :, 0] + R[:, 2:3, 1] * X[None, :, 1]
    u, v = q[..., 0], q[..., 1]
    npts = X.shape[0]
    # normal equations of rows [1,0,-u | u*Pz-Px], [0,1,-v | v*Pz-Py]
    su, sv = u.sum(1), v.sum(1)
    suv2 = (u * u + v * v).sum(1)
    bx = (u * Pz - Px)
    by = (v * Pz - Py)
    r0, r1 = bx.sum(1), by.sum(1)
    r2 = -(u * bx + v * by).sum(1)
    n = q.shape[0]
    A = np.zeros((n, 3, 3))
    A[:, 0, 0] = npts
    A[:, 1, 1] = npts
    A[:, 0, 2] = A[:, 2, 0] = -su
    A[:, 1, 2] = A[:, 2, 1] = -sv
    A[:, 2, 2] = suv2
    rhs = np.stack((r0, r1, r2), 1)
    return np.linalg.solve(A, rhs[..., None])[..., 0]


def rot2vec(R: np.ndarray) -> np.ndarray:
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    w = np.arccos(np.clip((tr - 1.0) / 2.0, -1.0, 1.0))
    eps = np.finfo(np.float32).eps
    with np.errstate(divide="ignore", invalid="ignore"):
        d = w / (2.0 * np.sin(w))
    c = np.stack((R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]), 1)
    r = d[:, None] * c
    r[w < eps] = 0.0
    return r


def solve_ippe(img_pts: np.ndarray, K=K_DEFAULT, D=D_DEFAULT, large: bool = False,
               both: bool = False):
    """img_pts f32/f64[n,4,2] (LB,LT,RT,RB px) -> rvec f64[n,3], tvec f64[n,3]."""
    pts = np.asarray(img_pts, np.float32).astype(np.float64).reshape(-1, 4, 2)
    # cv::undistortPoints writes its result in the input depth (CV_32FC2 for Point2f input,
    # /root/reference/src/pnp_solver.cpp:37-44), so IPPE sees FP32-rounded normalised points.
    q = undistort_points(pts, K, D).astype(np.float32).astype(np.float64)
    obj = object_points(large)
    a, b = obj[0, 1], obj[1, 2]                      # half width (model y), half height (model z)
    Xc = np.stack((obj[:, 1], obj[:, 2]), 1)         # canonical X = model y, Y = model z
    H = _homography_rect_to_quad(a, b, q)
    Ra, Rb = _ippe_rotations(H)
    sols = []
    # canonical -> model: [X,Y,Z]_c = C @ [x,y,z]_m with rows (0,1,0),(0,0,1),(1,0,0), det = +1
    C = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]])
    for R in (Ra, Rb):
        t = _ippe_translation(Xc, q, R)
        Rm = R @ C
        P = np.einsum("nij,kj->nki", Rm, obj) + t[:, None, :]
        proj = P[..., :2] / P[..., 2:3]
        err = np.sqrt(((proj - q) ** 2).sum((1, 2)) / (2.0 * 4))
        sols.append((Rm, t, err))
    first = sols[0][2].astype(np.float32) < sols[1][2].astype(np.float32)   # OpenCV compares floats
    # OpenCV: if (err1 < err2) keep order else swap  (ties -> second)
    R1 = np.where(first[:, None, None], sols[0][0], sols[1][0])
    t1 = np.where(first[:, None], sols[0][1], sols[1][1])
    if not both:
        return rot2vec(R1), t1
    R2 = np.where(first[:, None, None], sols[1][0], sols[0][0])
    t2 = np.where(first[:, None], sols[1][1], sols[0][1])
    e1 = np.where(first, sols[0][2], sols[1][2])
    e2 = np.where(first, sols[1][2], sols[0][2])
    return rot2vec(R1), t1, rot2vec(R2), t2, e1, e2


def solve_cv2(img_pts: np.ndarray, K=K_DEFAULT, D=D_DEFAULT, large: bool = False):
    """The binary oracle: cv2.solvePnP(SOLVEPNP_IPPE) per quad, exactly the reference's call."""
    import cv2
    Km = np.asarray(K, np.float64).reshape(3, 3)
    Dm = np.asarray(D, np.float64).reshape(1, 5)
    obj = object_points(large)
    pts = np.asarray(img_pts, np.float32).reshape(-1, 4, 2)
    rv = np.empty((pts.shape[0], 3))
    tv = np.empty((pts.shape[0], 3))
    ok = np.zeros(pts.shape[0], bool)
    for i in range(pts.shape[0]):
        try:
            o, r, t = cv2.solvePnP(obj, pts[i], Km, Dm, flags=cv2.SOLVEPNP_IPPE)
        except cv2.error:
            o = False
        ok[i] = bool(o)
        if o:
            rv[i] = r.reshape(3)
            tv[i] = t.reshape(3)
    return rv, tv, ok


def project(rvec, tvec, K=K_DEFAULT, D=D_DEFAULT, large=False):
    import cv2
    Km = np.asarray(K, np.float64).reshape(3, 3)
    Dm = np.asarray(D, np.float64).reshape(1, 5)
    p, _ = cv2.projectPoints(object_points(large), np.asarray(rvec, np.float64),
                             np.asarray(tvec, np.float64), Km, Dm)
    return p.reshape(4, 2)


def synth_quads(n: int, seed: int = 0, noise_px: float = 0.5, K=K_DEFAULT, D=D_DEFAULT,
                img_w: int = 640, img_h: int = 480):
    """seeded armor poses projected through K/D (+ pixel noise).  The
    generator lives with the other synthetic inputs (irmv_detection_b200/synth.py: armor_quads)."""
    from irmv_detection_b200 import synth
    return synth.armor_quads(n, seed, noise_px, K, D, img_w, img_h)


def refine_lm_cv2(img_pts: np.ndarray, rvecs: np.ndarray, tvecs: np.ndarray, K=K_DEFAULT, D=D_DEFAULT,
                  large: bool = False):
    """Oracle of the optional LM stage (NOT in the reference's call): cv2.solvePnPRefineLM with a
    tight termination criterion, started from the given poses."""
    import cv2
    obj = object_points(large)
    Km = np.asarray(K, np.float64).reshape(3, 3)
    Dm = np.asarray(D, np.float64).reshape(-1, 1)
    out_r, out_t = np.empty_like(rvecs), np.empty_like(tvecs)
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_COUNT, 200, 1e-15)
    for i in range(len(img_pts)):
        r, t = cv2.solvePnPRefineLM(obj, np.asarray(img_pts[i], np.float64).reshape(4, 1, 2), Km, Dm,
                                    rvecs[i].reshape(3, 1).copy(), tvecs[i].reshape(3, 1).copy(), crit)
        out_r[i], out_t[i] = r.ravel(), t.ravel()
    return out_r, out_t


def reprojection_rmse_px(img_pts, rvecs, tvecs, K=K_DEFAULT, D=D_DEFAULT, large=False):
    import cv2
    obj = object_points(large)
    Km = np.asarray(K, np.float64).reshape(3, 3)
    Dm = np.asarray(D, np.float64).reshape(-1, 1)
    out = np.empty(len(img_pts))
    for i in range(len(img_pts)):
        pr, _ = cv2.projectPoints(obj, rvecs[i].reshape(3, 1), tvecs[i].reshape(3, 1), Km, Dm)
        out[i] = np.sqrt(np.mean((pr.reshape(4, 2) - np.asarray(img_pts[i], np.float64).reshape(4, 2)) ** 2))
    return out
