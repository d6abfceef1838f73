"""Oracle: DLT triangulation (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates, in float64 numpy/scipy:
  * utils.DLT                          (reference utils.py:19-34)
  * utils.get_projection_matrix        (reference utils.py:425-435, 803-805)
  * cv.undistortPoints(..., P=cmtx)    (reference call sites utils.py:1314-1315;
                                        algorithm = OpenCV 4.x undistortPoints,
                                        5 fixed-point iterations, pinned against
                                        cv2 outputs in tests/golden/)
  * utils.triangulate_points           (reference utils.py:1277-1336)
  * pose_estimation.get_pose_3D        (reference pose_estimation.py:11-65)
and extends utils.DLT to V confidence-weighted views the way SURVEY.md §8(c)
prescribes (rows of utils.py:21-24 scaled by w_v, in view order; B = A^T A
utils.py:28; SVD of B utils.py:29; Vh[3,:3]/Vh[3,3] utils.py:34).  The V=2,
w=1 case is pinned to utils.DLT by tests/test_oracle_golden.py.
"""
import numpy as np
from scipy import linalg


def projection_matrix(cmtx, R, T):
    """P = K [R|T]  (utils.py:425-435; inline form utils.py:1318-1319)."""
    cmtx = np.asarray(cmtx, dtype=np.float64)
    Rt = np.hstack([np.asarray(R, dtype=np.float64).reshape(3, 3),
                    np.asarray(T, dtype=np.float64).reshape(3, 1)])
    return cmtx @ Rt


def dlt_pair(P1, P2, point1, point2):
    """Literal restatement of utils.DLT (utils.py:19-34): 2 views, unweighted."""
    A = np.array([point1[1] * P1[2, :] - P1[1, :],
                  P1[0, :] - point1[0] * P1[2, :],
                  point2[1] * P2[2, :] - P2[1, :],
                  P2[0, :] - point2[0] * P2[2, :]]).reshape(4, 4)
    B = A.T @ A
    _, _, Vh = linalg.svd(B, full_matrices=False)
    return Vh[3, 0:3] / Vh[3, 3]


def dlt_rows(kpts, P):
    """Rows of the weighted DLT system, A: (N, 2V, 4).

    kpts: (N, V, 3) = [x, y, w];  P: (V, 3, 4).  Row order and signs follow
    utils.py:21-24; each row of view v is scaled by w_v.
    """
    kpts = np.asarray(kpts, dtype=np.float64)
    P = np.asarray(P, dtype=np.float64)
    x = kpts[:, :, 0, None]
    y = kpts[:, :, 1, None]
    w = kpts[:, :, 2, None]
    r1 = w * (y * P[None, :, 2, :] - P[None, :, 1, :])
    r2 = w * (P[None, :, 0, :] - x * P[None, :, 2, :])
    A = np.stack([r1, r2], axis=2)                # (N, V, 2, 4)
    return A.reshape(kpts.shape[0], -1, 4)


def dlt_weighted_loop(kpts, P):
    """V-view weighted DLT, one scipy SVD per point (the reference's per-point style)."""
    A = dlt_rows(kpts, P)
    out = np.empty((A.shape[0], 3))
    for n in range(A.shape[0]):
        B = A[n].T @ A[n]
        _, _, Vh = linalg.svd(B, full_matrices=False)
        out[n] = Vh[3, 0:3] / Vh[3, 3]
    return out


def dlt_weighted(kpts, P):
    """V-view weighted DLT, batched LAPACK SVD of B = A^T A (same maths as the loop)."""
    A = dlt_rows(kpts, P)
    B = np.einsum('nri,nrj->nij', A, A)
    _, _, Vh = np.linalg.svd(B)
    return Vh[:, 3, 0:3] / Vh[:, 3, 3:4]


def dlt_weighted_polished(kpts, P, iters=3):
    """The oracle's answer with LAPACK's rounding noise removed (not a reference function).

    For the mm-scale rigs of SURVEY.md section 8(d) the largest eigenvalue of B is ~1e13 times the
    smallest gap, so LAPACK's float64 eigenvector carries ~1e-9 relative error --
    the size of the fp64 tolerance itself.  This takes the batched-SVD answer and applies ``iters``
    steps of plain inverse iteration  h <- B^{-1} h  on the full 4x4 B in numpy longdouble
    (80-bit on x86; B is positive definite, convergence factor lambda_1/lambda_2 per step), i.e. it
    solves the SAME eigenproblem to ~1e-18.  Generic in h: no (X,1) parametrisation.
    """
    A = dlt_rows(kpts, P)
    B64 = np.einsum('nri,nrj->nij', A, A)
    _, _, Vh = np.linalg.svd(B64)
    h = Vh[:, 3, :].astype(np.longdouble)
    Al = _dlt_rows_longdouble(kpts, P)
    B = np.einsum('nri,nrj->nij', Al, Al)
    for _ in range(iters):
        h = _solve4(B, h)
        h = h / np.sqrt(np.einsum('ni,ni->n', h, h))[:, None]
    return (h[:, :3] / h[:, 3:4]).astype(np.float64)


def _dlt_rows_longdouble(kpts, P):
    k = np.asarray(kpts, dtype=np.float64).astype(np.longdouble)
    Pl = np.asarray(P, dtype=np.float64).astype(np.longdouble)
    x, y, w = k[:, :, 0, None], k[:, :, 1, None], k[:, :, 2, None]
    r1 = w * (y * Pl[None, :, 2, :] - Pl[None, :, 1, :])
    r2 = w * (Pl[None, :, 0, :] - x * Pl[None, :, 2, :])
    return np.stack([r1, r2], axis=2).reshape(k.shape[0], -1, 4)


def _solve4(B, r):
    """Batched 4x4 SPD solve by Gaussian elimination without pivoting, longdouble."""
    M = np.concatenate([B.copy(), r[:, :, None]], axis=2)          # (N, 4, 5)
    for i in range(4):
        M[:, i, :] = M[:, i, :] / M[:, i, i:i + 1]
        for j in range(4):
            if j != i:
                M[:, j, :] = M[:, j, :] - M[:, j, i:i + 1] * M[:, i, :]
    return M[:, :, 4]


def undistort_points(pts, cmtx, dist, iters=5):
    """cv.undistortPoints(pts, cmtx, dist, None, cmtx) restated (utils.py:1314-1315).

    OpenCV's public undistortPoints runs exactly 5 fixed-point iterations of the
    inverse Brown model (criteria = MAX_ITER 5), then maps the normalised point
    back to pixels with P = cmtx.  dist = (k1, k2, p1, p2, k3).
    pts: (N, 2) -> (N, 2) float64.
    """
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    K = np.asarray(cmtx, dtype=np.float64)
    k = np.zeros(5)
    d = np.asarray(dist, dtype=np.float64).ravel()
    k[:d.size] = d[:5]
    k1, k2, p1, p2, k3 = k
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    x = (pts[:, 0] - cx) / fx
    y = (pts[:, 1] - cy) / fy
    x0, y0 = x.copy(), y.copy()
    frozen = np.zeros(x.shape, dtype=bool)
    for _ in range(iters):
        r2 = x * x + y * y
        icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2)
        bad = (icdist < 0) & ~frozen          # OpenCV: reset to the input and stop
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        xn = (x0 - dx) * icdist
        yn = (y0 - dy) * icdist
        xn = np.where(bad, x0, xn)
        yn = np.where(bad, y0, yn)
        x = np.where(frozen, x, xn)
        y = np.where(frozen, y, yn)
        frozen |= bad
    # re-project with P = cmtx (3x3): [xx, yy, ww] = K [x, y, 1]
    xx = K[0, 0] * x + K[0, 1] * y + K[0, 2]
    yy = K[1, 0] * x + K[1, 1] * y + K[1, 2]
    ww = 1.0 / (K[2, 0] * x + K[2, 1] * y + K[2, 2])
    return np.stack([xx * ww, yy * ww], axis=1)


def triangulate_pair_svdA(P1, P2, pts1, pts2):
    """cv.triangulatePoints restated (utils.py:1324-1327): per point, the right
    singular vector of the 4x4 A = [x P[2]-P[0]; y P[2]-P[1]] (both views)."""
    pts1 = np.asarray(pts1, dtype=np.float64)
    pts2 = np.asarray(pts2, dtype=np.float64)
    A = np.stack([pts1[:, 0, None] * P1[2] - P1[0],
                  pts1[:, 1, None] * P1[2] - P1[1],
                  pts2[:, 0, None] * P2[2] - P2[0],
                  pts2[:, 1, None] * P2[2] - P2[1]], axis=1)
    _, _, Vh = np.linalg.svd(A)
    return Vh[:, 3, :]                              # homogeneous (N, 4)


def triangulate_points(kpts_2d, cmtx1, dist1, R1, T1, cmtx2, dist2, R2, T2):
    """utils.triangulate_points restated (utils.py:1277-1336).

    kpts_2d: (..., 2 cameras, 2 xy) -> (..., 3) float64: undistort each camera's
    points (:1314-1315), P = K[R|T] (:1318-1319), 2-view null-space triangulation
    (:1324-1327), de-homogenise (:1331), restore leading dims (:1335).
    """
    kpts_2d = np.asarray(kpts_2d, dtype=np.float64)
    shape = list(kpts_2d.shape[:-2])
    k = kpts_2d.reshape(-1, 2, 2)
    u1 = undistort_points(k[:, 0, :], cmtx1, dist1)
    u2 = undistort_points(k[:, 1, :], cmtx2, dist2)
    P1 = projection_matrix(cmtx1, R1, T1)
    P2 = projection_matrix(cmtx2, R2, T2)
    h = triangulate_pair_svdA(P1, P2, u1, u2)
    return (h[:, :3] / h[:, 3:4]).reshape(shape + [3])


def get_pose_3d(camera_params, all_kpts_2d, world_trans_rot=None, camera_indices=None,
                ignore_nonlinear_distortions=False):
    """pose_estimation.get_pose_3D restated (pose_estimation.py:11-65).

    camera_params: dict id -> [cmtx, R, T, dist] (utils.py:828 order).
    all_kpts_2d: sequence of T arrays (J, 3, C) = [x, y, score] per camera.
    Per (frame, joint): the two highest-score cameras among the selected ones
    (argsort(conf)[-2:], :35-37), then the 2-view triangulation (:52).  The two
    top indices address ``camera_params`` by KEY (:44-45) -- which equals "by
    position" only for keys 0..n-1, the only case the reference exercises.
    """
    keys = list(camera_params.keys())
    params = {}
    for key in keys:
        cmtx, R, T, dist = camera_params[key]
        dist = np.asarray(dist, dtype=np.float64)
        if ignore_nonlinear_distortions:
            dist = dist * 0
        params[key] = (cmtx, dist, R, T)
    if camera_indices is None:
        camera_indices = keys
    pos = [keys.index(ci) for ci in camera_indices]
    frames = []
    for kp in all_kpts_2d:
        kp = np.asarray(kp, dtype=np.float64)
        out = np.empty((kp.shape[0], 3))
        for j in range(kp.shape[0]):
            sl = kp[j][:, pos]                       # (3 or 2, n_sel)
            if sl.shape[0] == 3:
                top = np.argsort(sl[2, :])[-2:]
            else:
                top = np.array([0, 1])
            pts = sl[:2, top].T                      # (2 cams, 2 xy)
            c0 = params[top[0]]
            c1 = params[top[1]]
            out[j] = triangulate_points(pts, *(list(c0) + list(c1)))
        frames.append(out)
    frames = np.array(frames)
    if world_trans_rot is not None:
        R_W0, _ = world_trans_rot
        frames = np.einsum('ij,tpj->tpi', np.linalg.inv(R_W0), frames)
    return frames
