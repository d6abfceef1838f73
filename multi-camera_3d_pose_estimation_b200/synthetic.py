"""Seeded synthetic rigs and inputs; shared by tests, bench and golden generation.

numpy only -- no CUDA, no oracle.  Units are millimetres and pixels.  The
rig conventions are the reference's: a camera is ``[cmtx(3,3), R(3,3), T(3,1),
dist(1,5)]`` (utils.py:828) and P = K [R|T] (utils.py:433-435).
"""
import numpy as np

DEFAULT_K = np.array([[1000.0, 0.0, 640.0], [0.0, 1000.0, 360.0], [0.0, 0.0, 1.0]])
DEFAULT_DIST = np.array([[0.01, -0.002, 5e-4, -3e-4, 1e-4]])


def _look_at(cam_pos, target=np.zeros(3), up=np.array([0.0, 1.0, 0.0])):
    """World->camera rotation with +z towards ``target``; T = -R @ cam_pos."""
    z = target - cam_pos
    z = z / np.linalg.norm(z)
    x = np.cross(up, z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z])
    T = -R @ cam_pos
    return R, T.reshape(3, 1)


SCENE_CENTRE = np.array([0.0, 0.0, 3000.0])


def ring_rig(n_cams, radius=3000.0, height=400.0, distortion=False, centre=SCENE_CENTRE):
    """``n_cams`` cameras on a ring of ``radius`` mm around ``centre``, all looking at it.  With the
    default centre (0,0,3000) the world origin lies on the ring near camera 0, as in a rig whose
    world frame is its first camera."""
    cams = {}
    centre = np.asarray(centre, dtype=np.float64)
    for v in range(n_cams):
        ang = 2.0 * np.pi * v / n_cams + 0.1
        pos = centre + np.array([radius * np.sin(ang), height * np.cos(3 * ang), -radius * np.cos(ang)])
        R, T = _look_at(pos, target=centre)
        K = DEFAULT_K.copy()
        K[0, 0] += 7.0 * v          # cameras are not identical
        K[1, 1] += 5.0 * v
        K[0, 2] += 3.0 * v
        K[1, 2] -= 2.0 * v
        dist = DEFAULT_DIST * (1.0 + 0.1 * v) if distortion else np.zeros((1, 5))
        cams[v] = [K, R, T, dist]
    return cams


def stereo_rig(distortion=True):
    """Config 1: camera 0 at the origin, camera 1 rotated about y and translated."""
    ang = np.deg2rad(-25.0)
    R1 = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    T1 = np.array([[-1200.0], [15.0], [300.0]])
    d0 = DEFAULT_DIST.copy() if distortion else np.zeros((1, 5))
    d1 = DEFAULT_DIST * 1.3 if distortion else np.zeros((1, 5))
    K1 = DEFAULT_K.copy()
    K1[0, 0], K1[1, 1], K1[0, 2], K1[1, 2] = 1012.0, 1009.0, 633.0, 371.0
    return {0: [DEFAULT_K.copy(), np.eye(3), np.zeros((3, 1)), d0],
            1: [K1, R1, T1, d1]}


def projection_matrices(cams):
    """(V, 3, 4) float64 stack of K [R|T] in key order."""
    return np.stack([c[0] @ np.hstack([c[1], np.asarray(c[2]).reshape(3, 1)]) for c in cams.values()])


def project(X, cam, distort=True):
    """Pinhole + 5-coefficient Brown projection of (..., 3) points (pose_refinement.py:134-174 maths)."""
    K, R, T, dist = cam
    Xc = X @ np.asarray(R).T + np.asarray(T).reshape(1, 3)
    x = Xc[..., 0] / Xc[..., 2]
    y = Xc[..., 1] / Xc[..., 2]
    if distort:
        k1, k2, p1, p2, k3 = np.asarray(dist).ravel()
        r2 = x * x + y * y
        rad = 1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3
        xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        x, y = xd, yd
    u = K[0, 0] * x + K[0, 1] * y + K[0, 2]
    v = K[1, 1] * y + K[1, 2]
    return np.stack([u, v], axis=-1)


def smooth_trajectory(n_frames, n_joints, rng, centre=(0.0, 0.0, 0.0), spread=400.0):
    """(T, J, 3) joints around ``centre``: a fixed skeleton offset plus a slow common drift."""
    t = np.arange(n_frames)[:, None, None]
    base = rng.normal(0.0, spread, size=(1, n_joints, 3))
    phase = rng.uniform(0, 2 * np.pi, size=(1, 1, 3))
    drift = 150.0 * np.sin(2 * np.pi * t / 97.0 + phase) + 60.0 * np.sin(2 * np.pi * t / 31.0 + 2 * phase)
    wobble = 8.0 * np.sin(2 * np.pi * t / 13.0 + rng.uniform(0, 2 * np.pi, size=(1, n_joints, 3)))
    return np.asarray(centre).reshape(1, 1, 3) + base + drift + wobble


def keypoints_from_trajectory(X, cams, rng, noise_px=1.0, distort=True):
    """Reference layout kpts_2d (T, J, 3, C) = [x, y, score] (pose_estimation.py:135)."""
    T_, J = X.shape[:2]
    C = len(cams)
    kp = np.empty((T_, J, 3, C))
    for c, cam in enumerate(cams.values()):
        uv = project(X, cam, distort=distort) + rng.normal(0.0, noise_px, size=(T_, J, 2))
        kp[:, :, 0, c] = uv[..., 0]
        kp[:, :, 1, c] = uv[..., 1]
    kp[:, :, 2, :] = rng.uniform(0.2, 1.0, size=(T_, J, C))
    return kp


def multiview_points(n_points, n_cams, seed=0, noise_px=1.0, dtype=np.float64):
    """Configs 2/5: (N, V, 3) [x, y, w] on a ring rig without distortion, plus P (V,3,4) and truth."""
    rng = np.random.default_rng(seed)
    cams = ring_rig(n_cams)
    X = SCENE_CENTRE + rng.normal(0.0, 400.0, size=(n_points, 3))
    kp = np.empty((n_points, n_cams, 3))
    for c, cam in enumerate(cams.values()):
        kp[:, c, :2] = project(X, cam, distort=False) + rng.normal(0.0, noise_px, size=(n_points, 2))
    kp[:, :, 2] = rng.uniform(0.2, 1.0, size=(n_points, n_cams))
    return kp.astype(dtype), projection_matrices(cams), X, cams


def gaussian_blob_heatmaps(n_maps, H=64, W=48, seed=0, sigma=2.0, noise=0.005, dtype=np.float32):
    """Config 3 heatmaps: one Gaussian blob per map at U(8,W-8) x U(8,H-8) plus N(0, noise)."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(8, W - 8, size=(n_maps, 1, 1))
    cy = rng.uniform(8, H - 8, size=(n_maps, 1, 1))
    amp = rng.uniform(0.3, 1.0, size=(n_maps, 1, 1))
    yy, xx = np.mgrid[0:H, 0:W]
    hm = amp * np.exp(-((xx[None] - cx) ** 2 + (yy[None] - cy) ** 2) / (2 * sigma * sigma))
    hm = hm + rng.normal(0.0, noise, size=hm.shape)
    return hm.astype(dtype), np.concatenate([cx.reshape(-1, 1), cy.reshape(-1, 1)], axis=1)


# COCO-17 skeleton used to make bone lengths meaningful for the refinement inputs.
COCO_BONES = [(0, 1), (0, 2), (1, 3), (2, 4), (5, 7), (7, 9), (6, 8), (8, 10), (11, 13), (13, 15),
              (12, 14), (14, 16), (5, 6), (5, 11), (6, 12), (11, 12)]

EXAMPLE_BODY_LENGTHS = {            # examples/body_part_lengths.yaml:1-13 (values are the fixture's)
    'left_shoulder_left_elbow': 38, 'left_elbow_left_wrist': 27,
    'right_shoulder_right_elbow': 38, 'right_elbow_right_wrist': 27,
    'left_hip_left_knee': 51, 'left_knee_left_ankle': 40,
    'right_hip_right_knee': 51, 'right_knee_right_ankle': 40,
    'left_hip_right_hip': 31, 'left_shoulder_left_hip': 54,
    'right_shoulder_right_hip': 54, 'left_shoulder_right_shoulder': 47}


def coco_skeleton_trajectory(n_frames, rng, centre=(0.0, 0.0, 3000.0), scale=10.0):
    """(T, 17, 3) articulated-ish COCO pose (mm) whose bone lengths are ~ scale x the yaml values."""
    rest = np.array([
        [0, -60, 0], [3, -63, 0], [-3, -63, 0], [7, -61, 0], [-7, -61, 0],      # head
        [23.5, -45, 0], [-23.5, -45, 0], [30, -8, 0], [-30, -8, 0],           # shoulders, elbows
        [32, 19, 0], [-32, 19, 0], [15.5, 9, 0], [-15.5, 9, 0],               # wrists, hips
        [17, 60, 0], [-17, 60, 0], [18, 100, 0], [-18, 100, 0]], dtype=np.float64) * scale
    t = np.arange(n_frames)[:, None, None]
    sway = np.concatenate([60.0 * np.sin(2 * np.pi * t / 120.0), 15.0 * np.sin(2 * np.pi * t / 45.0),
                           100.0 * np.sin(2 * np.pi * t / 200.0)], axis=2)
    limb = 12.0 * np.sin(2 * np.pi * t / 40.0 + rng.uniform(0, 2 * np.pi, size=(1, 17, 3)))
    return np.asarray(centre).reshape(1, 1, 3) + rest[None] + sway + limb


def refinement_inputs(n_frames, n_cams=2, seed=0, dtype=np.float64):
    """Config 4 inputs: (gaussians (T,C,17,6), initial trajectory (T,17,3), cams, truth).

    gaussians[t,c,j] = [mean_x, mean_y, var_x, cov, cov, var_y] with mean = projection + N(0,2),
    covariance [[4,.5],[.5,3]]; initial trajectory = truth + N(0,3 mm)  (SURVEY.md §8(d) row 4).
    """
    rng = np.random.default_rng(seed)
    cams = stereo_rig(distortion=True) if n_cams == 2 else ring_rig(n_cams, distortion=True)
    centre = (0.0, 0.0, 3000.0)
    truth = coco_skeleton_trajectory(n_frames, rng, centre=centre)
    g = np.empty((n_frames, n_cams, 17, 6))
    for c, cam in enumerate(cams.values()):
        g[:, c, :, :2] = project(truth, cam) + rng.normal(0.0, 2.0, size=(n_frames, 17, 2))
    g[..., 2:] = np.array([4.0, 0.5, 0.5, 3.0])
    init = truth + rng.normal(0.0, 3.0, size=truth.shape)
    return g.astype(dtype), init.astype(dtype), cams, truth
