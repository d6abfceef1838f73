"""Drop-in for the hot-path part of the reference's ``utils`` module.

Same names, argument meaning and return types as the reference (file:line cited per function); the
arithmetic runs in libmc3d.so on the GPU.  Only what the triangulation / refinement path and its CLIs
use is here; calibration, capture and GUI helpers are out of scope.
"""
import inspect
import os

import numpy as np
import yaml

from . import triangulation as _tri

# COCO-17 bones and joint names (reference utils.py:1068-1074 `CONNECTIVITY_DICT['coco']`,
# utils.py:1077-1161 `POINT_INFO['coco']`; only the fields the path reads are kept).
CONNECTIVITY_DICT = {
    'coco': [(0, 1), (0, 2), (1, 3), (2, 4), (5, 7), (7, 9), (6, 8), (8, 10), (11, 13), (13, 15), (12, 14), (14, 16),
             (5, 6), (5, 11), (6, 12), (11, 12)],
}
_COCO_NAMES = ['nose', 'left_eye', 'right_eye', 'left_ear', 'right_ear', 'left_shoulder', 'right_shoulder',
               'left_elbow', 'right_elbow', 'left_wrist', 'right_wrist', 'left_hip', 'right_hip', 'left_knee',
               'right_knee', 'left_ankle', 'right_ankle']
POINT_INFO = {'coco': {i: {'name': name, 'id': i} for i, name in enumerate(_COCO_NAMES)}}


def to_numpy(arr):
    """utils.py:1272-1273."""
    return arr if isinstance(arr, np.ndarray) else arr.numpy()


# ---- projection matrices (utils.py:425-435, 803-805) ---------------------------------------------
def _make_homogeneous_rep_matrix(R, t):
    H = np.eye(4)
    H[:3, :3] = R
    H[:3, 3] = np.asarray(t).reshape(3)
    return H


def get_projection_matrix(cmtx, R, T):
    return np.asarray(cmtx) @ _make_homogeneous_rep_matrix(R, T)[:3, :]


def calculate_projection_matrix(cmtx, rvec, tvec):
    return get_projection_matrix(cmtx, rvec, tvec)


# ---- triangulation -----------------------------------------------------------------------------------
def DLT(P1, P2, point1, point2):
    """Two-view DLT of one point (reference utils.py:19-34): smallest right singular vector of A^T A,
    de-homogenised.  Runs on the GPU through the 2-view weighted kernel with w = 1."""
    kp = np.array([[[point1[0], point1[1], 1.0], [point2[0], point2[1], 1.0]]], dtype=np.float64)
    P = np.stack([np.asarray(P1, dtype=np.float64), np.asarray(P2, dtype=np.float64)])
    return _tri.triangulate_multiview(kp, P)[0]


def triangulate_points(kpts_2d, cmtx1, dist1, R1, T1, cmtx2, dist2, R2, T2):
    """Reference utils.py:1277-1336: undistort both cameras' pixels (cv.undistortPoints with P=cmtx),
    build P = K[R|T], 2-view null-space triangulation, de-homogenise.

    kpts_2d (..., 2 cameras, 2 xy) ndarray or torch tensor -> new float64 ndarray (..., 3).
    """
    kpts_2d, cmtx1, dist1, R1, T1, cmtx2, dist2, R2, T2 = [
        np.asarray(to_numpy(a), dtype=np.float64) for a in (kpts_2d, cmtx1, dist1, R1, T1, cmtx2, dist2, R2, T2)]
    lead = list(kpts_2d.shape[:-2])
    pts = kpts_2d.reshape(-1, 2, 2)
    kp = np.concatenate([pts, np.ones((pts.shape[0], 2, 1))], axis=2)
    P = np.stack([get_projection_matrix(cmtx1, R1, T1), get_projection_matrix(cmtx2, R2, T2)])
    X = _tri.triangulate_multiview(kp, P, K=np.stack([cmtx1, cmtx2]),
                                   dist=np.stack([dist1.reshape(-1)[:5], dist2.reshape(-1)[:5]]))
    return X.reshape(lead + [3])


# ---- reprojection (utils.py:438-458, 558-567) ---------------------------------------------------------------------
def project_points(points_3d, K, R, T, dist_coeffs=None):
    """Project (N,3) or (Time,N,3) points to pixels with intrinsics K, extrinsics R (3,3), T and optional Brown
    distortion (k1,k2,p1,p2,k3) -- the cv.projectPoints wrapper of reference utils.py:438-458.  float64 in and out;
    runs on the GPU (csrc/refine.cu projection kernel)."""
    import torch
    from . import _lib
    pts = np.asarray(points_3d, dtype=np.float64)
    shape = pts.shape
    flat = np.ascontiguousarray(pts.reshape(-1, 3))
    d = np.zeros(5)
    if dist_coeffs is not None:
        dd = np.asarray(dist_coeffs, dtype=np.float64).reshape(-1)
        d[:min(5, dd.size)] = dd[:5]
    row = np.ascontiguousarray(np.concatenate([np.asarray(K, dtype=np.float64).reshape(9), np.asarray(R, dtype=np.float64).reshape(9),
                                               np.asarray(T, dtype=np.float64).reshape(3), d]))
    if not torch.cuda.is_available():
        raise _lib.Mc3dError('project_points needs a CUDA device (no CPU fallback)')
    dev = torch.device('cuda', torch.cuda.current_device())
    p = torch.as_tensor(flat).to(dev)
    out = torch.empty((flat.shape[0], 2), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mc3d_project_points_f64(p.data_ptr(), flat.shape[0], row.ctypes.data, 0, out.data_ptr(),
                                                      torch.cuda.current_stream().cuda_stream))
    res = out.cpu().numpy()
    return res.reshape(shape[:2] + (2,)) if len(shape) == 3 else res.reshape(-1, 2)


def compute_2d_coordinates(P, point_3d):
    """Undistorted projection of one 3D point with a 3x4 projection matrix (utils.py:558-567); host arithmetic on
    four numbers, kept for the plotting callers."""
    uv = np.asarray(P, dtype=np.float64) @ np.array([point_3d[0], point_3d[1], point_3d[2], 1.0])
    return np.array([uv[0], uv[1]]) / uv[2]


# ---- camera parameter files (formats of utils.py:750-793) -----------------------------------------
def read_camera_parameters(camera_name, params_dir=''):
    """`<name>.dat`: a header line, 3 rows of the intrinsic matrix, a header line, one row of
    distortion coefficients (utils.py:750-770).  Returns (cmtx (3,3), dist (1,5))."""
    with open(os.path.join(params_dir or os.getcwd(), camera_name + '.dat')) as fh:
        lines = fh.read().splitlines()
    cmtx = [[float(tok) for tok in lines[i].split()] for i in (1, 2, 3)]
    dist = [[float(tok) for tok in lines[5].split()]]
    return np.array(cmtx), np.array(dist)


def read_rotation_translation(camera_name, params_dir=''):
    """`rot_trans_<name>.dat`: header, 3 rows R, header, 3 rows T (utils.py:772-793)."""
    with open(os.path.join(params_dir or os.getcwd(), 'rot_trans_' + camera_name + '.dat')) as fh:
        lines = fh.read().splitlines()
    rot = [[float(tok) for tok in lines[i].split()] for i in (1, 2, 3)]
    trans = [[float(tok) for tok in lines[i].split()] for i in (5, 6, 7)]
    return np.array(rot), np.array(trans)


def get_params_from_name(camera_name, intrinsic_params_dir='', extrinsic_params_dir=''):
    """utils.py:807-828: returns (P, [cmtx, R, T, dist]); a part that fails to load is None (and
    a message is printed), exactly as upstream."""
    intrinsic_params_dir = intrinsic_params_dir or os.path.join(os.getcwd(), 'intrinsic_camera_parameters')
    extrinsic_params_dir = extrinsic_params_dir or os.path.join(os.getcwd(), 'extrinsic_camera_parameters')
    cmtx = dist = rvec = tvec = P = None
    try:
        cmtx, dist = read_camera_parameters(camera_name, params_dir=intrinsic_params_dir)
    except Exception:
        print(f'failed to load {camera_name} intrinsic params')
    try:
        rvec, tvec = read_rotation_translation(camera_name, params_dir=extrinsic_params_dir)
    except Exception:
        print(f'failed to load {camera_name} extrinsic params')
    try:
        P = calculate_projection_matrix(cmtx, rvec, tvec)
    except Exception:
        print(f'failed to compute {camera_name} projection')
    return P, [cmtx, rvec, tvec, dist]


# ---- skeleton helpers (utils.py:1175-1208) -----------------------------------------------------------
def generate_connectivity_names(connectivity_list, point_names):
    return {i: f"{point_names[s]['name']}_{point_names[e]['name']}" for i, (s, e) in enumerate(connectivity_list)}


def get_body_part_vects(pose, connectivity_type='coco'):
    bones = CONNECTIVITY_DICT[connectivity_type]
    names = generate_connectivity_names(bones, POINT_INFO[connectivity_type])
    return {names[i]: pose[:, e, :] - pose[:, s, :] for i, (s, e) in enumerate(bones)}


def get_body_part_lengths(pose, connectivity_type='coco'):
    """Per-frame bone lengths, torch in -> torch out, numpy in -> numpy out (utils.py:1197-1208).
    Reporting helper (pose_refinement.py:1239-1247); the optimiser computes bones inside its kernels."""
    import torch
    vects = get_body_part_vects(pose, connectivity_type)
    return {k: (torch.norm(v, dim=1) if isinstance(v, torch.Tensor) else np.linalg.norm(v, axis=1))
            for k, v in vects.items()}


def rotation_conversion(rotation_rep, to_vector=True):
    """Axis-angle <-> rotation matrix (Rodrigues), reference utils.py:1219-1268: a (3,3) input with
    to_vector=False (and a 3-vector with to_vector=True) passes through unchanged."""
    import torch
    was_numpy = isinstance(rotation_rep, np.ndarray)
    rep = torch.as_tensor(rotation_rep)
    is_matrix = tuple(rep.shape) == (3, 3)
    if is_matrix and to_vector:
        theta = torch.acos((torch.trace(rep) - 1) / 2)
        if torch.abs(theta) < 1e-6:
            return torch.zeros(3)
        axis = torch.stack([rep[2, 1] - rep[1, 2], rep[0, 2] - rep[2, 0], rep[1, 0] - rep[0, 1]]) / (2 * torch.sin(theta))
        res = theta * axis
    elif not is_matrix and not to_vector:
        theta = torch.norm(rep)
        if torch.abs(theta) < 1e-6:
            return torch.eye(3)
        ux, uy, uz = (rep / theta).reshape(3)
        Kx = torch.stack([torch.stack([torch.zeros_like(ux), -uz, uy]), torch.stack([uz, torch.zeros_like(ux), -ux]),
                          torch.stack([-uy, ux, torch.zeros_like(ux)])])
        res = torch.eye(3, dtype=rep.dtype) + torch.sin(theta) * Kx + (1 - torch.cos(theta)) * (Kx @ Kx)
    else:
        res = rep
    return np.array(res) if was_numpy else res


# ---- artefact / config helpers (utils.py:1365-1399) ---------------------------------------------------
def load_if_exists(path):
    if os.path.exists(path):
        return np.load(path)
    print(f'file does not exist at path {path}')
    return None


def load_config(config_path=None):
    if config_path is None:
        return {}
    with open(config_path) as fh:
        return yaml.safe_load(fh)


def get_function_defaults(func):
    return {k: p.default for k, p in inspect.signature(func).parameters.items()
            if p.default is not inspect.Parameter.empty}


def prepare_kwargs(func, user_kwargs):
    """Signature defaults overlaid with the yaml section; '.inf' -> inf, betas list -> tuple."""
    kwargs = dict(get_function_defaults(func))
    kwargs.update(user_kwargs or {})
    for k, v in list(kwargs.items()):
        if isinstance(v, str) and v == '.inf':
            kwargs[k] = np.inf
        if k == 'betas' and isinstance(v, list):
            kwargs[k] = tuple(v)
    return kwargs
