"""Drop-in for the 3D step of the reference's ``pose_estimation`` module (pose_estimation.py:11-65).

The 2D driver (mmdet/mmpose inference, OpenCV preview) is out of scope; this
module keeps ``get_pose_3D`` call-compatible and runs it as ONE kernel launch over all frames and
joints instead of one OpenCV call chain per joint.
"""
import numpy as np

from . import triangulation as _tri
from . import utils


def get_pose_3D(camera_params, all_kpts_2d, world_trans_rot=None, camera_indices=None,
                ignore_nonlinear_distortions=False):
    """Triangulate every (frame, joint) from its two most confident cameras.

    camera_params  dict id -> [cmtx, R, T, dist]            (utils.get_params_from_name order)
    all_kpts_2d    sequence of T arrays (J, 3, C) [x, y, score] (or (J, 2, C): cameras 0 and 1)
    camera_indices ids of the cameras (columns) to consider; default all
    Returns (T, J, 3) float64.

    Upstream behaviour kept on purpose: the two winning positions within the selected columns
    address ``camera_params`` BY KEY (pose_estimation.py:44-45), i.e. slot q always uses
    ``camera_params[q]``; with ``camera_indices`` other than 0..n-1 that pairs pixels of one camera
    with the parameters of another, exactly as the reference does.
    """
    keys = list(camera_params.keys())
    if camera_indices is None:
        camera_indices = keys
    positions = [keys.index(ci) for ci in camera_indices]
    n_sel = len(positions)
    if n_sel < 2:
        raise ValueError('need at least two cameras to triangulate')
    slots = []
    for q in range(n_sel):
        cmtx, R, T, dist = camera_params[q]          # KeyError if the ids are not 0..n-1, as upstream
        dist = np.asarray(dist, dtype=np.float64).reshape(-1)[:5]
        slots.append((np.asarray(cmtx, dtype=np.float64), np.asarray(R, dtype=np.float64),
                      np.asarray(T, dtype=np.float64), dist * 0 if ignore_nonlinear_distortions else dist))
    P = np.stack([utils.get_projection_matrix(c, R, T) for c, R, T, _ in slots])
    K = np.stack([c for c, _, _, _ in slots])
    D = np.stack([d for _, _, _, d in slots])

    kp = np.asarray(all_kpts_2d, dtype=np.float64)           # (T, J, 3|2, C)
    if kp.ndim != 4:
        raise ValueError(f'all_kpts_2d must stack to (T, J, 3, C), got {kp.shape}')
    kp = kp[:, :, :, positions]
    if kp.shape[2] == 2:                                     # no scores: cameras 0 and 1 (pose_estimation.py:39)
        score = np.zeros(kp.shape[:2] + (1, n_sel))
        score[..., 0] = 1.0
        score[..., 1] = 2.0
        kp = np.concatenate([kp, score], axis=2)
    frames = _tri.triangulate_multiview(np.ascontiguousarray(kp), P, K=K, dist=D, layout='n3v', mode='top2')
    if world_trans_rot is not None:
        R_W0, _ = world_trans_rot
        frames = np.einsum('ij,tpj->tpi', np.linalg.inv(np.asarray(R_W0, dtype=np.float64)), frames)
    return frames
