"""Drop-in for the decode half of the reference's ``mmpose_pose_estimation.PoseEstimator``.

Only the heatmap -> moments decode (mmpose_pose_estimation.py:114-215) is on the hot path; building and
running the mmdet / mmpose networks (``__init__``, ``predict``) is third-party DNN inference and out of
scope, so those raise NotImplementedError here.
"""
import numpy as np

from . import decode as _decode


class PoseEstimator:
    def __init__(self, det_config, det_checkpoint, pose_config, pose_checkpoint, device='cpu', det_cat_id=0,
                 bbox_thr=0.3, nms_thr=0.3, using_detector=True):
        # upstream's signature (mmpose_pose_estimation.py:82); building the mmdet / mmpose networks is not on the path
        raise NotImplementedError('mc3d_b200 accelerates the decode only; construct the mmpose model with the '
                                  "reference's PoseEstimator and call PoseEstimator.get_heatmap_means_cov from here")

    def predict(self, input_file, return_full_heatmaps=False):
        raise NotImplementedError('network inference is out of scope of mc3d_b200')

    @staticmethod
    def _is_torch(x):
        return type(x).__module__.startswith('torch')

    def get_heatmap_means_cov(self, heatmaps):
        """(J, H, W) heatmaps -> (J, 6) float64 [mean_x, mean_y, var_x, cov_xy, cov_xy, var_y] in heatmap pixels.

        As upstream (mmpose_pose_estimation.py:163-215): values < 0.01 are zeroed IN PLACE in the caller's array
        first (:166), all-zero maps give six zeros, a list recurses (:174-178).  Callable unbound with
        ``self=None`` as the reference's own tests of it would.
        """
        if isinstance(heatmaps, list):
            return np.array([PoseEstimator.get_heatmap_means_cov(self, h) for h in heatmaps])
        if PoseEstimator._is_torch(heatmaps) and heatmaps.is_cuda:
            _, mom = _decode.decode_heatmaps(heatmaps, want_kpts=False, write_back=heatmaps.is_contiguous())
            if not heatmaps.is_contiguous():
                heatmaps[heatmaps < 0.01] = 0
            return mom.cpu().numpy()
        heatmaps[heatmaps < 0.01] = 0                       # the upstream side effect, on the caller's array
        hm = heatmaps.numpy() if PoseEstimator._is_torch(heatmaps) else np.asarray(heatmaps)
        _, mom = _decode.decode_heatmaps(np.ascontiguousarray(hm, dtype=np.float32), want_kpts=False)
        return mom

    @staticmethod
    def get_heatmap_means_stds(heatmaps):
        """Means and per-axis standard deviations without thresholding (mmpose_pose_estimation.py:114-161).
        Returns (means, stds) as lists of (x, y) tuples; a list input recurses."""
        if isinstance(heatmaps, list):
            res = [PoseEstimator.get_heatmap_means_stds(h) for h in heatmaps]
            return [r[0] for r in res], [r[1] for r in res]
        if PoseEstimator._is_torch(heatmaps):
            hm = heatmaps if heatmaps.is_cuda else heatmaps.numpy()
        else:
            hm = np.asarray(heatmaps)
        if isinstance(hm, np.ndarray):
            hm = np.ascontiguousarray(hm, dtype=np.float32)
        _, mom = _decode.decode_heatmaps(hm, threshold=-np.inf, want_kpts=False)
        mom = mom.cpu().numpy() if not isinstance(mom, np.ndarray) else mom
        stds = np.sqrt(np.maximum(mom[:, [2, 5]], 0.0))
        return [tuple(m) for m in mom[:, :2].tolist()], [tuple(s) for s in stds.tolist()]
