"""Oracle: heatmap -> keypoint / Gaussian-moment decode (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates in numpy:
  * PoseEstimator.get_heatmap_means_cov   (reference mmpose_pose_estimation.py:163-215)
  * PoseEstimator.get_heatmap_means_stds  (reference mmpose_pose_estimation.py:114-161)
    -- pinned by tests/golden/heatmap_moments.npz (outputs of the unmodified reference).
  * the argmax / quarter-pixel keypoint decode that the reference obtains from mmpose
    (call sites mmpose_pose_estimation.py:253-259; consumed pose_estimation.py:104-105).

PARITY UNPINNED for the argmax decode only: mmpose is a third-party dependency that upstream neither
vendors nor pins (README.md:18, examples/model_paths.yaml:3-18) and it is not installed here, so no
reference output exists to pin it.  ``argmax_decode`` restates the published algorithm of mmpose 1.x's
MSRA heatmap codec (``get_heatmap_maximum`` + the 0.25-pixel shift towards the larger neighbour of
``MSRAHeatmap.decode``): flat argmax (first maximum), score = the maximum, locations with score <= 0
become (-1, -1), and for 1 < px < W-1, 1 < py < H-1 the keypoint moves 0.25 px along the sign of the
central difference.  Coordinates are heatmap pixels; the caller applies the bbox affine.
The argmax half (integer pixel, score, the (-1, -1) rule, first-of-ties) IS pinned by a third-party port of
mmpose's ``_get_max_preds``: tests/golden/argmax_vitpose.npz (HF transformers' ViTPose
``get_keypoint_predictions``); only the quarter-pixel shift rests on the restatement alone.
"""
import numpy as np


def heatmap_means_cov(heatmaps, threshold=0.01, mutate=True):
    """get_heatmap_means_cov restated (mmpose_pose_estimation.py:163-215).

    heatmaps (J, H, W) float32 -> (J, 6) float64 [mean_x, mean_y, var_x, cov_xy, cov_xy, var_y].
    Values < ``threshold`` are zeroed IN PLACE first (:166, quirk Q7); all-zero maps give six zeros
    (:191-193); arithmetic is float32 with float32 index grids (:181-182), results widened to float64.
    """
    hm = heatmaps if mutate else heatmaps.copy()
    hm[hm < threshold] = 0
    J, H, W = hm.shape
    y_grid = np.arange(H, dtype=np.float32).reshape(H, 1) * np.ones((1, W), dtype=np.float32)
    x_grid = np.ones((H, 1), dtype=np.float32) * np.arange(W, dtype=np.float32).reshape(1, W)
    out = np.zeros((J, 6), dtype=np.float64)
    for j in range(J):
        h = hm[j].astype(np.float32, copy=False)
        s = h.sum(dtype=np.float32)
        if s == 0:
            continue
        p = h / s
        mx = (x_grid * p).sum(dtype=np.float32)
        my = (y_grid * p).sum(dtype=np.float32)
        vx = ((x_grid - mx) ** 2 * p).sum(dtype=np.float32)
        vy = ((y_grid - my) ** 2 * p).sum(dtype=np.float32)
        cxy = ((x_grid - mx) * (y_grid - my) * p).sum(dtype=np.float32)
        out[j] = [mx, my, vx, cxy, cxy, vy]
    return out


def heatmap_means_cov_f64(heatmaps, threshold=0.01):
    """The same moments in float64 arithmetic (what the float32 results approximate); no mutation."""
    hm = np.asarray(heatmaps, dtype=np.float32).copy()
    hm[hm < threshold] = 0
    hm = hm.astype(np.float64)
    J, H, W = hm.shape
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    s = hm.sum(axis=(1, 2))
    ok = s > 0
    p = hm / np.where(ok, s, 1.0)[:, None, None]
    mx = (xx * p).sum(axis=(1, 2))
    my = (yy * p).sum(axis=(1, 2))
    dx = xx[None] - mx[:, None, None]
    dy = yy[None] - my[:, None, None]
    vx = (dx * dx * p).sum(axis=(1, 2))
    vy = (dy * dy * p).sum(axis=(1, 2))
    cxy = (dx * dy * p).sum(axis=(1, 2))
    out = np.stack([mx, my, vx, cxy, cxy, vy], axis=1)
    out[~ok] = 0.0
    return out


def heatmap_means_stds(heatmaps):
    """get_heatmap_means_stds restated (mmpose_pose_estimation.py:114-161): means and sqrt of the diagonal
    variances, no thresholding inside.  Returns (means (J,2), stds (J,2))."""
    m = heatmap_means_cov(np.array(heatmaps, dtype=np.float32), threshold=-np.inf, mutate=False)
    return m[:, :2], np.sqrt(m[:, [2, 5]])


def argmax_decode(heatmaps):
    """mmpose 1.x MSRA-style decode (see module docstring; PARITY UNPINNED).

    heatmaps (J, H, W) -> keypoints (J, 2) float32 [x, y] in heatmap pixels, scores (J,) float32.
    """
    hm = np.asarray(heatmaps, dtype=np.float32)
    J, H, W = hm.shape
    flat = hm.reshape(J, -1)
    idx = np.argmax(flat, axis=1)
    scores = flat[np.arange(J), idx].astype(np.float32)
    px = (idx % W).astype(np.int64)
    py = (idx // W).astype(np.int64)
    kp = np.stack([px, py], axis=1).astype(np.float32)
    for j in range(J):
        if not scores[j] > 0:
            kp[j] = -1.0
            continue
        x, y = px[j], py[j]
        if 1 < x < W - 1 and 1 < y < H - 1:
            dx = hm[j, y, x + 1] - hm[j, y, x - 1]
            dy = hm[j, y + 1, x] - hm[j, y - 1, x]
            kp[j, 0] += np.float32(np.sign(dx) * 0.25)
            kp[j, 1] += np.float32(np.sign(dy) * 0.25)
    return kp, scores
