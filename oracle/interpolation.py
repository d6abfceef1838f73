"""Oracle: pose_refinement.linear_interpolation (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates reference pose_refinement.py:15-84: per (joint, dim, t) a window of ~k frames centred on t (clipped at the
ends), outlier rejection by mean +- k_std*std AND median +- median_std*MAD, then a least-squares line through the
surviving samples evaluated at t (or their mean with use_rolling_average).  Quirk kept: with fewer than two
surviving samples the output stays 0 (the upstream `continue` skips the assignment, :62-64).
Pinned by tests/golden/interp.npz (outputs of the unmodified reference).
"""
import numpy as np


def linear_interpolation(points, k=5, k_std=2, median_std=2, use_rolling_average=False, filter_distance_from_median=True):
    points = np.array(points, dtype=np.float64)
    squeeze = points.ndim == 2
    p3 = points[:, :, None] if squeeze else points
    T, P, D = p3.shape
    out = np.zeros_like(p3)
    for p in range(P):
        for d in range(D):
            for t in range(T):
                lo, hi = max(0, t - k // 2), min(T, t + k // 2 + 1)
                w = p3[lo:hi, p, d]
                mean, std = np.mean(w), np.std(w)
                med = np.median(w)
                mad = np.median(np.abs(w - med))
                valid = np.abs(w - mean) <= k_std * std
                if filter_distance_from_median:
                    valid &= np.abs(w - med) <= median_std * mad
                vals = w[valid]
                if len(vals) < 2:
                    continue
                if use_rolling_average:
                    out[t, p, d] = np.mean(vals)
                else:
                    times = np.arange(lo, hi)[valid].astype(np.float64)
                    tm, vm = times.mean(), vals.mean()
                    sxx = ((times - tm) ** 2).sum()
                    slope = ((times - tm) * (vals - vm)).sum() / sxx
                    out[t, p, d] = vm + slope * (t - tm)
    return out[:, :, 0] if squeeze else out
