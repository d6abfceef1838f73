"""Oracle: trajectory refinement (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Closed-form numpy restatement of what ``pose_refinement.Optimized_3d_Pose_Estimation.sgd_optimize``
(reference pose_refinement.py:894-1096) computes with torch autograd on its default path
(``optimize_trajectory=True``, no extrinsic learning, no NN):

  * project_points_torch            pose_refinement.py:94-179   -> ``project`` (+ Jacobian)
  * Sigma^-1 precompute             pose_refinement.py:663-668  -> ``cov_inverse`` (camera-0 quirk Q1)
  * compute_likelihood_cost         pose_refinement.py:863-889  -> ``likelihood``
  * compute_smoothness_cost         pose_refinement.py:836-845  -> ``smoothness``
  * compute_body_length_cost        pose_refinement.py:848-860, utils.py:1185-1208 -> ``body_length``
  * nan_mean                        pose_refinement.py:221-229  (finite-masked means)
  * clip_grad_norm_(max_norm=1) + torch.optim.Adam step + running-mean early stopping
                                    pose_refinement.py:1002-1091 -> ``sgd_optimize``

The hand-derived gradients are checked against torch autograd of the same loss in
tests/test_oracle_refine.py, and the cost histories / trajectories against runs of the unmodified
reference (tests/golden/refine_T48.npz).  Arithmetic is float64 throughout (``dtype`` only rounds
the stored state, to mimic the reference's float32 default).
"""
import numpy as np

COCO_BONES = [(0, 1), (0, 2), (1, 3), (2, 4), (5, 7), (7, 9), (6, 8), (8, 10), (11, 13), (13, 15), (12, 14),
              (14, 16), (5, 6), (5, 11), (6, 12), (11, 12)]                   # utils.py:1070
COCO_NAMES = ['nose', 'left_eye', 'right_eye', 'left_ear', 'right_ear', 'left_shoulder', 'right_shoulder',
              'left_elbow', 'right_elbow', 'left_wrist', 'right_wrist', 'left_hip', 'right_hip', 'left_knee',
              'right_knee', 'left_ankle', 'right_ankle']                      # utils.py:1077-1161


def bone_table(body_lengths):
    """[(start, end, target_length)] for the yaml's bones, in yaml key order (pose_refinement.py:851)."""
    names = {f'{COCO_NAMES[s]}_{COCO_NAMES[e]}': (s, e) for s, e in COCO_BONES}       # utils.py:1175-1181
    return [(names[k][0], names[k][1], float(v)) for k, v in body_lengths.items()]    # KeyError as upstream


def project(X, cam, ignore_distortions=False, jac=False):
    """pose_refinement.py:134-174.  X (..., 3) -> pixels (..., 2) [and d pixel / d X (..., 2, 3)]."""
    K, R, T, dist = [np.asarray(a, dtype=np.float64) for a in cam]
    R = R.reshape(3, 3)
    T = T.reshape(3)
    Xc = X @ R.T + T
    z = Xc[..., 2]
    a = Xc[..., 0] / z
    b = Xc[..., 1] / z
    if ignore_distortions:
        xd, yd = a, b
    else:
        k1, k2, p1, p2, k3 = dist.reshape(-1)[:5]
        r2 = a * a + b * b
        rad = 1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3
        xd = a * rad + 2 * p1 * a * b + p2 * (r2 + 2 * a * a)
        yd = b * rad + p1 * (r2 + 2 * b * b) + 2 * p2 * a * b
    u = K[0, 0] * xd + K[0, 1] * yd + K[0, 2]
    v = K[1, 0] * xd + K[1, 1] * yd + K[1, 2]
    s = K[2, 0] * xd + K[2, 1] * yd + K[2, 2]
    pix = np.stack([u / s, v / s], axis=-1)
    if not jac:
        return pix
    # d pix / d (xd, yd)
    Jp = np.empty(X.shape[:-1] + (2, 2))
    Jp[..., 0, 0] = (K[0, 0] - pix[..., 0] * K[2, 0]) / s
    Jp[..., 0, 1] = (K[0, 1] - pix[..., 0] * K[2, 1]) / s
    Jp[..., 1, 0] = (K[1, 0] - pix[..., 1] * K[2, 0]) / s
    Jp[..., 1, 1] = (K[1, 1] - pix[..., 1] * K[2, 1]) / s
    # d (xd, yd) / d (a, b)
    Jd = np.zeros(X.shape[:-1] + (2, 2))
    if ignore_distortions:
        Jd[..., 0, 0] = 1.0
        Jd[..., 1, 1] = 1.0
    else:
        drad = k1 + 2 * k2 * r2 + 3 * k3 * r2 ** 2
        Jd[..., 0, 0] = rad + 2 * a * a * drad + 2 * p1 * b + 6 * p2 * a
        Jd[..., 0, 1] = 2 * a * b * drad + 2 * p1 * a + 2 * p2 * b
        Jd[..., 1, 0] = Jd[..., 0, 1]
        Jd[..., 1, 1] = rad + 2 * b * b * drad + 6 * p1 * b + 2 * p2 * a
    # d (a, b) / d Xc
    Jn = np.zeros(X.shape[:-1] + (2, 3))
    Jn[..., 0, 0] = 1 / z
    Jn[..., 0, 2] = -a / z
    Jn[..., 1, 1] = 1 / z
    Jn[..., 1, 2] = -b / z
    J = Jp @ Jd @ Jn @ R
    return pix, J


def cov_inverse(gaussians, eps=1e-6, dtype=np.float64, camera=0):
    """(T, J, 2, 2) inverse of camera 0's covariances + eps I (pose_refinement.py:663-668; quirk Q1:
    ``gaussians[:, 0]`` regardless of camera), computed in ``dtype`` like the reference.  ``camera`` selects another
    camera's covariances for the per-camera form (``per_camera_gaussians`` of sgd_optimize below)."""
    g = np.asarray(gaussians, dtype=dtype)
    cov = g[:, camera, :, 2:].reshape(g.shape[0], g.shape[2], 2, 2) + dtype(eps) * np.eye(2, dtype=dtype)
    return np.linalg.inv(cov).astype(dtype)


def _finite(a):
    return np.isfinite(a)


def likelihood(x, mu0, Sinv, cams, ignore_distortions=False, grad=True):
    """mean over finite (camera, frame, joint) entries of 0.5 d^T Sinv d, d = pi_c(x) - mu0
    (pose_refinement.py:866-889; camera-0 means for every camera, Q1).  Returns (cost, grad or None, count)."""
    total, count = 0.0, 0
    g = np.zeros_like(x) if grad else None
    parts = []
    per_camera = isinstance(mu0, (list, tuple))            # per-camera Gaussians: one (mu, Sinv) per camera
    for ci, cam in enumerate(cams):
        pix, J = project(x, cam, ignore_distortions, jac=True)
        mu_c, S_c = (mu0[ci], Sinv[ci]) if per_camera else (mu0, Sinv)
        d = pix - mu_c
        Sd = np.einsum('tjab,tjb->tja', S_c, d)
        q = 0.5 * np.einsum('tja,tja->tj', d, Sd)
        ok = _finite(q)
        parts.append((ok, q, d, J, S_c))
        total += q[ok].sum()
        count += int(ok.sum())
    cost = total / count if count else np.nan
    if grad:
        for ok, q, d, J, S_c in parts:
            Ssym_d = 0.5 * (np.einsum('tjab,tjb->tja', S_c, d) + np.einsum('tjba,tjb->tja', S_c, d))
            gi = np.einsum('tjak,tja->tjk', J, Ssym_d) / count
            gi[~ok] = 0.0                                   # masked entries carry no gradient (see module note)
            g += np.where(np.isfinite(gi), gi, 0.0)
    return cost, g, count


def smoothness(x, lam, grad=True):
    """lam * mean over finite t>=2 of ||x_t - 2 x_{t-1} + x_{t-2}||_F^2 (pose_refinement.py:836-845)."""
    if x.shape[0] < 3:
        return np.nan, (np.zeros_like(x) if grad else None), 0
    D = x[2:] - 2 * x[1:-1] + x[:-2]
    terms = (D * D).sum(axis=(1, 2))
    ok = _finite(terms)
    n = int(ok.sum())
    cost = lam * terms[ok].sum() / n if n else np.nan
    g = None
    if grad:
        g = np.zeros_like(x)
        Dm = np.where(ok[:, None, None], D, 0.0)
        g[2:] += Dm
        g[1:-1] += -2 * Dm
        g[:-2] += Dm
        g *= 2 * lam / n if n else np.nan
    return cost, g, n


def body_length(x, bones, lam, grad=True):
    """lam * ||a - mu b||^2 / ||a||^2 with mu = a.b / b.b over all (bone, frame) lengths
    (pose_refinement.py:848-860).  Non-finite lengths are left out of a.b, b.b and the residual (upstream
    has no defined behaviour there: it prints 'nan cost' and dies); ||a||^2 always counts every bone."""
    L = x.shape[0]
    s_idx = [b[0] for b in bones]
    e_idx = [b[1] for b in bones]
    a = np.array([b[2] for b in bones])[None, :].repeat(L, axis=0)        # (L, B)
    vec = x[:, e_idx, :] - x[:, s_idx, :]
    b = np.sqrt((vec * vec).sum(axis=2))
    ok = _finite(b)
    aa = (a * a).sum()
    ab = (a * b)[ok].sum()
    bb = (b * b)[ok].sum()
    mu = ab / bb
    res = np.where(ok, a - mu * b, 0.0)
    cost = lam * (res * res).sum() / aa
    g = None
    if grad:
        g = np.zeros_like(x)
        coef = np.where(ok, -2 * lam * mu * (a - mu * b) / aa, 0.0)
        with np.errstate(divide='ignore', invalid='ignore'):
            u = np.where((b > 0)[..., None] & ok[..., None], vec / b[..., None], 0.0)   # torch.norm subgradient 0 at 0
        contrib = coef[..., None] * u
        for k in range(len(bones)):
            g[:, e_idx[k], :] += contrib[:, k, :]
            g[:, s_idx[k], :] -= contrib[:, k, :]
    return cost, g, mu


def total_cost_and_grad(x, mu0, Sinv, cams, bones, lam_s, lam_b, ignore_distortions=False):
    costs = {}
    c, g, _ = likelihood(x, mu0, Sinv, cams, ignore_distortions)
    costs['likelihood_cost'] = c
    if lam_s > 0:
        cs, gs, _ = smoothness(x, lam_s)
        costs['smoothness_cost'] = cs
        g = g + gs
    if lam_b > 0:
        cb, gb, _ = body_length(x, bones, lam_b)
        costs['body_length_cost'] = cb
        g = g + gb
    costs['total_cost'] = sum(costs.values())
    return costs, g


def sgd_optimize(gaussians, initial_trajectory, cams, body_lengths, lr=0.001, betas=(0.9, 0.999), lambda_smooth=1.0,
                 lambda_body_length=1.0, patience=100, tolerance=1e-5, max_iter=1000, batch_size=None,
                 ignore_distortions=False, time_interval=(0, -1), dtype=np.float64, eps_adam=1e-8,
                 per_camera_gaussians=False, gaussian_cameras=None):
    """The reference's optimisation loop (pose_refinement.py:894-1096, default path).

    Returns dict(best, final, history) where history[name] is the reference's interleaved list
    [cost_batch..., running_mean, ...] (quirk Q5: the running mean is over the list that already contains the
    previous running means).  Quirks kept: time_interval slicing incl. the default [0,-1] dropping the last
    frame (Q3), max_iter + 1 iterations (Q4), half-overlapping batch windows stepped sequentially.

    ``per_camera_gaussians=True`` is the opt-in the survey asks for: camera c's projection is
    compared with camera c's OWN Gaussian -- ``gaussians[:, gaussian_cameras[c]]``, default c -- as the superseded
    ``Trajectory_Optimization`` indexes them (pose_refinement.py:499), instead of camera 0's for every camera.
    """
    t0, t1 = time_interval
    g_sub = np.asarray(gaussians, dtype=np.float64)[t0:t1]
    Sinv_all = cov_inverse(np.asarray(gaussians), dtype=dtype).astype(np.float64)[t0:t1]
    Time = len(g_sub)
    bs = Time if batch_size is None else batch_size
    Time = int(np.floor(Time / bs) * bs)
    x = np.asarray(initial_trajectory, dtype=dtype)[t0:t1].astype(np.float64)
    bones = bone_table(body_lengths)
    windows = [(s, s + bs) for s in range(0, Time - bs + 1, bs // 2)]
    mu0_all = g_sub[:, 0, :, :2].astype(dtype).astype(np.float64)
    if per_camera_gaussians:
        gcs = list(range(len(cams))) if gaussian_cameras is None else list(gaussian_cameras)
        mu_pc = [g_sub[:, c, :, :2].astype(dtype).astype(np.float64) for c in gcs]
        S_pc = [cov_inverse(np.asarray(gaussians), dtype=dtype, camera=c).astype(np.float64)[t0:t1] for c in gcs]
    names = ['total_cost', 'likelihood_cost'] + (['smoothness_cost'] if lambda_smooth > 0 else []) + \
            (['body_length_cost'] if lambda_body_length > 0 else [])
    hist = {n: [] for n in names}
    m = np.zeros_like(x)
    v = np.zeros_like(x)
    step = 0
    best_cost = np.inf
    best = None
    no_improve = 0
    it = 0
    b1, b2 = betas
    while no_improve < patience and it <= max_iter:
        for (f0, f1) in windows:
            xs = x[f0:f1]
            if per_camera_gaussians:
                costs, gs = total_cost_and_grad(xs, [a[f0:f1] for a in mu_pc], [a[f0:f1] for a in S_pc], cams, bones,
                                                lambda_smooth, lambda_body_length, ignore_distortions)
            else:
                costs, gs = total_cost_and_grad(xs, mu0_all[f0:f1], Sinv_all[f0:f1], cams, bones, lambda_smooth,
                                                lambda_body_length, ignore_distortions)
            g = np.zeros_like(x)
            g[f0:f1] = gs
            norm = np.sqrt((g * g).sum())
            g = g * min(1.0, 1.0 / (norm + 1e-6))
            step += 1
            m = m + (g - m) * (1 - b1)
            v = b2 * v + (1 - b2) * g * g
            bc1 = 1 - b1 ** step
            bc2 = 1 - b2 ** step
            denom = np.sqrt(v) / np.sqrt(bc2) + eps_adam
            x = (x - (lr / bc1) * (m / denom)).astype(dtype).astype(np.float64)
            m = m.astype(dtype).astype(np.float64)
            v = v.astype(dtype).astype(np.float64)
            for n in names:
                hist[n].append(float(costs[n]))
        for n in names:
            hist[n].append(float(np.mean(hist[n])))
        cur = hist['total_cost'][-1]
        if cur < best_cost - tolerance:
            best_cost = cur
            best = x.copy()
            no_improve = 0
        else:
            no_improve += 1
        if no_improve >= patience:
            break
        it += 1
    return {'best': best, 'final': x, 'history': hist, 'iterations': it}
