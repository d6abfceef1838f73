"""Oracle: learning one camera's extrinsics from sampled points (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Numpy restatement of the ``optimize_trajectory=False`` / ``extrinsic_optimization_IDs=[id]`` branch of
``pose_refinement.Optimized_3d_Pose_Estimation.sgd_optimize`` (reference pose_refinement.py:915-1096):

  * sample_gaussians        pose_refinement.py:684-706  N pixel samples per (frame, GT camera, joint)
  * construct_sample_cost   pose_refinement.py:800-831  samples triangulated by the two GT cameras
                            (utils.triangulate_points), projected with the learnt camera, scored against the
                            Gaussians of camera INDEX 2 (hard-coded upstream, :803) with the precomputed inverse
                            covariances -- which come from camera 0 (quirk Q1, :663-668)
  * the learnt parameters are the camera's full 3x3 R (upstream converts only the *initial* copy to axis-angle,
    :934, so the live matrix is optimised entry-wise) and T; Adam + clip_grad_norm_(1.0) over the 12 entries
  * smoothness / bone-length costs of the (fixed) trajectory are added to the total as constants (:984-986)

Pinned by tests/golden/extrinsic_T12.npz (runs of the unmodified reference with seeded numpy / random).
"""
import numpy as np

from . import refine as R_


def sample_gaussians(gaussians_subset, gt_ids, n_samples, rng=np.random):
    """pose_refinement.py:684-706, same loop order and the same numpy call, so a seeded global RNG reproduces upstream's
    samples exactly.  Returns (Time, J, N, 2 cameras, 2)."""
    g = np.asarray(gaussians_subset)
    T, J = g.shape[0], g.shape[2]
    out = np.empty((T, 2, J, n_samples, 2))
    for t in range(T):
        for cam in range(2):
            for j in range(J):
                mean = g[t, gt_ids[cam], j, :2]
                cov = g[t, gt_ids[cam], j, 2:].reshape(2, 2)
                out[t, cam, j] = rng.multivariate_normal(mean, cov, n_samples)
    return np.transpose(out, (0, 2, 3, 1, 4))


def sample_cost_and_grad(samples3d, K, Rm, Tv, dist, mean, Sinv, ignore_distortions=False):
    """cost = mean over finite samples of 0.5 d^T Sinv d, d = pi(R X + T) - mean[t, j]; gradient w.r.t. the 9
    entries of R and the 3 of T.  samples3d (T, J, N, 3), mean (T, J, 2), Sinv (T, J, 2, 2)."""
    X = np.asarray(samples3d, dtype=np.float64)
    cam = [K, np.asarray(Rm, dtype=np.float64).reshape(3, 3), np.asarray(Tv, dtype=np.float64).reshape(3), dist]
    flat = X.reshape(X.shape[0], -1, 3)
    pix, J = R_.project(flat, cam, ignore_distortions, jac=True)          # J = d pix / d X = Jc @ R
    pix = pix.reshape(X.shape[:3] + (2,))
    d = pix - mean[:, :, None, :]
    S = 0.5 * (Sinv + np.swapaxes(Sinv, -1, -2))[:, :, None]
    Sd = np.einsum('tjnab,tjnb->tjna', np.broadcast_to(S, d.shape[:3] + (2, 2)), d)
    q = 0.5 * np.einsum('tjna,tjna->tjn', d, np.einsum('tjab,tjnb->tjna', Sinv, d))
    ok = np.isfinite(q)
    n_ok = int(ok.sum())
    cost = q[ok].sum() / n_ok if n_ok else np.nan
    # d pix / d Xc = J @ R^-1 is awkward for a non-orthogonal R: recompute the camera-frame Jacobian directly
    Jc = _camera_jacobian(flat, cam, ignore_distortions).reshape(X.shape[:3] + (2, 3))
    gXc = np.einsum('tjnak,tjna->tjnk', Jc, Sd)
    gXc = np.where(ok[..., None] & np.isfinite(gXc), gXc, 0.0)
    dR = np.einsum('tjni,tjnk->ik', gXc, np.where(np.isfinite(X), X, 0.0)) / n_ok
    dT = gXc.sum(axis=(0, 1, 2)) / n_ok
    return cost, dR, dT, n_ok


def _camera_jacobian(X, cam, ignore_distortions):
    """d pixel / d (camera-frame point) for world points X (..., 3)."""
    K, Rm, Tv, dist = [np.asarray(a, dtype=np.float64) for a in cam]
    ident = [K, np.eye(3), np.zeros(3), dist]
    Xc = X @ Rm.reshape(3, 3).T + Tv.reshape(3)
    _, J = R_.project(Xc, ident, ignore_distortions, jac=True)
    return J


def optimize(samples3d, K, R0, T0, dist, mean, Sinv, lr=0.001, betas=(0.9, 0.999), patience=100, tolerance=1e-5,
             max_iter=1000, const_costs=None, ignore_distortions=False, dtype=np.float64, eps_adam=1e-8):
    """The reference's loop for this mode (pose_refinement.py:1002-1091).  const_costs: dict name -> value of the
    smoothness / bone-length costs of the fixed trajectory, in the order upstream lists them.  Returns dict(R, T,
    best_R, best_T, history, iterations); history[name] interleaves [cost, running mean, ...] (quirk Q5)."""
    const_costs = dict(const_costs or {})
    p = np.concatenate([np.asarray(R0, dtype=dtype).reshape(9), np.asarray(T0, dtype=dtype).reshape(3)]).astype(np.float64)
    m = np.zeros(12)
    v = np.zeros(12)
    names = ['total_cost'] + list(const_costs) + ['extrinsic_param_sample_cost']
    hist = {n: [] for n in names}
    best_cost, best, no_improve, it, step = np.inf, None, 0, 0, 0
    b1, b2 = betas
    while no_improve < patience and it <= max_iter:
        c, dR, dT, _ = sample_cost_and_grad(samples3d, K, p[:9].reshape(3, 3), p[9:], dist, mean, Sinv, ignore_distortions)
        costs = dict(const_costs)
        costs['extrinsic_param_sample_cost'] = c
        costs['total_cost'] = sum(costs.values())
        g = np.concatenate([dR.reshape(9), dT.reshape(3)])
        g = g * min(1.0, 1.0 / (np.sqrt((g * g).sum()) + 1e-6))
        step += 1
        m = m + (g - m) * (1 - b1)
        v = b2 * v + (1 - b2) * g * g
        denom = np.sqrt(v) / np.sqrt(1 - b2 ** step) + eps_adam
        p = (p - (lr / (1 - b1 ** step)) * (m / denom)).astype(dtype).astype(np.float64)
        for n in names:
            hist[n].append(float(costs[n]))
        for n in names:
            hist[n].append(float(np.mean(hist[n])))
        cur = hist['total_cost'][-1]
        if cur < best_cost - tolerance:
            best_cost, best, no_improve = cur, p.copy(), 0
        else:
            no_improve += 1
        if no_improve >= patience:
            break
        it += 1
    return {'R': p[:9].reshape(3, 3), 'T': p[9:].reshape(3, 1), 'best_R': None if best is None else best[:9].reshape(3, 3),
            'best_T': None if best is None else best[9:].reshape(3, 1), 'history': hist, 'iterations': it}


def joint_optimize(gaussians, initial_trajectory, cams, learn_ids, body_lengths, lr=0.001, betas=(0.9, 0.999), lambda_smooth=1.0,
                   lambda_body_length=1.0, patience=100, tolerance=1e-5, max_iter=1000, ignore_distortions=False,
                   time_interval=(0, -1), dtype=np.float64, eps_adam=1e-8):
    """Trajectory AND the extrinsics of the cameras in ``learn_ids`` optimised together: ``extrinsic_optimization_IDs``
    non-empty with ``optimize_trajectory=True`` (pose_refinement.py:931-961 for the parameters, :863-889 for the loss
    whose gradient now also reaches R, T, :1044-1050 one clip_grad_norm_ + Adam step over everything).  ``cams``: dict id ->
    [K, R, T, dist] with exact zeros of the learnt R, T already nudged (:937-938).  Single whole-window batch.
    Returns dict(final, best, cams, best_cams, history, iterations)."""
    t0, t1 = time_interval
    g_sub = np.asarray(gaussians, dtype=np.float64)[t0:t1]
    Sinv = R_.cov_inverse(np.asarray(gaussians), dtype=dtype).astype(np.float64)[t0:t1]
    mu0 = g_sub[:, 0, :, :2].astype(dtype).astype(np.float64)
    x = np.asarray(initial_trajectory, dtype=dtype)[t0:t1].astype(np.float64)
    bones = R_.bone_table(body_lengths)
    cams = {k: [np.asarray(a, dtype=dtype).astype(np.float64) for a in v] for k, v in cams.items()}
    ids = list(cams)
    names = ['total_cost', 'likelihood_cost'] + (['smoothness_cost'] if lambda_smooth > 0 else []) + \
            (['body_length_cost'] if lambda_body_length > 0 else [])
    hist = {n: [] for n in names}
    m_x, v_x = np.zeros_like(x), np.zeros_like(x)
    m_c = {k: np.zeros(12) for k in learn_ids}
    v_c = {k: np.zeros(12) for k in learn_ids}
    best_cost, best, best_cams, no_improve, it, step = np.inf, None, None, 0, 0, 0
    b1, b2 = betas
    while no_improve < patience and it <= max_iter:
        cam_list = [cams[k] for k in ids]
        costs, g_x = R_.total_cost_and_grad(x, mu0, Sinv, cam_list, bones, lambda_smooth, lambda_body_length, ignore_distortions)
        n_lik = R_.likelihood(x, mu0, Sinv, cam_list, ignore_distortions, grad=False)[2]
        g_c = {}
        for k in learn_ids:
            K, Rm, Tv, dist = cams[k]
            _, dR, dT, n_ok = sample_cost_and_grad(x[:, :, None, :], K, Rm, Tv, dist, mu0, Sinv, ignore_distortions)
            g_c[k] = np.concatenate([dR.reshape(9), dT.reshape(3)]) * (n_ok / n_lik)
        norm = np.sqrt((g_x * g_x).sum() + sum((g * g).sum() for g in g_c.values()))
        clip = min(1.0, 1.0 / (norm + 1e-6))
        step += 1
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step

        def adam(p, g, m, v):
            g = g * clip
            m = m + (g - m) * (1 - b1)
            v = b2 * v + (1 - b2) * g * g
            p = (p - (lr / bc1) * (m / (np.sqrt(v) / np.sqrt(bc2) + eps_adam))).astype(dtype).astype(np.float64)
            return p, m.astype(dtype).astype(np.float64), v.astype(dtype).astype(np.float64)

        x, m_x, v_x = adam(x, g_x, m_x, v_x)
        for k in learn_ids:
            p = np.concatenate([cams[k][1].reshape(9), cams[k][2].reshape(3)])
            p, m_c[k], v_c[k] = adam(p, g_c[k], m_c[k], v_c[k])
            cams[k] = [cams[k][0], p[:9].reshape(3, 3), p[9:].reshape(3, 1), cams[k][3]]
        for n in names:
            hist[n].append(float(costs[n]))
        for n in names:
            hist[n].append(float(np.mean(hist[n])))
        cur = hist['total_cost'][-1]
        if cur < best_cost - tolerance:
            best_cost, best, no_improve = cur, x.copy(), 0
            best_cams = {k: [a.copy() for a in v] for k, v in cams.items()}
        else:
            no_improve += 1
        if no_improve >= patience:
            break
        it += 1
    return {'final': x, 'best': best, 'cams': cams, 'best_cams': best_cams, 'history': hist, 'iterations': it}
