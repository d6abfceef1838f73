"""CPU: the algebra behind the two-phase refinement step (csrc/refine.cu costgrad_loop / mix_of).

The gradient of the reference's loss is linear in three scalars that need global sums,
    g = alpha g1 + sigma gs + beta (G2 - mu G3),  alpha = 1/N_lik, sigma = 2 lambda_s/N_s, beta = -2 lambda_b mu/(a.a),
so one pass can store the four component vectors (with G2' = G2 - mu_prev G3) and their ten dot products, and after one
reduction every thread knows the gradient and |g|^2.  Checked against the oracle's closed-form gradient, which is itself
pinned to the unmodified reference's runs (tests/test_oracle_refine.py)."""
import numpy as np
import pytest

from oracle import refine as R


def _components(x, mu0, Sinv, cams, bones, mu_prev):
    """Unscaled components exactly as pass 1 stores them, from the oracle's scaled pieces."""
    c_lik, g_lik, n_lik = R.likelihood(x, mu0, Sinv, cams)
    g1 = g_lik * n_lik                                               # likelihood() returns the gradient / N_lik
    c_s, g_s1, n_s = R.smoothness(x, 1.0)
    gs = g_s1 * n_s / 2.0                                             # smoothness(lam=1) returns 2/N_s * stencil
    s_idx, e_idx = [b[0] for b in bones], [b[1] for b in bones]
    a = np.array([b[2] for b in bones])
    vec = x[:, e_idx, :] - x[:, s_idx, :]
    b = np.sqrt((vec * vec).sum(axis=2))
    u = vec / b[..., None]
    G2 = np.zeros_like(x)
    G3 = np.zeros_like(x)
    for k in range(len(bones)):
        G2[:, e_idx[k]] += a[k] * u[:, k]
        G2[:, s_idx[k]] -= a[k] * u[:, k]
        G3[:, e_idx[k]] += vec[:, k]
        G3[:, s_idx[k]] -= vec[:, k]
    sums = dict(n_lik=n_lik, n_s=n_s, ab=(a[None] * b).sum(), bb=(b * b).sum(), aa=(a * a).sum() * x.shape[0])
    return g1, gs, G2 - mu_prev * G3, G3, sums


@pytest.mark.parametrize('mu_prev', [0.0, 'last_step'])
def test_gradient_is_linear_in_the_global_scalars(mu_prev):
    rng = np.random.default_rng(4)
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location('syn', os.path.join(ROOT, 'multi-camera_3d_pose_estimation_b200', 'synthetic.py'))
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    gs_, init, cams, _ = syn.refinement_inputs(30, n_cams=2, seed=8)
    x = init + rng.normal(0, 2.0, size=init.shape)
    lam_s, lam_b = 0.3, 1.7
    bones = R.bone_table(syn.EXAMPLE_BODY_LENGTHS)
    Sinv = R.cov_inverse(gs_)
    mu0 = gs_[:, 0, :, :2]
    cam_list = list(cams.values())
    _, g_ref = R.total_cost_and_grad(x, mu0, Sinv, cam_list, bones, lam_s, lam_b)

    if mu_prev == 'last_step':                                        # mu moves by ~1e-3 relative per step at most
        mu_prev = R.body_length(x, bones, lam_b, grad=False)[2] * (1.0 + 1e-3)
    g1, gs, G2p, G3, S = _components(x, mu0, Sinv, cam_list, bones, mu_prev)
    mu = S['ab'] / S['bb']
    alpha, sigma = 1.0 / S['n_lik'], 2.0 * lam_s / S['n_s']
    beta = -2.0 * lam_b * mu / S['aa']
    gamma = beta * (mu_prev - mu)                                     # coefficient of G3 once G2' absorbed mu_prev G3
    g = alpha * g1 + sigma * gs + beta * G2p + gamma * G3
    assert np.abs(g - g_ref).max() < 1e-12 * np.abs(g_ref).max()

    # |g|^2 as the quadratic form pass 2 evaluates from the ten dot products
    comps = [g1, gs, G2p, G3]
    coef = [alpha, sigma, beta, gamma]
    dots = {(i, j): float((comps[i] * comps[j]).sum()) for i in range(4) for j in range(i, 4)}
    gn2 = sum(coef[i] * coef[j] * dots[(i, j)] * (1 if i == j else 2) for i in range(4) for j in range(i, 4))
    assert np.isclose(gn2, (g_ref * g_ref).sum(), rtol=1e-10)
    if mu_prev:
        # with last step's mu the stored bone component is the small, already cancelled combination
        assert np.abs(G2p).max() < 0.2 * np.abs(G2p + mu_prev * G3).max()
