"""CPU: the refinement oracle against runs of the UNMODIFIED reference (tests/golden/refine_T48.npz) and its
hand-derived gradient against torch autograd of an independently written loss."""
import numpy as np
import pytest

from conftest import cams_from_golden, load_golden
from oracle import refine as R

RUNS = {
    'readme': dict(lr=0.01, lambda_smooth=1e-6, lambda_body_length=1, patience=100, max_iter=60, time_interval=[0, 40]),
    'defaults': dict(max_iter=25),
    'stop': dict(lr=0.01, lambda_smooth=1e-6, lambda_body_length=1, patience=3, tolerance=0.7, max_iter=200,
                 time_interval=[0, 48]),
    'nodist': dict(lr=0.005, lambda_smooth=0.5, lambda_body_length=0, max_iter=20, ignore_distortions=True,
                   time_interval=[4, 44]),
    'batch': dict(lr=0.01, lambda_smooth=1e-3, lambda_body_length=1, max_iter=12, batch_size=16, time_interval=[1, 47]),
}


def test_projection_matches_reference():
    g = load_golden('refine_T48.npz')
    cams = cams_from_golden(g, 2)
    for i in range(2):
        assert np.abs(R.project(g['init'], cams[i]) - g[f'proj_f64_cam{i}']).max() < 1e-10
        assert np.abs(R.project(g['init'], cams[i], True) - g[f'proj_nodist_f64_cam{i}']).max() < 1e-10
        assert np.abs(R.project(g['init'], cams[i]) - g[f'proj_f32_cam{i}']).max() < 1e-3      # reference in float32


@pytest.mark.parametrize('run', sorted(RUNS))
@pytest.mark.parametrize('tag', ['f64', 'f32'])
def test_optimisation_history_matches_reference(syn, run, tag):
    g = load_golden('refine_T48.npz')
    cams = list(cams_from_golden(g, 2).values())
    out = R.sgd_optimize(g['gaussians'], g['init'], cams, syn.EXAMPLE_BODY_LENGTHS,
                         dtype=np.float64 if tag == 'f64' else np.float32, **RUNS[run])
    rtol, atol = (1e-12, 1e-9) if tag == 'f64' else (1e-5, 1e-3)
    key = f'run_{run}_{tag}'
    for name, hist in out['history'].items():
        ref = g[f'{key}_{name}']
        assert len(hist) == len(ref)                        # same number of iterations, incl. early stopping
        assert np.max(np.abs(np.array(hist) - ref) / np.abs(ref)) < rtol
    assert np.abs(out['best'] - g[f'{key}_best']).max() < atol
    assert np.abs(out['final'] - g[f'{key}_final']).max() < atol


def test_optimisation_history_matches_reference_at_4000_frames(syn):
    """tests/golden/refine_T4000.npz (the unmodified reference on 4 000 frames; inputs from the seeded generator)."""
    g = load_golden('refine_T4000.npz')
    n, stride = int(g['n_frames']), int(g['stride'])
    gs, init, cams, _ = syn.refinement_inputs(n, n_cams=2, seed=int(g['seed']))
    out = R.sgd_optimize(gs, init, list(cams.values()), syn.EXAMPLE_BODY_LENGTHS, dtype=np.float64, lr=0.01, lambda_smooth=1e-6,
                         lambda_body_length=1, patience=100, max_iter=11, time_interval=[0, n])
    for name, hist in out['history'].items():
        ref = g[f'f64_{name}']
        assert len(hist) == len(ref) == 24             # 12 iterations, each followed by the running mean (Q5)
        assert np.max(np.abs(np.array(hist) - ref) / np.abs(ref)) < 1e-12
    assert np.abs(out['final'][::stride] - g['f64_final']).max() < 1e-9
    assert np.abs(out['best'][::stride] - g['f64_best']).max() < 1e-9
    assert abs(np.abs(out['final']).sum() - float(g['f64_final_abs_sum'])) / float(g['f64_final_abs_sum']) < 1e-12


def test_quirks_are_reproduced(syn):
    g = load_golden('refine_T48.npz')
    # Q3: the default time_interval [0, -1] drops the last frame; Q4: max_iter + 1 iterations; Q5: interleaved means
    assert g['run_defaults_f64_final'].shape[0] == 47
    assert len(g['run_defaults_f64_total_cost']) == 2 * (25 + 1)
    h = g['run_defaults_f64_total_cost']
    assert np.isclose(h[1], h[0]) and np.isclose(h[3], np.mean(h[:3]))
    # the reference dies on a NaN joint after one iteration (documented in make_golden.py)
    assert 'KeyError' in str(g['run_nan_f64_raised'])


def test_nan_masked_forward_costs_match_reference_iteration0(syn):
    g = load_golden('refine_T48.npz')
    cams = list(cams_from_golden(g, 2).values())
    x = g['init_nan'][0:40]
    gs = g['gaussians'][0:40]
    Sinv = R.cov_inverse(g['gaussians'])[0:40]
    lik, _, _ = R.likelihood(x, gs[:, 0, :, :2], Sinv, cams)
    sm, _, _ = R.smoothness(x, 1e-6)
    assert np.isclose(lik, float(g['run_nan_f64_iter0_likelihood_cost']), rtol=1e-12)
    assert np.isclose(sm, float(g['run_nan_f64_iter0_smoothness_cost']), rtol=1e-12)


def test_closed_form_gradient_matches_autograd(syn):
    import torch
    gs, init, cams, _ = syn.refinement_inputs(12, n_cams=3, seed=3)
    cams = list(cams.values())
    bones = R.bone_table(syn.EXAMPLE_BODY_LENGTHS)
    Sinv = R.cov_inverse(gs)
    mu0 = gs[:, 0, :, :2]
    costs, grad = R.total_cost_and_grad(init, mu0, Sinv, cams, bones, 0.37, 1.3)

    x = torch.tensor(init, dtype=torch.float64, requires_grad=True)
    lik = []
    for K, Rm, T, dist in cams:
        Xc = x @ torch.tensor(Rm).T + torch.tensor(T).reshape(1, 1, 3)
        a, b = Xc[..., 0] / Xc[..., 2], Xc[..., 1] / Xc[..., 2]
        k1, k2, p1, p2, k3 = [float(v) for v in np.asarray(dist).ravel()]
        r2 = a * a + b * b
        rad = 1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3
        xd = a * rad + 2 * p1 * a * b + p2 * (r2 + 2 * a * a)
        yd = b * rad + p1 * (r2 + 2 * b * b) + 2 * p2 * a * b
        pix = torch.stack([K[0, 0] * xd + K[0, 1] * yd + K[0, 2], K[1, 1] * yd + K[1, 2]], dim=-1)
        d = pix - torch.tensor(mu0)
        lik.append(0.5 * torch.einsum('tji,tjik,tjk->tj', d, torch.tensor(Sinv), d))
    L = torch.stack(lik).mean()
    D = x[2:] - 2 * x[1:-1] + x[:-2]
    Ls = 0.37 * (D ** 2).sum(dim=(1, 2)).mean()
    a_vec = torch.tensor([b[2] for b in bones]).repeat(12, 1)
    b_vec = torch.stack([torch.norm(x[:, e] - x[:, s], dim=1) for s, e, _ in bones], dim=1)
    mu = (a_vec * b_vec).sum() / (b_vec * b_vec).sum()
    Lb = 1.3 * ((a_vec - mu * b_vec) ** 2).sum() / (a_vec ** 2).sum()
    total = L + Ls + Lb
    total.backward()
    assert np.isclose(costs['total_cost'], float(total), rtol=1e-12)
    assert np.abs(grad - x.grad.numpy()).max() < 1e-12 * max(1.0, np.abs(grad).max())


PERCAM_RUNS = {'a': dict(lr=0.01, lambda_smooth=1e-3, lambda_body_length=1, max_iter=30, time_interval=[0, 32]),
               'b': dict(lr=0.005, lambda_smooth=0.5, lambda_body_length=0, max_iter=15, time_interval=[2, 30], ignore_distortions=True)}


@pytest.mark.parametrize('run', sorted(PERCAM_RUNS))
@pytest.mark.parametrize('tag', ['f64', 'f32'])
def test_per_camera_gaussians_match_the_reference_loop(syn, run, tag):
    """The opt-in form: tests/golden/refine_percam_T32.npz is the reference's own sgd_optimize with
    its likelihood cost re-indexed by `camera_index` as Trajectory_Optimization does (pose_refinement.py:499); three cameras
    so that camera 0's Gaussian differs from the others'."""
    g = load_golden('refine_percam_T32.npz')
    cams = list(cams_from_golden(g, 3).values())
    out = R.sgd_optimize(g['gaussians'], g['init'], cams, syn.EXAMPLE_BODY_LENGTHS, per_camera_gaussians=True,
                         dtype=np.float64 if tag == 'f64' else np.float32, **PERCAM_RUNS[run])
    rtol, atol = (1e-12, 1e-9) if tag == 'f64' else (1e-5, 1e-3)
    key = f'run_{run}_{tag}'
    for name, hist in out['history'].items():
        ref = g[f'{key}_{name}']
        assert len(hist) == len(ref)
        assert np.max(np.abs(np.array(hist) - ref) / np.abs(ref)) < rtol, name
    assert np.abs(out['best'] - g[f'{key}_best']).max() < atol
    # and it is NOT what upstream's default computes: camera-0 Gaussians for every camera give another loss
    quirk = R.sgd_optimize(g['gaussians'], g['init'], cams, syn.EXAMPLE_BODY_LENGTHS, dtype=np.float64, **dict(PERCAM_RUNS[run], max_iter=0))
    assert abs(quirk['history']['likelihood_cost'][0] - g[f'run_{run}_f64_likelihood_cost'][0]) > 1e-3 * abs(g[f'run_{run}_f64_likelihood_cost'][0])
