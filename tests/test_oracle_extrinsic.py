"""CPU: oracle/extrinsic.py against runs of the unmodified reference (tests/golden/extrinsic_T12.npz)."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import extrinsic as E
from oracle import refine as R

GOLD = os.path.join(ROOT, 'tests', 'golden', 'extrinsic_T12.npz')


def _cams(g):
    return {c: [g[f'cam{c}_K'], g[f'cam{c}_R'], g[f'cam{c}_T'], g[f'cam{c}_dist']] for c in range(3)}


def _kw(g, key):
    out = {}
    for item in g[f'{key}_kw']:
        k, v = str(item).split('=')
        out[k] = float(v) if ('.' in v or 'e' in v) else int(v)
    return out


def test_sampling_reproduces_the_reference_draws():
    g = np.load(GOLD)
    np.random.seed(3)
    s = E.sample_gaussians(g['gaussians'][0:12], [0, 1], 6)
    assert s.shape == (12, 17, 6, 2, 2)
    assert np.array_equal(s, g['f64_plain_samples'])


@pytest.mark.parametrize('key', ['f64_plain', 'f64_consts', 'f64_stop', 'f32_plain', 'f32_consts', 'f32_stop'])
def test_optimisation_matches_reference_runs(key):
    g = np.load(GOLD)
    cams = _cams(g)
    kw = _kw(g, key)
    dt = np.float64 if key.startswith('f64') else np.float32
    gs = g['gaussians'].astype(dt).astype(np.float64)
    mean = gs[:, 2, :, :2]                                           # camera index 2, hard-coded upstream (:803)
    Sinv = R.cov_inverse(g['gaussians'], dtype=dt).astype(np.float64)         # camera 0's covariances (Q1)
    lam_s, lam_b = kw.pop('lambda_smooth'), kw.pop('lambda_body_length')
    consts = {}
    x = g['initial'].astype(dt).astype(np.float64)
    if lam_s > 0:
        consts['smoothness_cost'] = R.smoothness(x, lam_s, grad=False)[0]
    if lam_b > 0:
        consts['body_length_cost'] = R.body_length(x, R.bone_table(_lengths()), lam_b, grad=False)[0]
    K, R0, T0, dist = [np.asarray(a, dtype=dt).astype(np.float64) for a in cams[2]]
    import random
    random.seed(3)                                                   # the golden runs seeded random with 3
    R0[R0 == 0] = dt(random.random() / 10 ** 6)                      # exact zeros are nudged upstream (:937-938)
    T0[T0 == 0] = dt(random.random() / 10 ** 6)
    res = E.optimize(g[f'{key}_samples3d'], K, R0, T0, dist, mean, Sinv, lr=1e-3, const_costs=consts, dtype=dt, **kw)
    rtol = 1e-9 if dt == np.float64 else 2e-5
    for name in ['total_cost', 'extrinsic_param_sample_cost'] + list(consts):
        ref = g[f'{key}_hist_{name}']
        got = np.array(res['history'][name])
        assert len(got) == len(ref), (name, len(got), len(ref))
        assert np.allclose(got, ref, rtol=rtol, atol=0), (name, np.max(np.abs(got - ref) / np.abs(ref)))
    atol = 1e-9 if dt == np.float64 else 1e-4
    assert np.allclose(res['R'], g[f'{key}_R'], atol=atol) and np.allclose(res['T'], g[f'{key}_T'], atol=atol * 1e3)
    assert np.allclose(res['best_R'], g[f'{key}_best_R'], atol=atol)


def _lengths():
    import importlib.util
    spec = importlib.util.spec_from_file_location('syn', os.path.join(ROOT, 'multi-camera_3d_pose_estimation_b200', 'synthetic.py'))
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    return dict(syn.EXAMPLE_BODY_LENGTHS)


@pytest.mark.parametrize('key', ['f64_joint', 'f32_joint'])
def test_joint_camera_and_trajectory_optimisation_matches_reference_runs(key):
    """extrinsic_optimization_IDs=[2] with optimize_trajectory=True (pose_refinement.py:931-961)."""
    import random
    g = np.load(GOLD)
    dt = np.float64 if key.startswith('f64') else np.float32
    cams = {k: [np.asarray(a, dtype=dt).astype(np.float64) for a in v] for k, v in _cams(g).items()}
    random.seed(3)
    cams[2][1][cams[2][1] == 0] = dt(random.random() / 10 ** 6)
    cams[2][2][cams[2][2] == 0] = dt(random.random() / 10 ** 6)
    res = E.joint_optimize(g['gaussians'], g['initial'], cams, [2], _lengths(), lr=1e-3, lambda_smooth=1e-3,
                           lambda_body_length=1.0, max_iter=14, time_interval=(0, 12), dtype=dt)
    rtol = 1e-9 if dt == np.float64 else 2e-5
    for name, got in res['history'].items():
        ref = g[f'{key}_hist_{name}']
        assert len(got) == len(ref) == 30 and np.allclose(got, ref, rtol=rtol, atol=0), name
    atol = 1e-9 if dt == np.float64 else 1e-4
    assert np.abs(res['final'] - g[f'{key}_traj']).max() < atol * 10
    assert np.abs(res['cams'][2][1] - g[f'{key}_R']).max() < atol and np.abs(res['cams'][2][2] - g[f'{key}_T']).max() < atol * 1e3
    assert np.abs(res['best_cams'][2][1] - g[f'{key}_best_R']).max() < atol


def test_camera_subset_in_the_likelihood_matches_reference_run():
    """camera_IDs=[0, 2] out of three cameras (pose_refinement.py:866): pins the oracle path the GPU test
    test_sgd_optimize_matches_oracle[subset_of_cameras] compares against."""
    g = np.load(GOLD)
    cams = _cams(g)
    out = R.sgd_optimize(g['gaussians'], g['initial'], [cams[0], cams[2]], _lengths(), lr=0.01, lambda_smooth=1e-3,
                         lambda_body_length=1.0, max_iter=8, time_interval=(0, 12))
    for name, hist in out['history'].items():
        ref = g[f'f64_subset_hist_{name}']
        assert len(hist) == len(ref) and np.allclose(hist, ref, rtol=1e-12, atol=0), name
    assert np.abs(out['final'] - g['f64_subset_traj']).max() < 1e-9
