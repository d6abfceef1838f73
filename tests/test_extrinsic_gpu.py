"""GPU: learning one camera's extrinsics from sampled points (csrc/extrinsic.cu behind
Optimized_3d_Pose_Estimation.sgd_optimize(extrinsic_optimization_IDs=[2], optimize_trajectory=False)) against runs of the
unmodified reference (tests/golden/extrinsic_T12.npz) and against the oracle."""
import os
import random

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, 'tests', 'golden', 'extrinsic_T12.npz')


def _kw(g, key):
    out = {}
    for item in g[f'{key}_kw']:
        k, v = str(item).split('=')
        out[k] = float(v) if ('.' in v or 'e' in v) else int(v)
    return out


@pytest.mark.parametrize('key', ['f64_plain', 'f64_consts', 'f64_stop', 'f32_plain', 'f32_consts', 'f32_stop'])
def test_extrinsics_from_samples_match_reference_runs(key, capsys):
    import torch
    import mc3d_b200.pose_refinement as pr
    import mc3d_b200.synthetic as syn
    g = np.load(GOLD)
    cams = {c: [g[f'cam{c}_K'], g[f'cam{c}_R'], g[f'cam{c}_T'], g[f'cam{c}_dist']] for c in range(3)}
    dt = torch.float64 if key.startswith('f64') else torch.float32
    np.random.seed(3)
    random.seed(3)
    opt = pr.Optimized_3d_Pose_Estimation(g['gaussians'].copy(), g['initial'].copy(),
                                          decomposed_cam_params_initial={i: list(cams[i]) for i in cams},
                                          body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), N_sample_points=6, torch_dtype=dt)
    opt.sgd_optimize(extrinsic_optimization_IDs=[2], optimize_trajectory=False, GT_camera_IDs=[0, 1], lr=1e-3, print_frequency=10,
                     time_interval=[0, 12], **_kw(g, key))
    assert np.array_equal(opt.samples, g[f'{key}_samples'])            # same draws as upstream
    s3 = opt.samples_3d.numpy().astype(np.float64)
    tol3 = 1e-9 if dt == torch.float64 else 1e-3
    assert np.abs(s3 - g[f'{key}_samples3d']).max() < tol3 * (1 if dt == torch.float32 else 5000)
    rtol = 1e-9 if dt == torch.float64 else 1e-4                         # north-star tolerance on the loss
    for name, vals in opt.all_costs_total.items():
        ref = g[f'{key}_hist_{name}']
        got = np.array([float(v) for v in vals])
        assert len(got) == len(ref), (name, len(got), len(ref))
        assert np.allclose(got, ref, rtol=rtol, atol=0), (name, np.max(np.abs(got - ref) / np.abs(ref)))
    assert list(opt.all_costs_total) == [k.split('_hist_')[1] for k in g.files if k.startswith(f'{key}_hist_')]
    atol = 1e-9 if dt == torch.float64 else 2e-4
    assert np.abs(opt.decomposed_cam_params[2][1].numpy() - g[f'{key}_R']).max() < atol
    assert np.abs(opt.decomposed_cam_params[2][2].numpy() - g[f'{key}_T']).max() < atol * 1e3
    assert np.abs(opt.best_decomposed_cam_params[2][1].numpy() - g[f'{key}_best_R']).max() < atol
    assert np.abs(opt.best_decomposed_cam_params[2][2].numpy() - g[f'{key}_best_T']).max() < atol * 1e3
    assert tuple(opt.decomposed_cam_params[2][1].shape) == (3, 3) and tuple(opt.decomposed_cam_params[2][2].shape) == (3, 1)
    assert isinstance(opt.all_costs_total['total_cost'][0], torch.Tensor)
    out = capsys.readouterr().out
    assert 'Iteration 0: total_cost:' in out and 'extrinsic_param_sample_cost' in out
    if key.endswith('stop'):
        assert 'Early stopping at iteration' in out


def test_extrinsic_gradient_matches_oracle_on_a_larger_sample_set():
    """The 14 sums of the cost/gradient kernel against the oracle's closed form, incl. a non-finite sample."""
    import ctypes
    import torch
    from mc3d_b200 import _lib
    from oracle import extrinsic as E
    import mc3d_b200.synthetic as syn
    rng = np.random.default_rng(12)
    T_, J_, N_ = 40, 17, 9
    cams = syn.ring_rig(3, distortion=True)
    X = syn.SCENE_CENTRE + rng.normal(0, 300, size=(T_, J_, N_, 3))
    X[3, 4, 2] = np.nan
    K, Rm, Tv, dist = cams[2]
    Rm = Rm + rng.normal(0, 1e-3, size=(3, 3))                          # not orthogonal: the 9 entries are free
    mean = syn.project(X[:, :, 0], cams[2]) + rng.normal(0, 2, size=(T_, J_, 2))
    cov = np.array([[4.0, 0.5], [0.5, 3.0]])
    Sinv = np.broadcast_to(np.linalg.inv(cov), (T_, J_, 2, 2)).copy()
    c, dR, dT, n_ok = E.sample_cost_and_grad(X, K, Rm, Tv, dist, mean, Sinv)
    dev = 'cuda:0'
    s3 = torch.tensor(X, device=dev)
    mu = torch.tensor(mean, device=dev)
    S = torch.tensor(np.stack([Sinv[..., 0, 0], Sinv[..., 0, 1], Sinv[..., 1, 1]], axis=-1), device=dev)
    params = torch.zeros(48, dtype=torch.float64, device=dev)
    params[:9] = torch.tensor(Rm.reshape(9))
    params[9:12] = torch.tensor(np.asarray(Tv).reshape(3))
    ctrl = torch.zeros(64 + 2 * 8, dtype=torch.float64, device=dev)
    ctrl[35] = ctrl[51] = float('inf')
    pb = _lib.ExtrinsicProblem()
    pb.n_frames, pb.n_joints, pb.n_samples, pb.hist_capacity = T_, J_, N_, 8
    pb.patience, pb.max_iter = 100, 100
    pb.lr, pb.beta1, pb.beta2, pb.eps, pb.tolerance = 0.0, 0.9, 0.999, 1e-8, 1e-5      # lr 0: the sums are what is checked
    for i in range(9):
        pb.K[i] = float(np.asarray(K).reshape(9)[i])
    for i in range(5):
        pb.dist[i] = float(np.asarray(dist).reshape(-1)[i])
    pb.samples3d, pb.mean, pb.S, pb.params, pb.ctrl = s3.data_ptr(), mu.data_ptr(), S.data_ptr(), params.data_ptr(), ctrl.data_ptr()
    _lib.check(_lib.lib().mc3d_extrinsic_run_f64(ctypes.byref(pb), 0, 1, None))
    torch.cuda.synchronize()
    hist = ctrl[64:66].cpu().numpy()
    assert np.isclose(hist[0], c, rtol=1e-12)
    # the step kernel zeroes the sums it consumed; recompute them with a stopped-state-free second launch on parity 1
    _lib.check(_lib.lib().mc3d_extrinsic_run_f64(ctypes.byref(pb), 1, 1, None))
    torch.cuda.synchronize()
    m = params[12:24].cpu().numpy()                                      # Adam's first moments after two identical steps
    gref = np.concatenate([dR.reshape(9), dT.reshape(3)])
    gclip = gref * min(1.0, 1.0 / (np.sqrt((gref ** 2).sum()) + 1e-6))
    assert np.allclose(m, gclip * (1 - 0.9 ** 2), rtol=1e-10, atol=1e-18)


@pytest.mark.parametrize('graph', ['1', '0'])
@pytest.mark.parametrize('key', ['f64_joint', 'f32_joint'])
def test_cameras_and_trajectory_learnt_together_match_reference_runs(key, graph, monkeypatch):
    """extrinsic_optimization_IDs=[2] with optimize_trajectory=True (pose_refinement.py:931-961): the trajectory's three
    phases plus the camera gradient / joint clip / camera Adam kernels of csrc/extrinsic.cu, the learnt cameras resident in
    device memory (mc3d_refine_problem.cams_dev), two steps per CUDA-graph replay (graph = '1') or eager launches."""
    import torch
    monkeypatch.setenv('MC3D_JOINT_GRAPH', graph)
    import mc3d_b200.pose_refinement as pr
    import mc3d_b200.synthetic as syn
    g = np.load(GOLD)
    cams = {c: [g[f'cam{c}_K'], g[f'cam{c}_R'], g[f'cam{c}_T'], g[f'cam{c}_dist']] for c in range(3)}
    dt = torch.float64 if key.startswith('f64') else torch.float32
    np.random.seed(3)
    random.seed(3)
    opt = pr.Optimized_3d_Pose_Estimation(g['gaussians'].copy(), g['initial'].copy(),
                                          decomposed_cam_params_initial={i: list(cams[i]) for i in cams},
                                          body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=dt)
    opt.sgd_optimize(extrinsic_optimization_IDs=[2], optimize_trajectory=True, lr=1e-3, lambda_smooth=1e-3, lambda_body_length=1.0,
                     max_iter=14, print_frequency=np.inf, time_interval=[0, 12])
    assert (opt.joint_graph_replays > 0) == (graph == '1')           # 15 steps: 2 eager, 6 replays of 2, 1 eager
    rtol = 1e-9 if dt == torch.float64 else 1e-4
    for name, vals in opt.all_costs_total.items():
        ref = g[f'{key}_hist_{name}']
        got = np.array([float(v) for v in vals])
        assert len(got) == len(ref) == 30, (name, len(got))
        assert np.allclose(got, ref, rtol=rtol, atol=0), (name, np.max(np.abs(got - ref) / np.abs(ref)))
    atol = 1e-9 if dt == torch.float64 else 2e-4
    assert np.abs(opt.trajectory.numpy() - g[f'{key}_traj']).max() < atol * 10
    assert np.abs(opt.best_trajectory.numpy() - g[f'{key}_best_traj']).max() < atol * 10
    assert np.abs(opt.decomposed_cam_params[2][1].numpy() - g[f'{key}_R']).max() < atol
    assert np.abs(opt.decomposed_cam_params[2][2].numpy() - g[f'{key}_T']).max() < atol * 1e3
    assert np.abs(opt.best_decomposed_cam_params[2][1].numpy() - g[f'{key}_best_R']).max() < atol
    assert np.abs(opt.best_decomposed_cam_params[2][2].numpy() - g[f'{key}_best_T']).max() < atol * 1e3


def test_superseded_extrinsic_parameter_refinement_class_matches_reference():
    """pose_refinement.ExtrinsicParameterRefinement (reference :233-362) against a run of the unmodified
    class (tests/golden/epr_T10.npz, float32 -- upstream raises in float64): same numpy draws, the triangulated samples, the
    cost of every iteration (the mean log-likelihood it minimises, wrong sign and (T, T, J) broadcast included) and R, T."""
    import torch
    import mc3d_b200.pose_refinement as pr
    g = load_golden('epr_T10.npz')
    assert 'expected scalar type' in str(g['f64_raised'])                     # what upstream does with torch_dtype=float64
    cams = {i: [g[f'cam{i}_{n}'] for n in ('K', 'R', 'T', 'dist')] for i in range(3)}
    np.random.seed(4)
    opt = pr.ExtrinsicParameterRefinement(g['gaussians'].copy(), decomposed_cam_params={i: list(cams[i]) for i in cams},
                                          N_sample_points=5, torch_dtype=torch.float32)
    best = opt.optimize(learning_rate=1e-3, max_iter=24, patience=10, print_frequency=1000)
    assert np.array_equal(np.asarray(opt.samples), g['f32_samples'])           # identical random stream
    assert np.abs(opt.samples_3d.numpy().astype(np.float64) - g['f32_samples3d']).max() < 1e-2   # float32 storage of ~3 m coordinates
    costs = np.array(opt.costs)
    assert len(costs) == len(g['f32_costs'])
    assert np.max(np.abs(costs - g['f32_costs']) / np.abs(g['f32_costs'])) < 1e-3
    assert np.abs(opt.R.numpy() - g['f32_R']).max() < 1e-4 and np.abs(opt.T.numpy() - g['f32_T']).max() < 1e-2
    assert np.abs(best[0].numpy() - g['f32_best_R']).max() < 1e-4 and np.abs(best[1].numpy() - g['f32_best_T']).max() < 1e-2
    # the loss it builds is callable like upstream's (R, T) -> scalar tensor
    val = opt.loss_function(opt.R, opt.T)
    assert isinstance(val, torch.Tensor) and val.ndim == 0
