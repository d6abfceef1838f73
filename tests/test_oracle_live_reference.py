"""CPU, build container only: the oracle against the UNMODIFIED reference imported live from /root/reference, on
randomised inputs (fresh rigs, shapes, weights and hyper-parameters per seed) -- the committed goldens pin fixed inputs,
this pins the restatement itself.  Skipped wherever /root/reference is absent (the GPU box): nothing here is needed by the
GPU tests, smoke() or bench.py.
"""
import contextlib
import importlib.util
import io
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, rel_err
from oracle import decode as D
from oracle import dlt as O
from oracle import interpolation as I
from oracle import refine as R

REFERENCE = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='the reference tree exists in the build container only')


@pytest.fixture(scope='module')
def ref():
    spec = importlib.util.spec_from_file_location('mc3d_make_golden', os.path.join(GOLDEN, 'make_golden.py'))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    saved_path = list(sys.path)
    try:
        mods = mg.import_reference(REFERENCE)
    finally:
        sys.path[:] = saved_path          # the reference's module names (utils, ...) must not shadow anything later
    return dict(zip(('utils', 'refine', 'mm', 'pe'), mods))


def _random_rig(syn, rng, n_cams):
    """ring_rig with per-seed intrinsics, distortion and a small extra rotation per camera."""
    cams = syn.ring_rig(n_cams, radius=float(rng.uniform(2000, 4500)), distortion=True)
    out = {}
    for i, (K, Rm, T, dist) in cams.items():
        K = np.array(K, dtype=np.float64)
        K[0, 0] *= rng.uniform(0.8, 1.3)
        K[1, 1] *= rng.uniform(0.8, 1.3)
        K[0, 2] += rng.uniform(-40, 40)
        K[1, 2] += rng.uniform(-40, 40)
        dist = np.asarray(dist, dtype=np.float64) * rng.uniform(0.3, 1.5, size=np.shape(dist))
        out[i] = [K, np.asarray(Rm, dtype=np.float64), np.asarray(T, dtype=np.float64), dist]
    return out


@pytest.mark.parametrize('seed', [101, 102, 103])
def test_dlt_and_triangulate_points(ref, syn, seed):
    rng = np.random.default_rng(seed)
    cams = _random_rig(syn, rng, 2)
    X = syn.smooth_trajectory(5, 17, rng, centre=(0, 0, 3000.0)) + rng.normal(0, 300, size=3)
    kp = syn.keypoints_from_trajectory(X, cams, rng)                         # (5, 17, 3, 2)
    P = syn.projection_matrices(cams)
    pts = kp[:, :, :2, :].reshape(-1, 2, 2)
    want = np.array([ref['utils'].DLT(P[0], P[1], p[:, 0], p[:, 1]) for p in pts])
    got = np.array([O.dlt_pair(P[0], P[1], p[:, 0], p[:, 1]) for p in pts])
    assert np.array_equal(got, want)
    k3 = np.concatenate([np.transpose(pts, (0, 2, 1)), np.ones((len(pts), 2, 1))], axis=2)
    assert rel_err(O.dlt_weighted_polished(k3, P), want).max() < 1e-9
    pair = np.ascontiguousarray(np.transpose(pts, (0, 2, 1))).reshape(5, 17, 2, 2)
    c0, c1 = cams[0], cams[1]
    args = (c0[0], c0[3], c0[1], c0[2], c1[0], c1[3], c1[1], c1[2])
    assert rel_err(O.triangulate_points(pair, *args), ref['utils'].triangulate_points(pair, *args)).max() < 1e-10


@pytest.mark.parametrize('seed,n_cams', [(111, 2), (112, 3), (113, 4)])
def test_get_pose_3d_top2(ref, syn, seed, n_cams):
    rng = np.random.default_rng(seed)
    cams = _random_rig(syn, rng, n_cams)
    X = syn.smooth_trajectory(6, 17, rng, centre=(0, 0, 3000.0))
    kp = syn.keypoints_from_trajectory(X, cams, rng)
    kp[1, 4, 2, :] = kp[1, 4, 2, 0]                                          # tied scores: argsort order decides
    frames = list(kp)
    for kw in ({}, {'ignore_nonlinear_distortions': True}):
        with contextlib.redirect_stdout(io.StringIO()):
            want = ref['pe'].get_pose_3D({i: [np.copy(a) for a in cams[i]] for i in cams}, [f.copy() for f in frames], **kw)
        got = O.get_pose_3d(cams, frames, **kw)
        assert rel_err(got, want).max() < 1e-10


@pytest.mark.parametrize('seed', [121, 122])
def test_heatmap_moments(ref, syn, seed):
    rng = np.random.default_rng(seed)
    hm, _ = syn.gaussian_blob_heatmaps(9, seed=seed, noise=float(rng.uniform(0.0, 0.02)))
    hm *= rng.uniform(0.2, 3.0, size=(9, 1, 1)).astype(np.float32)
    a, b = hm.copy(), hm.copy()
    want = ref['mm'].PoseEstimator.get_heatmap_means_cov(None, a)
    got = D.heatmap_means_cov(b)
    assert np.array_equal(a, b)                                              # both thresholded in place (Q7)
    # float32 reductions in a different summation order; noise above the threshold makes variances of several hundred px^2
    assert np.allclose(got, want, rtol=2e-6, atol=2e-5)
    assert np.allclose(D.heatmap_means_cov_f64(hm), want, rtol=1e-5, atol=5e-5)


@pytest.mark.parametrize('seed', [131, 132, 133])
def test_refinement_histories(ref, syn, seed):
    import torch
    rng = np.random.default_rng(seed)
    T = int(rng.integers(20, 40))
    g, init, cams, _ = syn.refinement_inputs(T, n_cams=2, seed=seed)
    kw = dict(lr=float(rng.choice([1e-3, 5e-3, 1e-2])), lambda_smooth=float(10.0 ** rng.uniform(-6, 0)),
              lambda_body_length=float(rng.choice([0.0, 0.3, 1.0])), max_iter=int(rng.integers(6, 14)),
              patience=int(rng.integers(2, 50)), tolerance=float(10.0 ** rng.uniform(-6, -1)),
              time_interval=[int(rng.integers(0, 4)), T - int(rng.integers(0, 4))],
              ignore_distortions=bool(rng.integers(0, 2)))
    torch.set_num_threads(1)
    with contextlib.redirect_stdout(io.StringIO()):
        opt = ref['refine'].Optimized_3d_Pose_Estimation(
            g.copy(), init.copy(), decomposed_cam_params_initial={i: [np.asarray(a).copy() for a in cams[i]] for i in cams},
            body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=torch.float64)
        opt.sgd_optimize(**ref['utils'].prepare_kwargs(opt.sgd_optimize, kw))
    out = R.sgd_optimize(g, init, list(cams.values()), syn.EXAMPLE_BODY_LENGTHS, dtype=np.float64, **kw)
    for name, hist in out['history'].items():
        want = np.array([float(h) for h in opt.all_costs_total[name]])
        assert len(hist) == len(want), (name, kw)
        assert np.max(np.abs(np.array(hist) - want) / np.abs(want)) < 1e-11, (name, kw)
    assert np.abs(out['best'] - opt.best_trajectory.numpy()).max() < 1e-9
    assert np.abs(out['final'] - opt.trajectory.detach().numpy()).max() < 1e-9
    # the three public cost methods on the final trajectory (they read the state sgd_optimize leaves behind)
    opt.compute_likelihood_cost()
    opt.compute_smoothness_cost()
    opt.compute_body_length_cost()
    t0, t1 = kw['time_interval']
    x = opt.trajectory.detach().numpy()
    Sinv = R.cov_inverse(g)[t0:t1]
    cam_list = [[np.asarray(a, dtype=np.float64) for a in cams[i]] for i in cams]
    lik, _, _ = R.likelihood(x, g[t0:t1, 0, :, :2], Sinv, cam_list, kw['ignore_distortions'], grad=False)
    sm, _, _ = R.smoothness(x, kw['lambda_smooth'], grad=False)
    bl, _, _ = R.body_length(x, R.bone_table(syn.EXAMPLE_BODY_LENGTHS), kw['lambda_body_length'], grad=False)
    assert np.isclose(lik, float(opt.likelihood_cost), rtol=1e-11)
    assert np.isclose(sm, float(opt.smoothness_cost), rtol=1e-11)
    assert np.isclose(bl, float(opt.body_length_cost), rtol=1e-11, atol=1e-300)


@pytest.mark.parametrize('seed', [141, 142])
def test_linear_interpolation(ref, syn, seed):
    rng = np.random.default_rng(seed)
    X = syn.smooth_trajectory(int(rng.integers(30, 70)), 17, rng, centre=(0, 0, 3000.0)) + rng.normal(0, 2.0, size=(1, 17, 3))
    spikes = rng.random(X.shape) < 0.05
    X[spikes] += rng.normal(0, 120.0, size=int(spikes.sum()))
    kw = dict(k=int(rng.integers(3, 10)), k_std=float(rng.uniform(1.0, 3.0)), median_std=float(rng.uniform(2.0, 5.0)),
              use_rolling_average=bool(rng.integers(0, 2)), filter_distance_from_median=bool(rng.integers(0, 2)))
    with contextlib.redirect_stdout(io.StringIO()):
        want = ref['refine'].linear_interpolation(X.copy(), **kw)
    got = I.linear_interpolation(X.copy(), **kw)
    assert np.abs(got - want).max() < 1e-9, kw
