"""GPU parity: csrc/refine.cu through the drop-in class against runs of the unmodified reference
(tests/golden/refine_T48.npz) and against the oracle.

Tolerance (BASELINE.json north_star): the loss history at equal iterations within 1e-4 relative.  The float64
path is held to 1e-9; trajectories to 1e-6 mm (float64) / 5e-2 mm (float32 state, the reference's own float32
autograd noise over tens of Adam steps).
"""
import numpy as np
import pytest

from conftest import cams_from_golden, load_golden
from oracle import refine as R
from test_oracle_refine import PERCAM_RUNS, RUNS

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4


@pytest.fixture(scope='module')
def pr():
    import __graft_entry__ as g
    g.build()
    import mc3d_b200.pose_refinement as m
    return m


def _cam_params(g, n):
    return {i: [np.asarray(a).copy() for a in cams_from_golden(g, n)[i]] for i in range(n)}


def _history(opt):
    return {k: np.array([float(v) for v in vals]) for k, vals in opt.all_costs_total.items()}


def test_project_points_torch_matches_reference(pr):
    import torch
    g = load_golden('refine_T48.npz')
    cams = cams_from_golden(g, 2)
    for i in range(2):
        K, Rm, T, dist = cams[i]
        for tag, dt, tol in (('f64', torch.float64, 1e-9), ('f32', torch.float32, 1e-3)):
            out = pr.project_points_torch(g['init'], K, Rm, T, dist, torch_dtype=dt)
            assert isinstance(out, torch.Tensor) and out.dtype == dt and tuple(out.shape) == (48, 17, 2) and not out.is_cuda
            assert np.abs(out.numpy() - g[f'proj_{tag}_cam{i}']).max() < tol
            out = pr.project_points_torch(torch.tensor(g['init']), torch.tensor(K), torch.tensor(Rm), torch.tensor(T),
                                          torch.tensor(dist), torch_dtype=dt, ignore_distortions=True)
            assert np.abs(out.numpy() - g[f'proj_nodist_{tag}_cam{i}']).max() < tol
        sub = pr.project_points_torch(g['init'], K, Rm, T, dist, indicies=[3, 4, 9], torch_dtype=torch.float64)
        assert np.abs(sub.numpy() - g[f'proj_f64_cam{i}'][[3, 4, 9]]).max() < 1e-9


@pytest.mark.parametrize('tag,sweep', [('f32', '0'), ('f64', '2')])
def test_persistent_kernel_forms_not_taken_by_default(pr, syn, tag, sweep, monkeypatch):
    """The float two-pass form (MC3D_REFINE_SWEEP=0) and the double-state sweep (=2) of the persistent kernel against the
    reference's 'readme' run: both are A/B paths that must stay correct."""
    import torch
    import mc3d_b200.utils as u
    monkeypatch.setenv('MC3D_REFINE_SWEEP', sweep)
    g = load_golden('refine_T48.npz')
    dt = torch.float64 if tag == 'f64' else torch.float32
    opt = pr.Optimized_3d_Pose_Estimation(g['gaussians'].copy(), g['init'].copy(), decomposed_cam_params_initial=_cam_params(g, 2),
                                          body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=dt)
    opt.sgd_optimize(**u.prepare_kwargs(opt.sgd_optimize, RUNS['readme']))
    key = f'run_readme_{tag}'
    for name, h in _history(opt).items():
        ref = g[f'{key}_{name}']
        assert len(h) == len(ref)
        assert np.max(np.abs(h - ref) / np.abs(ref)) < (1e-9 if tag == 'f64' else LOSS_RTOL), name
    assert np.abs(opt.trajectory.numpy() - g[f'{key}_final']).max() < (1e-6 if tag == 'f64' else 5e-2)


@pytest.mark.parametrize('tag', ['f64', 'f32'])
def test_repeated_pass1_gives_the_same_run(pr, syn, tag, monkeypatch):
    """The persistent kernel folds 1/N_lik and 2 lambda_s/N_s into a stored gradient component with the counts of the previous
    step and repeats pass 1 when a step's own counts differ (a value turned non-finite).  mc3d_refine_problem.test_flags bit 0
    declares the counts changed at every fourth step: the run must equal the undisturbed one (the repeated pass adds its sums
    in another order, hence the float tolerance).  4 000 frames: many blocks, in the fused sweep (float) and the two passes (double)."""
    import torch
    import mc3d_b200.utils as u
    n = 4000
    gs, init, cams, _ = syn.refinement_inputs(n, n_cams=2, seed=78)
    dt = torch.float64 if tag == 'f64' else torch.float32
    kw = dict(lr=0.01, lambda_smooth=1e-6, lambda_body_length=1, patience=100, max_iter=13, time_interval=[0, n])

    def run(flags):
        monkeypatch.setenv('MC3D_REFINE_TEST_FLAGS', flags)
        opt = pr.Optimized_3d_Pose_Estimation(gs.copy(), init.copy(), decomposed_cam_params_initial={i: [np.asarray(a).copy() for a in cams[i]] for i in cams},
                                              body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=dt)
        opt.sgd_optimize(**u.prepare_kwargs(opt.sgd_optimize, kw))
        return _history(opt), opt.trajectory.numpy(), np.array(opt.best_trajectory)

    h0, x0, b0 = run('0')
    h1, x1, b1 = run('1')
    rtol = 1e-10 if tag == 'f64' else 1e-6          # (the body-length cost is a small difference of large sums)
    for name in h0:
        assert len(h0[name]) == len(h1[name]) == 28
        assert np.max(np.abs(h0[name] - h1[name]) / np.abs(h0[name])) < rtol, name
    atol = 1e-9 if tag == 'f64' else 1e-3
    assert np.abs(x0 - x1).max() < atol and np.abs(b0 - b1).max() < atol


@pytest.mark.parametrize('tag', ['f64', 'f32'])
def test_sgd_optimize_matches_reference_at_4000_frames(pr, syn, tag):
    """tests/golden/refine_T4000.npz: the unmodified reference on 4 000 frames x 17 joints (inputs from the seeded generator,
    only results stored).  68 000 joint-frames put hundreds of thread blocks -- each with neighbours on both sides -- into the
    persistent kernel's block-owned ranges (float state: the fused sweep)."""
    import torch
    import mc3d_b200.utils as u
    g = load_golden('refine_T4000.npz')
    n, stride = int(g['n_frames']), int(g['stride'])
    gs, init, cams, _ = syn.refinement_inputs(n, n_cams=2, seed=int(g['seed']))
    dt = torch.float64 if tag == 'f64' else torch.float32
    opt = pr.Optimized_3d_Pose_Estimation(gs.copy(), init.copy(), decomposed_cam_params_initial={i: [np.asarray(a).copy() for a in cams[i]] for i in cams},
                                          body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=dt)
    kw = dict(lr=0.01, lambda_smooth=1e-6, lambda_body_length=1, patience=100, max_iter=11, time_interval=[0, n])
    opt.sgd_optimize(**u.prepare_kwargs(opt.sgd_optimize, kw))
    hist = _history(opt)
    rtol = 1e-9 if tag == 'f64' else LOSS_RTOL
    for name, h in hist.items():
        ref = g[f'{tag}_{name}']
        assert len(h) == len(ref) == 24, (name, len(h), len(ref))      # 12 iterations, each followed by the running mean (Q5)
        assert np.max(np.abs(h - ref) / np.abs(ref)) < rtol, (name, np.max(np.abs(h - ref) / np.abs(ref)))
    atol = 1e-6 if tag == 'f64' else 5e-2
    assert np.abs(opt.trajectory.numpy()[::stride] - g[f'{tag}_final']).max() < atol
    assert np.abs(np.array(opt.best_trajectory)[::stride] - g[f'{tag}_best']).max() < atol
    total = np.abs(opt.trajectory.numpy().astype(np.float64)).sum()
    assert abs(total - float(g[f'{tag}_final_abs_sum'])) / total < (1e-12 if tag == 'f64' else 1e-6)


@pytest.mark.parametrize('run', sorted(RUNS))
@pytest.mark.parametrize('tag', ['f64', 'f32'])
def test_sgd_optimize_matches_reference_runs(pr, syn, run, tag, capsys):
    import torch
    g = load_golden('refine_T48.npz')
    dt = torch.float64 if tag == 'f64' else torch.float32
    opt = pr.Optimized_3d_Pose_Estimation(g['gaussians'].copy(), g['init'].copy(),
                                          decomposed_cam_params_initial=_cam_params(g, 2),
                                          body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=dt)
    import mc3d_b200.utils as u
    opt.sgd_optimize(**u.prepare_kwargs(opt.sgd_optimize, RUNS[run]))
    key = f'run_{run}_{tag}'
    hist = _history(opt)
    rtol = 1e-9 if tag == 'f64' else LOSS_RTOL
    assert set(hist) == {k[len(key) + 1:] for k in g.files if k.startswith(key) and k.endswith('_cost')}
    for name, h in hist.items():
        ref = g[f'{key}_{name}']
        assert len(h) == len(ref), (name, len(h), len(ref))           # same iteration count incl. early stopping
        assert np.max(np.abs(h - ref) / np.abs(ref)) < rtol, name
    atol = 1e-6 if tag == 'f64' else 5e-2
    assert isinstance(opt.best_trajectory, torch.Tensor) and not opt.best_trajectory.is_cuda
    assert np.abs(np.array(opt.best_trajectory) - g[f'{key}_best']).max() < atol
    assert np.abs(opt.trajectory.numpy() - g[f'{key}_final']).max() < atol
    # element types of the history follow upstream: tensors for step costs, numpy scalars for running means
    tc = opt.all_costs_total['total_cost']
    steps_per_iteration = 3 if run == 'batch' else 1                   # one Adam step per window, then the running mean
    assert all(isinstance(c, torch.Tensor) for c in tc[:steps_per_iteration]) and isinstance(tc[steps_per_iteration], np.floating)
    out = capsys.readouterr().out
    assert 'Iteration 0: total_cost:' in out
    if run == 'stop':
        assert 'Early stopping at iteration' in out


@pytest.mark.parametrize('dt_name', ['f64', 'f32'])
def test_exchange_and_persistent_kernel_paths_on_one_gpu(pr, syn, dt_name, monkeypatch):
    """Every way to run the same optimisation on one GPU must agree: the graph of three kernels without and with an
    exchange block, and the two-phase step (gradient components + one reduction of 17 sums; the default) as a graph
    of two kernels (tickets, sequence flags, rank-ordered sums; this GPU is its own only peer) and as the persistent
    cooperative kernel (LL words, grid barriers)."""
    import torch
    dt = torch.float64 if dt_name == 'f64' else torch.float32
    gs, init, cams, _ = syn.refinement_inputs(120, n_cams=2, seed=23)
    init[40, 3] = np.nan
    kw = dict(lr=0.01, lambda_smooth=1e-3, lambda_body_length=1.0, max_iter=60, time_interval=[0, 120], patience=10 ** 6,
              print_frequency=np.inf)
    from mc3d_b200 import refinement as rf
    made = []
    orig = rf.PeerExchange.__init__

    def spy(self, *a, **k):
        made.append(1)
        orig(self, *a, **k)
    monkeypatch.setattr(rf.PeerExchange, '__init__', spy)

    def run(peer, fused, two_phase='0'):
        monkeypatch.setenv('MC3D_REFINE_PEER', peer)
        monkeypatch.setenv('MC3D_REFINE_FUSED', fused)
        monkeypatch.setenv('MC3D_REFINE_TWO_PHASE', two_phase)
        opt = pr.Optimized_3d_Pose_Estimation(gs.copy(), init.copy(), decomposed_cam_params_initial={i: list(cams[i]) for i in cams},
                                              body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=dt)
        opt.sgd_optimize(**kw)
        return opt

    plain = run('0', '1')
    assert not made
    runs = {'three_kernels': run('1', '0'), 'two_phase_graph': run('1', '0', '1'), 'two_phase_persistent': run('1', '1', '1')}
    assert len(made) == 3, 'the exchange block was not allocated'
    h1 = np.array([float(v) for v in plain.all_costs_total['total_cost']])
    # not bitwise: the block partial sums are added with double atomics in arrival order on every path
    rtol, atol = (1e-11, 1e-9) if dt_name == 'f64' else (1e-6, 1e-3)
    for name, other in runs.items():
        h2 = np.array([float(v) for v in other.all_costs_total['total_cost']])
        assert len(h1) == len(h2) == 122 and np.allclose(h1, h2, rtol=rtol, atol=0), name
        assert np.allclose(plain.trajectory.numpy(), other.trajectory.numpy(), rtol=0, atol=atol, equal_nan=True), name
        assert np.allclose(plain.best_trajectory.numpy(), other.best_trajectory.numpy(), rtol=0, atol=atol, equal_nan=True), name
        assert np.isnan(other.trajectory.numpy()[40, 3]).all()
        assert other.iterations == plain.iterations == 61


def test_large_shard_two_phase_persistent_agrees_with_three_kernel_graph(syn, monkeypatch):
    """60 000 frames x 17 joints (the 80-register, 3-CTA/SM build of the persistent kernel; BASELINE config 4 is
    100 000): cost history of the default path against the graph of three kernels, float state."""
    import torch
    from mc3d_b200 import refinement as rf
    n = 60_000
    gs, init, cams, _ = syn.refinement_inputs(n, n_cams=2, seed=31)
    rows = rf.camera_rows(cams, list(cams))

    def history(peer):
        monkeypatch.setenv('MC3D_REFINE_PEER', peer)
        eng = rf.RefineEngine(init, gs, rows, syn.EXAMPLE_BODY_LENGTHS, torch_dtype=torch.float32, device='cuda:0', lr=0.01,
                              betas=(0.9, 0.999), lambda_smooth=1e-6, lambda_body_length=1.0, patience=10 ** 9, tolerance=1e-5,
                              max_iter=10 ** 9, ignore_distortions=False, window=(0, n), n_window_frames=n, hist_capacity=64)
        plan = eng.plan()
        eng.run(24)
        h = eng.history(24)[:, 0].copy()
        x = eng.trajectory().cpu().numpy()
        eng.close()
        return plan, h, x

    p1, h1, x1 = history('1')
    p0, h0, x0 = history('0')
    assert 'two-phase' in p1 and 'persistent' in p1 and 'three-phase' in p0
    assert np.all(np.diff(h1) < 0)                                      # the loss goes down every step
    assert np.allclose(h1, h0, rtol=1e-6)
    assert np.abs(x1 - x0).max() < 1e-3


def test_persistent_kernel_early_stop_matches_three_kernel_graph(pr, syn, monkeypatch):
    """Early stopping inside the persistent kernel (it leaves its loop; the state is stored at both parities)."""
    import torch
    gs, init, cams, _ = syn.refinement_inputs(60, n_cams=2, seed=24)
    kw = dict(lr=0.01, lambda_smooth=1e-3, lambda_body_length=1.0, max_iter=500, time_interval=[0, 60], patience=7,
              tolerance=1e3, print_frequency=np.inf)

    def run(peer):
        monkeypatch.setenv('MC3D_REFINE_PEER', peer)
        opt = pr.Optimized_3d_Pose_Estimation(gs.copy(), init.copy(), decomposed_cam_params_initial={i: list(cams[i]) for i in cams},
                                              body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=torch.float64)
        opt.sgd_optimize(**kw)
        return opt
    a, b = run('0'), run('1')
    assert a.iterations == b.iterations and a.iterations < 400
    ha = np.array([float(v) for v in a.all_costs_total['total_cost']])
    hb = np.array([float(v) for v in b.all_costs_total['total_cost']])
    assert len(ha) == len(hb) and np.allclose(ha, hb, rtol=1e-11)
    assert np.allclose(a.best_trajectory.numpy(), b.best_trajectory.numpy(), atol=1e-9)


def test_gradient_kernel_matches_oracle(pr, syn):
    import torch
    from mc3d_b200 import refinement as rf
    gs, init, cams, _ = syn.refinement_inputs(64, n_cams=3, seed=21)
    init[5, 2] = np.nan                                                 # a frozen joint: masked everywhere
    rows = rf.camera_rows(cams, list(cams))
    eng = rf.RefineEngine(init, gs, rows, syn.EXAMPLE_BODY_LENGTHS, torch_dtype=torch.float64, device='cuda:0', lr=0.01,
                          betas=(0.9, 0.999), lambda_smooth=0.4, lambda_body_length=1.2, patience=10, tolerance=1e-5,
                          max_iter=5, ignore_distortions=False, window=(0, 64), n_window_frames=64, hist_capacity=16)
    eng.phases.phase(eng.problem, 0, 0, True, None)
    eng.phases.phase(eng.problem, 1, 0, True, None)
    torch.cuda.synchronize()
    acc = eng.ctrl[:8].cpu().numpy()
    Sinv = R.cov_inverse(gs)
    bones = R.bone_table(syn.EXAMPLE_BODY_LENGTHS)
    cl, gl, nl = R.likelihood(init, gs[:, 0, :, :2], Sinv, list(cams.values()))
    cs, gsm, ns = R.smoothness(init, 0.4)
    cb, gb, mu = R.body_length(init, bones, 1.2)
    assert np.isclose(acc[0] / acc[1], cl, rtol=1e-12) and acc[1] == nl
    assert np.isclose(0.4 * acc[2] / acc[3], cs, rtol=1e-12) and acc[3] == ns
    assert np.isclose(acc[4] / acc[5], mu, rtol=1e-12)
    ref = gl + gsm + gb
    ref[5, 2] = 0.0                                                      # the NaN joint itself gets no gradient
    got = eng.g.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - np.nan_to_num(ref)).max() < 1e-12 * max(1.0, np.abs(np.nan_to_num(ref)).max())
    assert np.isclose(acc[7], (got ** 2).sum(), rtol=1e-12)


@pytest.mark.parametrize('case', ['batch', 'three_cams_f32', 'subset_of_cameras'])
def test_sgd_optimize_matches_oracle(pr, syn, case):
    import torch
    n_cams = 3 if case != 'batch' else 2
    gs, init, cams, _ = syn.refinement_inputs(90, n_cams=n_cams, seed=22)
    kw = dict(lr=0.01, lambda_smooth=1e-3, lambda_body_length=1.0, max_iter=12, time_interval=[3, 83], patience=50)
    dt, np_dt, rtol = torch.float64, np.float64, 1e-9
    cam_ids = None
    if case == 'batch':
        kw['batch_size'] = 32                                            # windows [0,32) [16,48) [32,64): 3 Adam steps / iteration
    if case == 'three_cams_f32':
        dt, np_dt, rtol = torch.float32, np.float32, LOSS_RTOL
    if case == 'subset_of_cameras':
        cam_ids = [0, 2]
    opt = pr.Optimized_3d_Pose_Estimation(gs.copy(), init.copy(), decomposed_cam_params_initial={i: list(cams[i]) for i in cams},
                                          body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), camera_IDs=cam_ids, torch_dtype=dt)
    opt.sgd_optimize(print_frequency=5, **kw)
    use = list(cams.values()) if cam_ids is None else [cams[i] for i in cam_ids]
    ref = R.sgd_optimize(gs, init, use, syn.EXAMPLE_BODY_LENGTHS, dtype=np_dt, **kw)
    hist = _history(opt)
    for name, h in ref['history'].items():
        assert len(hist[name]) == len(h)
        assert np.max(np.abs(hist[name] - np.array(h)) / np.abs(h)) < rtol, name
    assert np.abs(opt.trajectory.numpy() - ref['final']).max() < (1e-6 if dt == torch.float64 else 5e-2)
    assert tuple(opt.trajectory.shape) == (80, 17, 3)


def test_nan_joint_is_masked_not_fatal(pr, syn):
    """Upstream masks non-finite terms in the forward value only and then dies (KeyError) at iteration 1; here the
    same masked objective is optimised with the NaN joint frozen.  Iteration-0 costs equal upstream's."""
    import torch
    g = load_golden('refine_T48.npz')
    opt = pr.Optimized_3d_Pose_Estimation(g['gaussians'].copy(), g['init_nan'].copy(),
                                          decomposed_cam_params_initial=_cam_params(g, 2),
                                          body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=torch.float64)
    opt.sgd_optimize(lr=0.01, lambda_smooth=1e-6, lambda_body_length=0, max_iter=5, time_interval=[0, 40], print_frequency=np.inf)
    hist = _history(opt)
    for name in ('total_cost', 'likelihood_cost', 'smoothness_cost'):
        assert np.isclose(hist[name][0], float(g[f'run_nan_f64_iter0_{name}']), rtol=1e-10)
    assert len(hist['total_cost']) == 12 and np.isfinite(hist['total_cost']).all()
    final = opt.trajectory.numpy()
    assert np.isnan(final[7, 3]).all() and np.isfinite(np.delete(final.reshape(40, -1), [9, 10, 11], axis=1)).all()


def test_unsupported_modes_and_errors(pr, syn):
    gs, init, cams, _ = syn.refinement_inputs(8, seed=1)
    mk = lambda **k: pr.Optimized_3d_Pose_Estimation(gs, init, decomposed_cam_params_initial={i: list(cams[i]) for i in cams}, **k)
    with pytest.raises(NotImplementedError):
        mk(body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS)).sgd_optimize(randomize_params=True)
    with pytest.raises(NotImplementedError):
        mk(body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS)).sgd_optimize(use_NN=True)
    with pytest.raises(NotImplementedError):                          # learnt cameras need whole-window batches
        mk(body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS)).sgd_optimize(extrinsic_optimization_IDs=[1], batch_size=4, time_interval=[0, 8])
    with pytest.raises(TypeError):                                    # upstream's default GT_camera_IDs expression fails the same way
        mk(body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS)).sgd_optimize(extrinsic_optimization_IDs=[1], optimize_trajectory=False)
    with pytest.raises(AttributeError):
        mk(body_lengths=None).sgd_optimize(max_iter=1)                   # upstream: create_body_length_vect on None
    with pytest.raises(KeyError):
        mk(body_lengths={'left_elbow_nose': 3.0}).sgd_optimize(max_iter=1)


def test_cli_round_trip(pr, syn, tmp_path):
    """`pose_refinement.py --refinement_types SGD` on a recording directory (pose_refinement.py:1099-1256)."""
    import pickle
    import yaml
    gs, init, cams, _ = syn.refinement_inputs(30, seed=23)
    root = tmp_path / 'session'
    run = root / 'recordings' / '0'
    ext = root / 'extrinsic_camera_parameters'
    intr = tmp_path / 'intrinsic_camera_parameters'
    for d in (run, ext, intr):
        d.mkdir(parents=True)
    names = {0: 'camA', 1: 'camB'}
    for i, nm in names.items():
        K, Rm, T, dist = cams[i]
        with open(intr / f'{nm}.dat', 'w') as fh:
            fh.write('intrinsic:\n' + '\n'.join(' '.join(repr(float(v)) for v in row) for row in K) + '\ndistortion:\n' +
                     ' '.join(repr(float(v)) for v in dist.ravel()) + '\n')
        with open(ext / f'rot_trans_{nm}.dat', 'w') as fh:
            fh.write('R:\n' + '\n'.join(' '.join(repr(float(v)) for v in row) for row in Rm) + '\nT:\n' +
                     '\n'.join(repr(float(v)) for v in T.ravel()) + '\n')
    with open(ext / 'camera_names.pkl', 'wb') as fh:
        pickle.dump((names, 'camA'), fh)
    np.save(run / 'kpts_3d.npy', init)
    np.save(run / 'heatmaps_2d.npy', gs)
    np.save(run / 'kpts_2d.npy', np.zeros((30, 17, 3, 2)))
    with open(run / 'recording_log.yaml', 'w') as fh:
        yaml.safe_dump({'kpts_3d': str(run / 'kpts_3d.npy'), 'heatmaps_2d': str(run / 'heatmaps_2d.npy'),
                        'kpts_2d': str(run / 'kpts_2d.npy'), 'recording_paths': [], 'estimator_model': 'x',
                        'detector_model': 'y'}, fh)
    with open(tmp_path / 'params.yaml', 'w') as fh:
        yaml.safe_dump({'SGD': {'max_iter': 15, 'lr': 0.01, 'lambda_smooth': 1e-6, 'lambda_body_length': 1,
                                'patience': 100, 'time_interval': [0, 30]}}, fh)
    with open(tmp_path / 'lengths.yaml', 'w') as fh:
        yaml.safe_dump({'my_lengths': dict(syn.EXAMPLE_BODY_LENGTHS)}, fh, sort_keys=False)
    pr.main(['--run_path', str(run), '--refinement_types', 'SGD', '--intrinsic_params_dir', str(intr),
             '--refinement_params_yaml', str(tmp_path / 'params.yaml'), '--body_part_lengths_yaml', str(tmp_path / 'lengths.yaml')])
    saved = np.load(run / 'kpts_3d_SGD.npy')
    ref = R.sgd_optimize(gs, init, list(cams.values()), syn.EXAMPLE_BODY_LENGTHS, dtype=np.float32, lr=0.01,
                         lambda_smooth=1e-6, lambda_body_length=1, max_iter=15, time_interval=[0, 30])
    assert saved.shape == (30, 17, 3) and np.abs(saved - ref['best']).max() < 5e-2


def test_long_run_at_scale_is_consistent_with_oracle_statistics(pr, syn):
    """Config-4 shape at reduced length (T=4000): the kernel's cost after 200 iterations equals the oracle's cost
    function evaluated on the kernel's own trajectory (size-independent property)."""
    import torch
    gs, init, cams, _ = syn.refinement_inputs(4000, seed=24)
    opt = pr.Optimized_3d_Pose_Estimation(gs, init, decomposed_cam_params_initial={i: list(cams[i]) for i in cams},
                                          body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=torch.float64)
    opt.sgd_optimize(lr=0.01, lambda_smooth=1e-6, lambda_body_length=1, max_iter=199, time_interval=[0, 4000],
                     print_frequency=np.inf, patience=1000)
    hist = _history(opt)
    assert len(hist['total_cost']) == 400
    assert hist['total_cost'][-2] < hist['total_cost'][0]
    # re-evaluate with the oracle at the state BEFORE the last step: run one fewer iteration and compare the next cost
    opt2 = pr.Optimized_3d_Pose_Estimation(gs, init, decomposed_cam_params_initial={i: list(cams[i]) for i in cams},
                                           body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=torch.float64)
    opt2.sgd_optimize(lr=0.01, lambda_smooth=1e-6, lambda_body_length=1, max_iter=198, time_interval=[0, 4000],
                      print_frequency=np.inf, patience=1000)
    x = opt2.trajectory.numpy()
    costs, _ = R.total_cost_and_grad(x, gs[:, 0, :, :2], R.cov_inverse(gs), list(cams.values()),
                                     R.bone_table(syn.EXAMPLE_BODY_LENGTHS), 1e-6, 1.0)
    assert np.isclose(hist['total_cost'][-2], costs['total_cost'], rtol=1e-9)


@pytest.mark.parametrize('tag', ['f64', 'f32'])
def test_public_cost_methods_on_the_final_trajectory(pr, syn, tag):
    """compute_likelihood_cost / compute_smoothness_cost / compute_body_length_cost (pose_refinement.py:836-889) set the
    cost attributes from the state sgd_optimize leaves behind; the oracle's costs of the same trajectory are the check
    (tests/test_oracle_live_reference.py holds the oracle to upstream's own three methods)."""
    import torch
    g = load_golden('refine_T48.npz')
    dt = torch.float64 if tag == 'f64' else torch.float32
    cams = _cam_params(g, 2)
    opt = pr.Optimized_3d_Pose_Estimation(g['gaussians'].copy(), g['init'].copy(), decomposed_cam_params_initial=cams,
                                          body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=dt)
    with pytest.raises(AttributeError):
        opt.compute_smoothness_cost()                                  # no trajectory before sgd_optimize, as upstream
    kw = dict(lr=0.01, lambda_smooth=1e-3, lambda_body_length=0.5, max_iter=6, time_interval=[2, 45], print_frequency=np.inf)
    opt.sgd_optimize(**kw)
    opt.compute_likelihood_cost()
    opt.compute_smoothness_cost()
    opt.compute_body_length_cost()
    x = opt.trajectory.numpy().astype(np.float64)
    gs = g['gaussians'].astype(np.float32 if tag == 'f32' else np.float64).astype(np.float64)
    cam_list = [[np.asarray(a, dtype=np.float32 if tag == 'f32' else np.float64).astype(np.float64) for a in cams[i]] for i in cams]
    Sinv = R.cov_inverse(gs)[2:45]
    lik, _, _ = R.likelihood(x, gs[2:45, 0, :, :2], Sinv, cam_list, False, grad=False)
    sm, _, _ = R.smoothness(x, kw['lambda_smooth'], grad=False)
    bl, _, _ = R.body_length(x, R.bone_table(syn.EXAMPLE_BODY_LENGTHS), kw['lambda_body_length'], grad=False)
    rtol = 1e-9 if tag == 'f64' else LOSS_RTOL
    for got, want in ((opt.likelihood_cost, lik), (opt.smoothness_cost, sm), (opt.body_length_cost, bl)):
        assert isinstance(got, torch.Tensor) and got.dtype == dt and got.dim() == 0
        assert abs(float(got) - want) <= rtol * abs(want), (float(got), want)
    assert tuple(opt.create_body_length_vect().shape) == (43 * len(syn.EXAMPLE_BODY_LENGTHS),)
    d = torch.randn(5, 17, 2, dtype=dt)
    Si = torch.eye(2, dtype=dt).expand(5, 17, 2, 2)
    assert torch.allclose(opt.gaussian_likelihood(d, torch.zeros_like(d), None, cov_inv=Si), -0.5 * (d * d).sum(-1))


@pytest.mark.parametrize('run', sorted(PERCAM_RUNS))
@pytest.mark.parametrize('tag', ['f64', 'f32'])
def test_per_camera_gaussians_opt_in_matches_the_reference_loop(pr, syn, run, tag):
    """`per_camera_gaussians=True` (mc3d_refine_problem.gauss_cam_stride): the reference's own loop with the
    likelihood indexed per camera (tests/golden/make_golden.py::golden_refine_percam), three cameras."""
    import torch
    import mc3d_b200.utils as u
    g = load_golden('refine_percam_T32.npz')
    dt = torch.float64 if tag == 'f64' else torch.float32
    opt = pr.Optimized_3d_Pose_Estimation(g['gaussians'].copy(), g['init'].copy(), decomposed_cam_params_initial=_cam_params(g, 3),
                                          body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=dt, per_camera_gaussians=True)
    opt.sgd_optimize(print_frequency=np.inf, **{k: v for k, v in u.prepare_kwargs(opt.sgd_optimize, PERCAM_RUNS[run]).items()
                                               if k != 'print_frequency'})
    key = f'run_{run}_{tag}'
    hist = _history(opt)
    rtol = 1e-9 if tag == 'f64' else LOSS_RTOL
    for name, h in hist.items():
        ref = g[f'{key}_{name}']
        assert len(h) == len(ref), (name, len(h), len(ref))
        assert np.max(np.abs(h - ref) / np.abs(ref)) < rtol, name
    atol = 1e-6 if tag == 'f64' else 5e-2
    assert np.abs(np.array(opt.best_trajectory) - g[f'{key}_best']).max() < atol
    # the default (upstream's camera-0 Gaussians for every camera) is a different loss on this rig
    ref0 = g[f'{key}_likelihood_cost'][0]
    dflt = pr.Optimized_3d_Pose_Estimation(g['gaussians'].copy(), g['init'].copy(), decomposed_cam_params_initial=_cam_params(g, 3),
                                           body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=dt)
    dflt.sgd_optimize(print_frequency=np.inf, **dict({k: v for k, v in u.prepare_kwargs(dflt.sgd_optimize, PERCAM_RUNS[run]).items()
                                                      if k != 'print_frequency'}, max_iter=0))
    assert abs(float(dflt.all_costs_total['likelihood_cost'][0]) - ref0) > 1e-3 * abs(ref0)
