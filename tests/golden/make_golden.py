"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

Usage (build container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_golden.py [--reference /root/reference]

The reference is imported as-is with harmless ``sys.modules`` stubs for imports the
hot path never touches (matplotlib -- pose_refinement.py:367 ``import
matplotlib.pyplot`` -- and mmpose/tqdm -- mmpose_pose_estimation.py:5-8).  Inputs come
from ``multi-camera_3d_pose_estimation_b200/synthetic.py`` with fixed seeds; only
inputs, outputs and library versions are stored.  No reference source is copied.
"""
import argparse
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference(ref_dir):
    for name in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.animation', 'mmpose', 'mmpose.apis',
                 'mmpose.structures', 'mmpose.utils', 'mmpose.evaluation', 'mmpose.evaluation.functional',
                 'mmdet', 'mmdet.apis']:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules['mmpose.apis'].inference_topdown = None
    sys.modules['mmpose.apis'].init_model = None
    sys.modules['mmpose.structures'].merge_data_samples = None
    sys.modules['mmpose.utils'].adapt_mmdet_pipeline = None
    sys.modules['mmpose.evaluation.functional'].nms = None
    sys.path.insert(0, ref_dir)
    import utils as ref_utils
    import pose_refinement as ref_refine
    import mmpose_pose_estimation as ref_mm
    import pose_estimation as ref_pe
    return ref_utils, ref_refine, ref_mm, ref_pe


def versions():
    import cv2
    import scipy
    import torch
    return np.array([f'numpy {np.__version__}', f'scipy {scipy.__version__}', f'cv2 {cv2.__version__}',
                     f'torch {torch.__version__}'])


def golden_dlt(ref_utils, syn, out):
    import cv2 as cv
    rng = np.random.default_rng(11)
    cams = syn.stereo_rig(distortion=True)
    X = syn.smooth_trajectory(8, 17, rng, centre=(0, 0, 3000.0))
    kp = syn.keypoints_from_trajectory(X, cams, rng)               # (8,17,3,2)
    P = syn.projection_matrices(cams)
    pts = kp[:, :, :2, :].reshape(-1, 2, 2)                        # (N, xy, cam)
    dlt = np.array([ref_utils.DLT(P[0], P[1], p[:, 0], p[:, 1]) for p in pts])
    pair = np.ascontiguousarray(np.transpose(pts, (0, 2, 1)))      # (N, cam, xy)
    c0, c1 = cams[0], cams[1]
    tri = ref_utils.triangulate_points(pair.reshape(8, 17, 2, 2), c0[0], c0[3], c0[1], c0[2],
                                       c1[0], c1[3], c1[1], c1[2])
    und0 = cv.undistortPoints(pair[:, 0, :][:, None, :], c0[0], c0[3], None, c0[0])[:, 0, :]
    und1 = cv.undistortPoints(pair[:, 1, :][:, None, :], c1[0], c1[3], None, c1[0])[:, 0, :]
    proj_pts = X[:3]
    proj = {f'proj_cam{i}': ref_utils.project_points(proj_pts, cams[i][0], cams[i][1], cams[i][2], cams[i][3]) for i in cams}
    proj['proj_flat_nodist'] = ref_utils.project_points(proj_pts.reshape(-1, 3), c1[0], c1[1], c1[2])
    proj['uv_c2d'] = ref_utils.compute_2d_coordinates(P[1], X[0, 0])
    np.savez(os.path.join(out, 'dlt_stereo.npz'), P=P, pts=pts, dlt=dlt, pair=pair, tri=tri, proj_pts=proj_pts, **proj,
             und0=und0, und1=und1, versions=versions(),
             **{f'cam{i}_{n}': np.asarray(cams[i][k]) for i in cams for k, n in enumerate(['K', 'R', 'T', 'dist'])})


def golden_pose3d(ref_pe, syn, out):
    rng = np.random.default_rng(12)
    for n_cams, tag in [(2, 'c2'), (3, 'c3')]:
        cams = syn.stereo_rig() if n_cams == 2 else syn.ring_rig(3, distortion=True)
        centre = (0, 0, 3000.0)
        X = syn.smooth_trajectory(6, 17, rng, centre=centre)
        kp = syn.keypoints_from_trajectory(X, cams, rng)
        cam_params = {i: [np.asarray(a) for a in cams[i]] for i in cams}
        with contextlib.redirect_stdout(io.StringIO()):
            p3d = ref_pe.get_pose_3D(cam_params, list(kp))
            p3d_nod = ref_pe.get_pose_3D(cam_params, list(kp), ignore_nonlinear_distortions=True)
            Rw = syn._look_at(np.array([100.0, 50.0, -500.0]))[0]
            p3d_w = ref_pe.get_pose_3D(cam_params, list(kp), world_trans_rot=(Rw, np.zeros(3)))
        np.savez(os.path.join(out, f'pose3d_{tag}.npz'), kpts=kp, p3d=p3d, p3d_nodist=p3d_nod, p3d_world=p3d_w,
                 Rw=Rw, versions=versions(),
                 **{f'cam{i}_{n}': np.asarray(cams[i][k]) for i in cams
                    for k, n in enumerate(['K', 'R', 'T', 'dist'])})


def golden_config1(ref_pe, syn, out):
    """BASELINE.json configs[0] at its full size: 2-camera COCO-17 triangulation of 400 synthetic frames through the
    reference's own ``get_pose_3D`` (the record_and_estimate_pose 3D step).  Keypoints are rounded to float32 first so
    that the fixture stores them in half the bytes; the reference ran on exactly those values."""
    rng = np.random.default_rng(0)
    cams = syn.stereo_rig(distortion=True)
    X = syn.smooth_trajectory(400, 17, rng, centre=(0, 0, 3000.0))
    kp32 = syn.keypoints_from_trajectory(X, cams, rng).astype(np.float32)          # (400, 17, 3, 2)
    kp = kp32.astype(np.float64)
    cam_params = {i: [np.asarray(a) for a in cams[i]] for i in cams}
    with contextlib.redirect_stdout(io.StringIO()):
        p3d = ref_pe.get_pose_3D(cam_params, list(kp))
    np.savez_compressed(os.path.join(out, 'pose3d_config1.npz'), kpts=kp32, p3d=p3d, versions=versions(),
                        **{f'cam{i}_{n}': np.asarray(cams[i][k]) for i in cams for k, n in enumerate(['K', 'R', 'T', 'dist'])})


def golden_moments(ref_mm, syn, out):
    import torch
    hm, _ = syn.gaussian_blob_heatmaps(17, seed=13)
    hm[5] = 0.004                       # below threshold everywhere -> six zeros (:191-193)
    hm[6, :, :] = 0.0
    hm[6, 10, 7] = 0.7                  # single pixel
    moments_np = ref_mm.PoseEstimator.get_heatmap_means_cov(None, hm.copy())
    moments_t = ref_mm.PoseEstimator.get_heatmap_means_cov(None, torch.tensor(hm.copy()))
    means, stds = ref_mm.PoseEstimator.get_heatmap_means_stds(torch.tensor(np.where(hm < 0.01, 0, hm)))
    np.savez(os.path.join(out, 'heatmap_moments.npz'), heatmaps=hm, moments=moments_np, moments_torch_in=moments_t,
             means=np.array(means), stds=np.array(stds), versions=versions())


def golden_argmax(syn, out):
    """NOT the reference: an independent third-party restatement of the argmax half of mmpose's decode.

    mmpose itself is neither vendored nor pinned upstream and is not installed here, so the quarter-pixel
    decode stays "parity unpinned".  What can be pinned is its first half: HF transformers' ViTPose image
    processor carries a port of mmpose's ``_get_max_preds`` (flat argmax -> (x, y), score = maximum,
    (-1, -1) where the maximum is <= 0) as ``get_keypoint_predictions``; its outputs on fixed heatmaps are
    stored here and both the oracle and the CUDA kernel are held to them (integer pixel + score)."""
    import transformers
    from transformers.models.vitpose.image_processing_vitpose import get_keypoint_predictions
    hm, _ = syn.gaussian_blob_heatmaps(17, seed=21)
    hm[2] = 0.0                          # maximum <= 0 -> (-1, -1), score 0
    hm[3] = -0.5                         # all negative -> (-1, -1)
    hm[4, 0, 0] = 3.0                    # maximum in a corner
    hm[5, 63, 47] = 3.0                  # ... and in the last pixel
    hm[6, 20, 11] = hm[6, 40, 30] = 2.5  # a tie: the first one in flat order wins
    hm[7, 31, 0] = 2.0                   # on the left border
    coords, scores = get_keypoint_predictions(hm[None].copy())
    np.savez_compressed(os.path.join(out, 'argmax_vitpose.npz'), heatmaps=hm, coords=coords[0], scores=scores[0, :, 0],
                        versions=np.array([f'numpy {np.__version__}', f'transformers {transformers.__version__}']))


def golden_refine(ref_refine, ref_utils, syn, out):
    import torch
    torch.manual_seed(0)
    torch.set_num_threads(1)
    g, init, cams, _ = syn.refinement_inputs(48, n_cams=2, seed=14)
    # a NaN joint in the initial trajectory exercises nan_mean masking (pose_refinement.py:221-229)
    lengths = dict(syn.EXAMPLE_BODY_LENGTHS)
    store = dict(gaussians=g, init=init, versions=versions(),
                 **{f'cam{i}_{n}': np.asarray(cams[i][k]) for i in cams
                    for k, n in enumerate(['K', 'R', 'T', 'dist'])})
    # projection
    for tag, dt in [('f32', torch.float32), ('f64', torch.float64)]:
        for i in cams:
            K, R, T, dist = cams[i]
            pr = ref_refine.project_points_torch(init, K, R, T, dist, torch_dtype=dt)
            store[f'proj_{tag}_cam{i}'] = pr.numpy()
            pr = ref_refine.project_points_torch(init, K, R, T, dist, torch_dtype=dt, ignore_distortions=True)
            store[f'proj_nodist_{tag}_cam{i}'] = pr.numpy()
    # optimisation runs
    runs = {
        'readme': dict(lr=0.01, lambda_smooth=1e-6, lambda_body_length=1, patience=100, max_iter=60,
                       time_interval=[0, 40]),
        'defaults': dict(max_iter=25),
        'stop': dict(lr=0.01, lambda_smooth=1e-6, lambda_body_length=1, patience=3, tolerance=0.7, max_iter=200,
                     time_interval=[0, 48]),
        'nodist': dict(lr=0.005, lambda_smooth=0.5, lambda_body_length=0, max_iter=20, ignore_distortions=True,
                       time_interval=[4, 44]),
        # half-overlapping windows of 16 frames stepped one after the other (pose_refinement.py:786-796, :1006):
        # 46 frames -> Time = 32 -> windows [0,16) [8,24) [16,32), three Adam steps per iteration
        'batch': dict(lr=0.01, lambda_smooth=1e-3, lambda_body_length=1, max_iter=12, batch_size=16, time_interval=[1, 47]),
    }
    for tag, dt in [('f32', torch.float32), ('f64', torch.float64)]:
        for rname, kw in runs.items():
            cam_params = {i: [np.asarray(a).copy() for a in cams[i]] for i in cams}
            with contextlib.redirect_stdout(io.StringIO()):
                opt = ref_refine.Optimized_3d_Pose_Estimation(g.copy(), init.copy(),
                                                              decomposed_cam_params_initial=cam_params,
                                                              body_lengths=lengths, torch_dtype=dt)
                kwargs = ref_utils.prepare_kwargs(opt.sgd_optimize, kw)
                opt.sgd_optimize(**kwargs)
            key = f'run_{rname}_{tag}'
            for cname, hist in opt.all_costs_total.items():
                store[f'{key}_{cname}'] = np.array([float(h) for h in hist], dtype=np.float64)
            store[f'{key}_best'] = opt.best_trajectory.numpy()
            store[f'{key}_final'] = opt.trajectory.detach().numpy()
    # NaN-masked variant: one joint-frame of the initial trajectory is NaN
    init_nan = init.copy()
    init_nan[7, 3, :] = np.nan
    cam_params = {i: [np.asarray(a).copy() for a in cams[i]] for i in cams}
    with contextlib.redirect_stdout(io.StringIO()):
        opt = ref_refine.Optimized_3d_Pose_Estimation(g.copy(), init_nan.copy(),
                                                      decomposed_cam_params_initial=cam_params,
                                                      body_lengths=lengths, torch_dtype=torch.float64)
        # The reference masks non-finite entries in the FORWARD value (nan_mean) but autograd still
        # propagates NaN into the gradient, the clip turns every gradient NaN, and iteration 1 dies
        # with KeyError at pose_refinement.py:1053.  Only the iteration-0 costs are defined.
        try:
            opt.sgd_optimize(lr=0.01, lambda_smooth=1e-6, lambda_body_length=0, max_iter=5, time_interval=[0, 40])
            store['run_nan_f64_raised'] = np.array('none')
        except KeyError as e:
            store['run_nan_f64_raised'] = np.array(f'KeyError {e}')
    for cname, hist in opt.all_costs_total.items():
        store[f'run_nan_f64_iter0_{cname}'] = np.array(float(hist[0]) if len(hist) else np.nan)
    store['init_nan'] = init_nan
    np.savez(os.path.join(out, 'refine_T48.npz'), **store)


def golden_refine_large(ref_refine, ref_utils, syn, out):
    """The unmodified reference at 4 000 frames (68 000 joint-frames: hundreds of thread blocks, each with neighbours, in the
    GPU kernels' block-owned ranges).  Inputs come from the seeded generator, so only results are stored: the cost histories
    and every 97th frame of the final / best trajectories."""
    import torch
    torch.manual_seed(0)
    n = 4000
    g, init, cams, _ = syn.refinement_inputs(n, n_cams=2, seed=77)
    lengths = dict(syn.EXAMPLE_BODY_LENGTHS)
    kw = dict(lr=0.01, lambda_smooth=1e-6, lambda_body_length=1, patience=100, max_iter=11, time_interval=[0, n])
    store = dict(versions=versions(), n_frames=np.array(n), seed=np.array(77), stride=np.array(97))
    for tag, dt in [('f32', torch.float32), ('f64', torch.float64)]:
        cam_params = {i: [np.asarray(a).copy() for a in cams[i]] for i in cams}
        with contextlib.redirect_stdout(io.StringIO()):
            opt = ref_refine.Optimized_3d_Pose_Estimation(g.copy(), init.copy(), decomposed_cam_params_initial=cam_params,
                                                          body_lengths=lengths, torch_dtype=dt)
            opt.sgd_optimize(**ref_utils.prepare_kwargs(opt.sgd_optimize, kw))
        for cname, hist in opt.all_costs_total.items():
            store[f'{tag}_{cname}'] = np.array([float(h) for h in hist], dtype=np.float64)
        store[f'{tag}_final'] = opt.trajectory.detach().numpy()[::97]
        store[f'{tag}_best'] = opt.best_trajectory.numpy()[::97]
        store[f'{tag}_final_abs_sum'] = np.array(np.abs(opt.trajectory.detach().numpy().astype(np.float64)).sum())
    np.savez_compressed(os.path.join(out, 'refine_T4000.npz'), **store)


def golden_interp(ref_refine, syn, out):
    rng = np.random.default_rng(15)
    X = syn.smooth_trajectory(60, 17, rng, centre=(0, 0, 3000.0))
    X += rng.normal(0, 2.0, size=X.shape)
    spikes = rng.random(X.shape) < 0.04
    X[spikes] += rng.normal(0, 150.0, size=int(spikes.sum()))
    X[10:13, 2, :] = X[9, 2, :]                      # a constant stretch: std = mad = 0
    store = dict(points=X, versions=versions())
    store['default'] = ref_refine.linear_interpolation(X)
    store['k9'] = ref_refine.linear_interpolation(X, k=9, k_std=1.5, median_std=3)
    store['rolling'] = ref_refine.linear_interpolation(X, k=7, use_rolling_average=True)
    store['nomedian'] = ref_refine.linear_interpolation(X, k=4, filter_distance_from_median=False)
    store['two_dim'] = ref_refine.linear_interpolation(X[:, :, 0])
    np.savez(os.path.join(out, 'interp.npz'), **store)


def golden_extrinsic(ref_refine, syn, out):
    """Learning camera 2's extrinsics from samples (pose_refinement.py:684-706, :800-831, :915-1091): four runs of the
    unmodified reference with numpy / random / torch seeded -- float64 and float32, with and without the constant
    smoothness / bone-length terms -- storing the samples it drew, the cost histories and the learnt R, T."""
    import random
    import torch
    gs, init, cams, _ = syn.refinement_inputs(12, n_cams=3, seed=5)
    cams = {k: [np.array(c, dtype=np.float64) for c in v] for k, v in cams.items()}
    cams[2][2] = cams[2][2] + np.array([[15.0], [-10.0], [20.0]])          # camera 2 starts 27 mm off
    store = dict(gaussians=gs, initial=init, versions=versions())
    for cid, cam in cams.items():
        for nm, arr in zip(('K', 'R', 'T', 'dist'), cam):
            store[f'cam{cid}_{nm}'] = np.asarray(arr)
    runs = {'plain': dict(lambda_smooth=0, lambda_body_length=0, max_iter=25),
            'consts': dict(lambda_smooth=1e-3, lambda_body_length=1.0, max_iter=10),
            'stop': dict(lambda_smooth=0, lambda_body_length=0, max_iter=200, patience=4, tolerance=0.05)}
    for dt_name, dt in (('f64', torch.float64), ('f32', torch.float32)):
        for name, kw in runs.items():
            np.random.seed(3)
            random.seed(3)
            torch.manual_seed(3)
            opt = ref_refine.Optimized_3d_Pose_Estimation(gs.copy(), init.copy(),
                                                          decomposed_cam_params_initial={i: list(cams[i]) for i in cams},
                                                          body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), N_sample_points=6, torch_dtype=dt)
            with contextlib.redirect_stdout(io.StringIO()):
                opt.sgd_optimize(extrinsic_optimization_IDs=[2], optimize_trajectory=False, GT_camera_IDs=[0, 1], lr=1e-3,
                                 print_frequency=1000, time_interval=[0, 12], **kw)
            key = f'{dt_name}_{name}'
            store[f'{key}_samples'] = np.asarray(opt.samples)
            store[f'{key}_samples3d'] = opt.samples_3d.numpy().astype(np.float64)
            for cost, vals in opt.all_costs_total.items():
                store[f'{key}_hist_{cost}'] = np.array([float(v) for v in vals])
            store[f'{key}_R'] = opt.decomposed_cam_params[2][1].detach().numpy().astype(np.float64)
            store[f'{key}_T'] = opt.decomposed_cam_params[2][2].detach().numpy().astype(np.float64)
            store[f'{key}_best_R'] = opt.best_decomposed_cam_params[2][1].numpy().astype(np.float64)
            store[f'{key}_best_T'] = opt.best_decomposed_cam_params[2][2].numpy().astype(np.float64)
            store[f'{key}_kw'] = np.array([f'{k}={v}' for k, v in kw.items()])
    # cameras and trajectory learnt together (extrinsic_optimization_IDs with optimize_trajectory=True, :931-961)
    for dt_name, dt in (('f64', torch.float64), ('f32', torch.float32)):
        np.random.seed(3)
        random.seed(3)
        torch.manual_seed(3)
        opt = ref_refine.Optimized_3d_Pose_Estimation(gs.copy(), init.copy(),
                                                      decomposed_cam_params_initial={i: list(cams[i]) for i in cams},
                                                      body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=dt)
        with contextlib.redirect_stdout(io.StringIO()):
            opt.sgd_optimize(extrinsic_optimization_IDs=[2], optimize_trajectory=True, lr=1e-3, lambda_smooth=1e-3,
                             lambda_body_length=1.0, max_iter=14, print_frequency=1000, time_interval=[0, 12])
        key = f'{dt_name}_joint'
        for cost, vals in opt.all_costs_total.items():
            store[f'{key}_hist_{cost}'] = np.array([float(v) for v in vals])
        store[f'{key}_traj'] = opt.trajectory.detach().numpy().astype(np.float64)
        store[f'{key}_best_traj'] = opt.best_trajectory.numpy().astype(np.float64)
        store[f'{key}_R'] = opt.decomposed_cam_params[2][1].detach().numpy().astype(np.float64)
        store[f'{key}_T'] = opt.decomposed_cam_params[2][2].detach().numpy().astype(np.float64)
        store[f'{key}_best_R'] = opt.best_decomposed_cam_params[2][1].numpy().astype(np.float64)
        store[f'{key}_best_T'] = opt.best_decomposed_cam_params[2][2].numpy().astype(np.float64)
    # plain trajectory optimisation with a subset of the cameras in the likelihood (camera_IDs, pose_refinement.py:866)
    opt = ref_refine.Optimized_3d_Pose_Estimation(gs.copy(), init.copy(), decomposed_cam_params_initial={i: list(cams[i]) for i in cams},
                                                  body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), camera_IDs=[0, 2], torch_dtype=torch.float64)
    with contextlib.redirect_stdout(io.StringIO()):
        opt.sgd_optimize(lr=0.01, lambda_smooth=1e-3, lambda_body_length=1.0, max_iter=8, print_frequency=1000, time_interval=[0, 12])
    for cost, vals in opt.all_costs_total.items():
        store[f'f64_subset_hist_{cost}'] = np.array([float(v) for v in vals])
    store['f64_subset_traj'] = opt.trajectory.detach().numpy().astype(np.float64)
    np.savez_compressed(os.path.join(out, 'extrinsic_T12.npz'), **store)


SURFACE = {
    'utils': ['DLT', 'triangulate_points', 'get_projection_matrix', 'calculate_projection_matrix', 'get_params_from_name',
              '_make_homogeneous_rep_matrix', 'project_points', 'compute_2d_coordinates', 'rotation_conversion',
              'prepare_kwargs', 'load_config', 'get_function_defaults', 'get_body_part_lengths', 'get_body_part_vects',
              'read_camera_parameters', 'read_rotation_translation', 'generate_connectivity_names'],
    'pose_refinement': ['linear_interpolation', 'project_points_torch', 'gaussian_likelihood', 'nan_mean',
                        'Optimized_3d_Pose_Estimation', 'Optimized_3d_Pose_Estimation.sgd_optimize',
                        'Optimized_3d_Pose_Estimation.create_batch_indices', 'Optimized_3d_Pose_Estimation.sample_gaussians',
                        'Optimized_3d_Pose_Estimation.compute_likelihood_cost',
                        'Optimized_3d_Pose_Estimation.compute_smoothness_cost',
                        'Optimized_3d_Pose_Estimation.compute_body_length_cost',
                        'Optimized_3d_Pose_Estimation.gaussian_likelihood',
                        'Optimized_3d_Pose_Estimation.create_body_length_vect',
                        'ExtrinsicParameterRefinement', 'ExtrinsicParameterRefinement.sample_gaussians',
                        'ExtrinsicParameterRefinement.construct_loss', 'ExtrinsicParameterRefinement.optimize'],
    'pose_estimation': ['get_pose_3D'],
    'mmpose_pose_estimation': ['PoseEstimator', 'PoseEstimator.predict', 'PoseEstimator.get_heatmap_means_cov',
                               'PoseEstimator.get_heatmap_means_stds'],
}


def signature_of(module, dotted):
    """[[parameter name, repr(default) or None], ...] of ``module.<dotted>`` (classes: their __init__ without self)."""
    import inspect
    obj = module
    for part in dotted.split('.'):
        obj = getattr(obj, part)
    params = list(inspect.signature(obj).parameters.values())
    return [[p.name, None if p.default is inspect.Parameter.empty else repr(p.default)] for p in params]


def golden_surface(ref_utils, ref_refine, ref_mm, ref_pe, out):
    """Names, parameter order and defaults of the reference's call surface for the path."""
    import json
    mods = {'utils': ref_utils, 'pose_refinement': ref_refine, 'pose_estimation': ref_pe, 'mmpose_pose_estimation': ref_mm}
    surface = {m: {name: signature_of(mods[m], name) for name in names} for m, names in SURFACE.items()}
    with open(os.path.join(out, 'surface.json'), 'w') as fh:
        json.dump(surface, fh, indent=1, sort_keys=True)
        fh.write('\n')


def golden_refine_percam(ref_refine, ref_utils, syn, out):
    """Per-camera Gaussians (the opt-in of SURVEY.md section 8a, Q1).  The reference's vectorised class compares EVERY
    camera's projection with camera 0's Gaussian (pose_refinement.py:663, :885); its superseded Trajectory_Optimization
    indexes the Gaussian of the camera at hand (:499).  This golden is the reference's own optimisation loop
    (sgd_optimize: costs, autograd, clip_grad_norm_, Adam, early stopping -- all unmodified) with ONE method replaced:
    the likelihood cost below is assembled from the reference's project_points_torch, its gaussian_likelihood method and
    nan_mean exactly like :866-889, but with `camera_index` where upstream has the literal 0."""
    import torch
    torch.manual_seed(0)
    torch.set_num_threads(1)

    class PerCameraGaussians(ref_refine.Optimized_3d_Pose_Estimation):
        def compute_likelihood_cost(self):
            if not hasattr(self, '_percam_inv'):
                eye = 1e-6 * torch.eye(2, dtype=self.gaussians.dtype)
                self._percam_inv = [torch.linalg.inv(self.gaussians[:, ci, :, 2:].reshape(self.gaussians.shape[0], self.n_joints, 2, 2)
                                                     + eye).to(self.torch_dtype) for ci in self.camera_indices]
                self._percam_t0 = self._percam_time_interval[0]
            terms = []
            for slot, (ci, cid) in enumerate(zip(self.camera_indices, self.camera_IDs)):
                proj = ref_refine.project_points_torch(self.trajectory, *self.decomposed_cam_params[cid], indicies=self.indicies,
                                                       torch_dtype=self.torch_dtype, ignore_distortions=self.ignore_distortions)
                sub = self.gaussians_subset[self.indicies, ci]
                inv = self._percam_inv[slot][self._percam_t0:][self.indicies]
                terms.append(-self.gaussian_likelihood(proj, sub[..., :2], sub[..., 2:].reshape(len(self.indicies), self.n_joints, 2, 2),
                                                       cov_inv=inv))
            self.likelihood_costs = terms
            self.likelihood_cost = ref_refine.nan_mean(terms)

    g, init, cams, _ = syn.refinement_inputs(32, n_cams=3, seed=23)
    lengths = dict(syn.EXAMPLE_BODY_LENGTHS)
    store = dict(gaussians=g, init=init, versions=versions(),
                 **{f'cam{i}_{n}': np.asarray(cams[i][k]) for i in cams for k, n in enumerate(['K', 'R', 'T', 'dist'])})
    runs = {'a': dict(lr=0.01, lambda_smooth=1e-3, lambda_body_length=1, max_iter=30, time_interval=[0, 32]),
            'b': dict(lr=0.005, lambda_smooth=0.5, lambda_body_length=0, max_iter=15, time_interval=[2, 30], ignore_distortions=True)}
    for tag, dt in [('f32', torch.float32), ('f64', torch.float64)]:
        for rname, kw in runs.items():
            cam_params = {i: [np.asarray(a).copy() for a in cams[i]] for i in cams}
            with contextlib.redirect_stdout(io.StringIO()):
                opt = PerCameraGaussians(g.copy(), init.copy(), decomposed_cam_params_initial=cam_params, body_lengths=lengths,
                                         torch_dtype=dt)
                opt._percam_time_interval = kw['time_interval']
                opt.sgd_optimize(**ref_utils.prepare_kwargs(opt.sgd_optimize, kw))
            key = f'run_{rname}_{tag}'
            for cname, hist in opt.all_costs_total.items():
                store[f'{key}_{cname}'] = np.array([float(h) for h in hist], dtype=np.float64)
            store[f'{key}_best'] = opt.best_trajectory.numpy()
            store[f'{key}_final'] = opt.trajectory.detach().numpy()
    np.savez(os.path.join(out, 'refine_percam_T32.npz'), **store)


def golden_epr(ref_refine, syn, out):
    """The superseded ExtrinsicParameterRefinement class (pose_refinement.py:233-362), unmodified, with a recording wrapper
    around the loss it builds: samples drawn, their triangulation, the cost of every iteration, final and best R, T."""
    import torch
    gs, init, cams, _ = syn.refinement_inputs(10, n_cams=3, seed=9)
    cams = {k: [np.array(c, dtype=np.float64) for c in v] for k, v in cams.items()}
    cams[2][2] = cams[2][2] + np.array([[12.0], [-8.0], [15.0]])
    store = dict(gaussians=gs, versions=versions())
    for cid, cam in cams.items():
        for nm, arr in zip(('K', 'R', 'T', 'dist'), cam):
            store[f'cam{cid}_{nm}'] = np.asarray(arr)

    class Recording(ref_refine.ExtrinsicParameterRefinement):
        def construct_loss(self):
            inner = super().construct_loss()
            self.costs = []

            def wrapped(R, T):
                c = inner(R, T)
                self.costs.append(float(c))
                return c
            self.loss_function = wrapped
            return wrapped

    for dt_name, dt in (('f32', torch.float32), ('f64', torch.float64)):
        np.random.seed(4)
        torch.manual_seed(4)
        try:
            opt = Recording(gs.copy(), decomposed_cam_params={i: list(cams[i]) for i in cams}, N_sample_points=5, torch_dtype=dt)
            with contextlib.redirect_stdout(io.StringIO()):
                best = opt.optimize(learning_rate=1e-3, max_iter=24, patience=10)
        except Exception as exc:                                  # recorded, not hidden
            store[f'{dt_name}_raised'] = np.array(f'{type(exc).__name__}: {exc}'[:300])
            continue
        store[f'{dt_name}_samples'] = np.asarray(opt.samples)
        store[f'{dt_name}_samples3d'] = opt.samples_3d.numpy().astype(np.float64)
        store[f'{dt_name}_costs'] = np.array(opt.costs)
        store[f'{dt_name}_R'] = opt.R.detach().numpy().astype(np.float64)
        store[f'{dt_name}_T'] = opt.T.detach().numpy().astype(np.float64)
        store[f'{dt_name}_best_R'] = best[0].numpy().astype(np.float64)
        store[f'{dt_name}_best_T'] = best[1].numpy().astype(np.float64)
    np.savez(os.path.join(out, 'epr_T10.npz'), **store)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    ap.add_argument('--only', nargs='*', default=None)
    args = ap.parse_args()
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        'mc3d_synthetic', os.path.join(ROOT, 'multi-camera_3d_pose_estimation_b200', 'synthetic.py'))
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    ref_utils, ref_refine, ref_mm, ref_pe = import_reference(args.reference)
    todo = args.only or ['dlt', 'pose3d', 'config1', 'moments', 'argmax', 'refine', 'refine_large', 'refine_percam', 'interp', 'extrinsic', 'epr', 'surface']
    if 'dlt' in todo:
        golden_dlt(ref_utils, syn, HERE)
    if 'pose3d' in todo:
        golden_pose3d(ref_pe, syn, HERE)
    if 'config1' in todo:
        golden_config1(ref_pe, syn, HERE)
    if 'moments' in todo:
        golden_moments(ref_mm, syn, HERE)
    if 'argmax' in todo:
        golden_argmax(syn, HERE)
    if 'refine' in todo:
        golden_refine(ref_refine, ref_utils, syn, HERE)
    if 'refine_large' in todo:
        golden_refine_large(ref_refine, ref_utils, syn, HERE)
    if 'refine_percam' in todo:
        golden_refine_percam(ref_refine, ref_utils, syn, HERE)
    if 'interp' in todo:
        golden_interp(ref_refine, syn, HERE)
    if 'extrinsic' in todo:
        golden_extrinsic(ref_refine, syn, HERE)
    if 'epr' in todo:
        golden_epr(ref_refine, syn, HERE)
    if 'surface' in todo:
        golden_surface(ref_utils, ref_refine, ref_mm, ref_pe, HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(('.npz', '.json')):
            print(f, os.path.getsize(os.path.join(HERE, f)), 'bytes')


if __name__ == '__main__':
    main()
