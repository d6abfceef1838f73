"""Drop-in for the hot-path part of the reference's ``pose_refinement`` module.

``Optimized_3d_Pose_Estimation(...).sgd_optimize(**yaml['SGD'])`` keeps the reference's constructor and
keyword names (pose_refinement.py:579, :894) and result attributes (``best_trajectory``, ``trajectory``,
``all_costs_total``, ...), but every iteration runs as CUDA kernels (csrc/refine.cu) with the state resident on
the GPU; under ``torchrun`` with several ranks the frames are sharded across the GPUs (refinement.py).

Upstream behaviour that is reproduced on purpose: camera 0's Gaussians are
used for every camera; Gaussian means are compared with image-pixel projections unscaled; the default
``time_interval=[0, -1]`` drops the last frame; ``max_iter + 1`` iterations run; the early-stopping statistic is a
running mean over a list that also holds its own earlier values.  ``per_camera_gaussians=True`` is NOT offered:
the kernel keeps one Gaussian per (frame, joint) because that is what upstream evaluates.

``sgd_optimize(extrinsic_optimization_IDs=[id], optimize_trajectory=False, GT_camera_IDs=[a, b])`` learns one
camera's extrinsics from points sampled from the two ground-truth cameras' Gaussians (pose_refinement.py:684-706,
:800-831, :915-1091; csrc/extrinsic.cu) -- SURVEY.md section 8(f3).

With ``optimize_trajectory=True`` the listed cameras are learnt together with the trajectory (:931-961; one GPU,
whole-window batches).

``ExtrinsicParameterRefinement`` (:233-362, upstream's first, superseded extrinsic learner) is mirrored with its quirks; its
loss / gradient evaluations are one kernel launch each.

Out of scope (raise NotImplementedError): ``use_NN``, ``randomize_params``, and the superseded ``Trajectory_Optimization``
class (``per_camera_gaussians=True`` gives its per-camera likelihood inside ``Optimized_3d_Pose_Estimation``).
"""
import argparse
import ctypes
import math
import os
import pickle as pk
import random
from pathlib import Path

import numpy as np
import yaml

from . import _lib
from . import refinement as _ref
from . import utils
from .interpolation import linear_interpolation  # noqa: F401  (module-level name upstream: pose_refinement.py:15)


def _torch():
    import torch
    return torch


# ---- module-level helpers of the reference -------------------------------------------------------------------------
def project_points_torch(points, K, R, T, dist_coeffs, indicies=None, torch_dtype=None, ignore_distortions=False):
    """Pinhole + Brown-distortion projection of a (Time, N, 3) trajectory to (Time, N, 2) pixels
    (pose_refinement.py:94-179), on the GPU.  Returns a CPU torch tensor like upstream (inputs may be numpy arrays,
    CPU or CUDA tensors; a CUDA input gives a CUDA result)."""
    torch = _torch()
    torch_dtype = torch_dtype or torch.float32
    pts = torch.as_tensor(points)
    if pts.dim() != 3 or pts.shape[2] != 3:
        raise AssertionError('points must have shape (Time, N, 3)')
    Kn, Tn = _ref._as_numpy(K), _ref._as_numpy(T)
    dn = np.asarray(_ref._as_numpy(dist_coeffs), dtype=np.float64)
    assert Kn.shape == (3, 3), 'K must have shape (3, 3)'
    assert Tn.shape in ((3, 1), (3,)), 'T must have shape (3,) or (3, 1)'
    assert dn.shape == (1, 5), 'dist_coeffs must have shape (1, 5)'
    row = _ref.camera_rows({0: [Kn, R, Tn, dn]}, [0])[0]
    # upstream rounds the camera to torch_dtype before projecting (:98-112)
    row = np.asarray(torch.tensor(row, dtype=torch_dtype).to(torch.float64).numpy())
    was_cuda = pts.is_cuda
    dev = pts.device if was_cuda else torch.device('cuda', torch.cuda.current_device() if torch.cuda.is_available() else 0)
    if indicies is not None:
        pts = pts[list(indicies)]
    if not was_cuda and not torch.cuda.is_available():
        raise _lib.Mc3dError('project_points_torch needs a CUDA device (no CPU fallback)')
    p = pts.to(device=dev, dtype=torch_dtype).contiguous()
    out = torch.empty(p.shape[:2] + (2,), dtype=torch_dtype, device=dev)
    fn = _lib.lib().mc3d_project_points_f32 if torch_dtype == torch.float32 else _lib.lib().mc3d_project_points_f64
    with torch.cuda.device(dev):
        _lib.check(fn(p.data_ptr(), p.shape[0] * p.shape[1], row.ctypes.data, int(bool(ignore_distortions)),
                      out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out if was_cuda else out.cpu()


def nan_mean(X):
    """Mean over the finite entries of a list of tensors (pose_refinement.py:221-229).  Small helper kept for API
    compatibility; the optimiser computes its masked means inside the kernels."""
    torch = _torch()
    stacked = torch.stack(X)
    mask = ~(torch.isnan(stacked) | torch.isinf(stacked))
    return torch.sum(stacked[mask]) / len(stacked[mask])


def gaussian_likelihood(x, mean, cov_mat, eps=1e-6, torch_dtype=None):
    """Log-likelihood of x under 2D Gaussians incl. the normalisation term (pose_refinement.py:182-218); API
    helper in plain torch ops on the caller's tensors (the optimiser does not call it)."""
    torch = _torch()
    torch_dtype = torch_dtype or torch.float32
    cov = cov_mat + eps * torch.eye(cov_mat.size(-1), device=cov_mat.device).expand_as(cov_mat)
    cov_inv = torch.linalg.inv(cov).to(torch_dtype)
    diff = x - mean
    quad = -0.5 * torch.einsum('...i,...ij,...j->...', diff, cov_inv, diff)
    norm = 0.5 * torch.log((2 * torch.pi) ** 2 * torch.det(cov) + eps)
    return quad - norm


# ---- the optimiser ----------------------------------------------------------------------------------------------------
class Optimized_3d_Pose_Estimation:
    """Maximum-likelihood trajectory refinement under smoothness and bone-length constraints.

    Args (pose_refinement.py:579):
      gaussians                      (Time, C, n_joints, 6) [mean_x, mean_y, var_x, cov, cov, var_y] per camera
      initial_trajectory             (Time, n_joints, 3)
      decomposed_cam_params_initial  dict id -> [K, R, T, dist]   (R/T None -> identity / zero)
      body_lengths                   dict bone name -> target length (utils.POINT_INFO naming)
      camera_IDs                     cameras in the likelihood (default: all)
      torch_dtype                    torch.float32 (default) or torch.float64: dtype of the optimiser state
      device                         extra: CUDA device (default: LOCAL_RANK under torchrun, else current)
    """

    def __init__(self, gaussians, initial_trajectory, decomposed_cam_params_initial=None, body_lengths=None,
                 camera_IDs=None, R_initial=None, T_initial=None, N_sample_points=100, torch_dtype=None, device=None,
                 per_camera_gaussians=False):
        """``per_camera_gaussians`` (not upstream; default off = upstream's behaviour): compare camera c's projection with
        camera c's OWN heatmap Gaussian instead of camera 0's for every camera.  Upstream's vectorised class reads
        ``gaussians[:, 0]`` for all cameras (pose_refinement.py:663, :885 -- SURVEY.md quirk Q1) where its superseded
        ``Trajectory_Optimization`` indexes ``camera_index`` (:499); this switch gives the latter form."""
        torch = _torch()
        torch_dtype = torch_dtype or torch.float32
        self.per_camera_gaussians = bool(per_camera_gaussians)
        if decomposed_cam_params_initial is None:
            raise TypeError("'NoneType' object is not iterable")                  # upstream iterates it unconditionally (:608)
        for cid in decomposed_cam_params_initial:
            if decomposed_cam_params_initial[cid][1] is None:
                decomposed_cam_params_initial[cid][1] = torch.eye(3)
            if decomposed_cam_params_initial[cid][2] is None:
                decomposed_cam_params_initial[cid][2] = torch.zeros(3, 1)
        self.gaussians = torch.as_tensor(gaussians).to(torch_dtype).clone()
        self.decomposed_cam_params_initial = {
            cid: [torch.as_tensor(np.asarray(_ref._as_numpy(cp)), dtype=torch_dtype) for cp in decomposed_cam_params_initial[cid]]
            for cid in decomposed_cam_params_initial}
        self.decomposed_cam_params = {cid: [cp.clone().detach() for cp in self.decomposed_cam_params_initial[cid]]
                                      for cid in self.decomposed_cam_params_initial}
        self.n_cams = self.gaussians.shape[1]
        self.N_sample_points = N_sample_points
        self.initial_trajectory = torch.as_tensor(initial_trajectory).to(torch_dtype).clone()
        self.torch_dtype = torch_dtype
        self.n_dims = self.initial_trajectory.shape[2]
        if self.n_dims != 3:
            raise NotImplementedError('only 3D trajectories are supported')
        self.n_joints = self.gaussians.shape[2]
        self.body_lengths = body_lengths
        self.camera_IDs = camera_IDs if camera_IDs is not None else list(decomposed_cam_params_initial.keys())
        self.camera_indices = [list(self.decomposed_cam_params.keys()).index(cid) for cid in self.camera_IDs]
        self.device = device
        self.best_trajectory = None
        self.best_decomposed_cam_params = None
        self.trajectory = None
        self.all_costs_total = None

    def _pick_device(self):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.Mc3dError('sgd_optimize needs a CUDA device (no CPU fallback)')
        if self.device is not None:
            return torch.device(self.device)
        if 'LOCAL_RANK' in os.environ and torch.distributed.is_available() and torch.distributed.is_initialized():
            return torch.device('cuda', int(os.environ['LOCAL_RANK']))
        return torch.device('cuda', torch.cuda.current_device())

    def create_batch_indices(self):
        """Half-overlapping windows of ``batch_size`` frames (pose_refinement.py:786-796)."""
        step = self.batch_size // 2
        return [list(range(s, s + self.batch_size)) for s in range(0, self.Time - self.batch_size + 1, step)]

    def sgd_optimize(self, extrinsic_optimization_IDs=[], optimize_trajectory=True, lr=0.001, betas=(0.9, 0.999),
                     lambda_smooth=1.0, lambda_body_length=1.0, patience=100, tolerance=1e-5, max_iter=1000,
                     print_frequency=100, batch_size=None, N_sample_points=100, GT_camera_IDs=None,
                     ignore_distortions=False, reset_camera_params=False, print_compute_times=False,
                     time_interval=[0, -1], randomize_params=False, use_NN=False):
        torch = _torch()
        if use_NN or randomize_params:
            raise NotImplementedError('use_NN / randomize_params are outside the accelerated path (SURVEY.md 8(f3))')
        if extrinsic_optimization_IDs is not None and optimize_trajectory is False:
            return self._learn_extrinsics_from_samples(extrinsic_optimization_IDs, GT_camera_IDs, lr=lr, betas=betas,
                                                       lambda_smooth=lambda_smooth, lambda_body_length=lambda_body_length,
                                                       patience=patience, tolerance=tolerance, max_iter=max_iter,
                                                       print_frequency=print_frequency, batch_size=batch_size,
                                                       ignore_distortions=ignore_distortions,
                                                       reset_camera_params=reset_camera_params, time_interval=time_interval)
        learn_ids = list(extrinsic_optimization_IDs) if extrinsic_optimization_IDs is not None else []
        if self.per_camera_gaussians and (learn_ids or not optimize_trajectory):
            raise NotImplementedError('per_camera_gaussians is offered for the trajectory optimisation only '
                                      '(the camera-learning kernels read camera 0\'s Gaussians like upstream)')
        if self.body_lengths is None:
            raise AttributeError("'NoneType' object has no attribute 'values'")   # create_body_length_vect, :770

        t0, t1 = time_interval[0], time_interval[1]
        gaussians_subset = self.gaussians[t0:t1]
        self.Time = len(gaussians_subset)
        if batch_size is None:
            batch_size = self.Time
        self.Time = int(np.floor(self.Time / batch_size) * batch_size)
        self.gaussians_subset = gaussians_subset[:self.Time]
        if reset_camera_params:
            self.decomposed_cam_params = {cid: [cp.clone().detach() for cp in self.decomposed_cam_params_initial[cid]]
                                          for cid in self.decomposed_cam_params_initial}
        self.n_cams = len(self.camera_IDs)
        self.GT_camera_IDs = GT_camera_IDs
        self.ignore_distortions = ignore_distortions
        self.extrinsic_optimization_IDs = extrinsic_optimization_IDs
        self.batch_size = batch_size
        self.lambda_smooth, self.lambda_body_length = lambda_smooth, lambda_body_length
        trajectory0 = self.initial_trajectory[t0:t1].clone().detach()
        batches = self.create_batch_indices()
        windows = [(b[0], b[-1] + 1) for b in batches]
        if not windows:
            raise ValueError('time_interval / batch_size leave no frames to optimise')
        self.indicies = batches[-1]                  # upstream leaves the last batch here (:1006); the cost methods read it

        dist_on = torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size() > 1
        if learn_ids:
            if dist_on or len(windows) > 1:
                raise NotImplementedError('cameras and trajectory are learnt together on one GPU, whole-window batches only')
            for ID in learn_ids:                                 # pose_refinement.py:931-943
                assert ID in self.decomposed_cam_params
                self.decomposed_cam_params_initial[ID][1] = utils.rotation_conversion(self.decomposed_cam_params_initial[ID][1], to_vector=True)
                if tuple(self.decomposed_cam_params[ID][1].shape) != (3, 3):
                    self.decomposed_cam_params[ID][1] = utils.rotation_conversion(
                        self.decomposed_cam_params[ID][1].reshape(3), to_vector=False).to(self.torch_dtype)
                Rl, Tl = self.decomposed_cam_params[ID][1], self.decomposed_cam_params[ID][2]
                Rl[Rl == 0] = random.random() / 10 ** 6
                Tl[Tl == 0] = random.random() / 10 ** 6
        if dist_on and len(windows) > 1:
            raise NotImplementedError('batch_size windows are sequential Adam steps and do not shard; run them on one GPU')
        comm = _ref.DistComm() if dist_on else _ref.LocalComm()
        device = self._pick_device()
        n_iters_max = int(max_iter) + 1 if math.isfinite(max_iter) else 2 ** 31 - 2      # `iteration <= max_iter` (Q4)
        hist_cap = min(n_iters_max * len(windows), 4_000_000)
        cam_rows = _ref.camera_rows(self.decomposed_cam_params, self.camera_IDs)
        cam_rows = torch.tensor(cam_rows, dtype=self.torch_dtype).to(torch.float64).numpy()   # cameras live in torch_dtype upstream

        engine = _ref.RefineEngine(trajectory0, self.gaussians_subset, cam_rows, self.body_lengths,
                                   torch_dtype=self.torch_dtype, device=device, lr=lr, betas=betas,
                                   lambda_smooth=lambda_smooth, lambda_body_length=lambda_body_length,
                                   patience=patience, tolerance=tolerance, max_iter=max_iter,
                                   ignore_distortions=ignore_distortions, window=windows[0],
                                   n_window_frames=batch_size, hist_capacity=hist_cap, comm=comm,
                                   use_exchange=False if learn_ids else None,
                                   gaussian_cameras=list(self.camera_indices) if self.per_camera_gaussians else None)
        self._engine = engine
        joint = _JointCameras(self, engine, learn_ids, lr, betas) if learn_ids else None
        names = ['total_cost', 'likelihood_cost'] + (['smoothness_cost'] if lambda_smooth > 0 else []) + \
                (['body_length_cost'] if lambda_body_length > 0 else [])
        col = {'total_cost': 0, 'likelihood_cost': 1, 'smoothness_cost': 2, 'body_length_cost': 3}
        chunk = max(1, int(print_frequency)) if print_frequency and math.isfinite(print_frequency) else 100
        chunk = min(chunk, 1000)
        iters_done, printed = 0, 0
        stopped_early = False
        with torch.cuda.device(device):
            while iters_done < n_iters_max:
                n = min(chunk, n_iters_max - iters_done)
                if joint is not None:
                    joint.steps(n)
                elif len(windows) == 1:
                    engine.run(n)
                else:
                    for _ in range(n):
                        for wi, (wb, we) in enumerate(windows):
                            engine.set_window(wb, we)
                            engine.one_step(end_of_iteration=(wi == len(windows) - 1))
                st = engine.state()
                iters_done = st['iterations']
                hist = engine.history(st['adam_step'])
                means = _running_means(hist[:, 0], len(windows))
                if print_frequency and math.isfinite(print_frequency):
                    while printed < iters_done:
                        if printed % int(print_frequency) == 0 and not (st['stopped'] and printed == iters_done - 1 and
                                                                        st['no_improve'] >= patience):
                            cur = {nm: _running_means(hist[:, col[nm]], len(windows))[printed] for nm in names}
                            print(f'Iteration {printed}: ' + ', '.join(f'{k}: {v:.2e}' for k, v in cur.items()))
                        printed += 1
                if st['stopped'] or np.isnan(means[-1] if len(means) else 0.0):
                    stopped_early = st['no_improve'] >= patience
                    break
        engine.release_graph()
        st = engine.state()
        hist = engine.history(st['adam_step'])
        if st['adam_step'] > hist_cap:
            import warnings
            warnings.warn(f'{st["adam_step"]} optimiser steps were taken but only the first {hist_cap} are recorded: all_costs_total and the '
                          'printed running means stop there (the early-stopping statistic on the device is unaffected)')
        if stopped_early:
            cur = {nm: _running_means(hist[:, col[nm]], len(windows))[-1] for nm in names}
            print(f"Early stopping at iteration {st['iterations'] - 1}. " + ', '.join(f'{k}: {v:.2e}' for k, v in cur.items()))

        # results, in the reference's containers
        self.trajectory = engine.trajectory().cpu()
        improved_once = math.isfinite(st['best'])
        self.best_trajectory = engine.best_trajectory().cpu() if improved_once else None
        engine.close()                       # sharded runs: unmap the peers' exchange blocks (collective)
        if joint is not None:
            joint.write_back()
            self.joint_graph_replays = joint.graph_replays        # two optimiser steps per replay (0: eager launches only)
        self.best_decomposed_cam_params = {k: [p.clone().detach() for p in self.decomposed_cam_params[k]]
                                           for k in self.decomposed_cam_params} if improved_once else None
        if joint is not None and improved_once:
            for ID, (Rb, Tb) in joint.best.items():
                self.best_decomposed_cam_params[ID][1], self.best_decomposed_cam_params[ID][2] = Rb, Tb
        self.all_costs_total = {}
        for nm in names:
            self.all_costs_total[nm] = _interleave_history(hist[:, col[nm]], len(windows), self.torch_dtype)
        self.iterations = st['iterations']
        return self


    # ---- the three cost terms as public methods (pose_refinement.py:836-889) --------------------------------------------
    def _evaluate_costs(self):
        """One launch of the cost kernel (phase 0 of csrc/refine.cu) on ``self.trajectory`` over the frames
        ``self.indicies`` with the settings of the last ``sgd_optimize`` call; returns the accumulator sums."""
        torch = _torch()
        for name in ('trajectory', 'gaussians_subset', 'Time'):
            if getattr(self, name, None) is None:
                raise AttributeError(f"'Optimized_3d_Pose_Estimation' object has no attribute '{name}'")   # as upstream before sgd_optimize
        idx = list(getattr(self, 'indicies', None) or range(self.Time))
        if idx != list(range(idx[0], idx[-1] + 1)):
            raise NotImplementedError('self.indicies must be a contiguous range of frames (create_batch_indices makes only those)')
        device = self._pick_device()
        cam_rows = _ref.camera_rows(self.decomposed_cam_params, self.camera_IDs)
        cam_rows = torch.tensor(cam_rows, dtype=self.torch_dtype).to(torch.float64).numpy()
        traj = torch.as_tensor(self.trajectory).detach()
        eng = _ref.RefineEngine(traj, self.gaussians_subset, cam_rows, self.body_lengths or {}, torch_dtype=self.torch_dtype,
                                device=device, lr=0.0, betas=(0.9, 0.999), lambda_smooth=1.0, lambda_body_length=1.0,
                                patience=1, tolerance=0.0, max_iter=1, ignore_distortions=getattr(self, 'ignore_distortions', False),
                                window=(idx[0], idx[-1] + 1), n_window_frames=len(idx), hist_capacity=4,
                                comm=_ref.LocalComm(), use_exchange=False,
                                gaussian_cameras=list(self.camera_indices) if self.per_camera_gaussians else None)
        with torch.cuda.device(device):
            eng.phases.phase(eng.problem, 0, 0, True, torch.cuda.current_stream().cuda_stream)
            acc = eng.ctrl[:8].cpu().numpy()
        return acc, eng.problem.aa

    def _as_cost(self, value):
        return _torch().tensor(value, dtype=self.torch_dtype)

    def compute_likelihood_cost(self):
        """Sets ``self.likelihood_cost``: mean over cameras, frames and joints of 0.5 d^T S d (pose_refinement.py:863-889;
        camera-0 Gaussians for every camera, non-finite entries dropped by nan_mean)."""
        acc, _ = self._evaluate_costs()
        self.likelihood_cost = self._as_cost(acc[0] / acc[1] if acc[1] else float('nan'))

    def compute_smoothness_cost(self):
        """Sets ``self.smoothness_cost`` = lambda_smooth * mean_t |x_t - 2 x_{t-1} + x_{t-2}|^2 (pose_refinement.py:836-845)."""
        acc, _ = self._evaluate_costs()
        self.smoothness_cost = self._as_cost(self.lambda_smooth * acc[2] / acc[3] if acc[3] else float('nan'))

    def compute_body_length_cost(self):
        """Sets ``self.body_length_cost`` = lambda_body_length * |a - mu b|^2 / |a|^2 (pose_refinement.py:848-860)."""
        if self.body_lengths is None:
            raise AttributeError("'NoneType' object has no attribute 'keys'")
        acc, aa = self._evaluate_costs()
        mu = acc[4] / acc[5]
        self.body_length_cost = self._as_cost(self.lambda_body_length * (acc[6] - 2.0 * mu * acc[4] + mu * mu * acc[5]) / aa)

    def create_body_length_vect(self):
        """Target bone lengths, each repeated ``batch_size`` times, in yaml key order (pose_refinement.py:768-781)."""
        torch = _torch()
        lengths = torch.tensor(list(self.body_lengths.values()), dtype=self.torch_dtype)
        return lengths.repeat_interleave(self.batch_size)

    def gaussian_likelihood(self, x, mean, cov_mat, eps=1e-6, cov_inv=None):
        """-0.5 d^T Sigma^-1 d without the normalisation term (pose_refinement.py:708-761); plain torch ops on the
        caller's tensors -- API helper, the optimiser evaluates this inside its kernels."""
        torch = _torch()
        if cov_inv is None:
            cov = cov_mat + eps * torch.eye(cov_mat.size(-1), device=cov_mat.device).expand_as(cov_mat)
            cov_inv = torch.linalg.inv(cov).to(self.torch_dtype)
        diff = x - mean
        return -0.5 * torch.einsum('...i,...ij,...j->...', diff, cov_inv, diff)

    # ---- one camera's extrinsics from sampled points (pose_refinement.py:684-706, :800-831, :915-1091) -----------------
    def sample_gaussians(self, N=None):
        """N pixel samples per (frame, GT camera, joint) from the 2D Gaussians, (Time, n_joints, N, 2, 2); the same
        numpy call in the same order as upstream (pose_refinement.py:684-706), so a seeded ``np.random`` gives the
        same draws."""
        if N is None:
            N = self.N_sample_points
        g = self.gaussians_subset
        means = g[:, self.GT_camera_IDs, :, :2]
        covs = g[:, self.GT_camera_IDs, :, 2:].reshape(self.Time, 2, self.n_joints, 2, 2)
        samples = np.empty((self.Time, 2, self.n_joints, N, 2))
        for t in range(self.Time):
            for cam in range(2):
                for point in range(self.n_joints):
                    samples[t, cam, point] = np.random.multivariate_normal(means[t, cam, point].numpy(), covs[t, cam, point].numpy(), N)
        self.samples = np.transpose(samples, (0, 2, 3, 1, 4))
        return self.samples

    def _fixed_trajectory_costs(self, trajectory, device, lambda_smooth, lambda_body_length):
        """Smoothness / bone-length cost of a trajectory that is not being optimised (constants of the total,
        pose_refinement.py:984-986), from one launch of the cost kernel."""
        torch = _torch()
        cam_rows = _ref.camera_rows(self.decomposed_cam_params, self.camera_IDs[:1])
        cam_rows = torch.tensor(cam_rows, dtype=self.torch_dtype).to(torch.float64).numpy()
        eng = _ref.RefineEngine(trajectory, self.gaussians_subset, cam_rows, self.body_lengths, torch_dtype=self.torch_dtype,
                                device=device, lr=0.0, betas=(0.9, 0.999), lambda_smooth=lambda_smooth,
                                lambda_body_length=lambda_body_length, patience=1, tolerance=0.0, max_iter=1,
                                ignore_distortions=True, window=(0, self.Time), n_window_frames=self.Time, hist_capacity=4,
                                comm=_ref.LocalComm(), use_exchange=False)
        with torch.cuda.device(device):
            eng.phases.phase(eng.problem, 0, 0, True, torch.cuda.current_stream().cuda_stream)
            acc = eng.ctrl[:8].cpu().numpy()
        out = {}
        if lambda_smooth > 0:
            out['smoothness_cost'] = lambda_smooth * acc[2] / acc[3] if acc[3] else float('nan')
        if lambda_body_length > 0:
            mu = acc[4] / acc[5]
            out['body_length_cost'] = lambda_body_length * (acc[6] - 2.0 * mu * acc[4] + mu * mu * acc[5]) / eng.problem.aa
        return out

    def _learn_extrinsics_from_samples(self, extrinsic_optimization_IDs, GT_camera_IDs, *, lr, betas, lambda_smooth,
                                       lambda_body_length, patience, tolerance, max_iter, print_frequency, batch_size,
                                       ignore_distortions, reset_camera_params, time_interval):
        torch = _torch()
        if batch_size is not None:
            raise NotImplementedError('batch_size windows are not supported when learning extrinsics from samples')
        if self.body_lengths is None:
            raise AttributeError("'NoneType' object has no attribute 'values'")   # create_body_length_vect, :770
        t0, t1 = time_interval[0], time_interval[1]
        self.gaussians_subset = self.gaussians[t0:t1]
        self.Time = len(self.gaussians_subset)
        if reset_camera_params:
            self.decomposed_cam_params = {cid: [cp.clone().detach() for cp in self.decomposed_cam_params_initial[cid]]
                                          for cid in self.decomposed_cam_params_initial}
        self.n_cams = len(self.camera_IDs)
        self.GT_camera_IDs = GT_camera_IDs
        self.ignore_distortions = ignore_distortions
        self.learning_extrinsics_from_samples = True
        self.extrinsic_optimization_IDs = extrinsic_optimization_IDs
        if self.GT_camera_IDs is None:
            raise TypeError("'NoneType' object is not iterable")                 # upstream's default expression, :920
        assert len(self.extrinsic_optimization_IDs) == 1
        assert len(self.GT_camera_IDs) == 2
        assert min([idx in self.decomposed_cam_params.keys() for idx in self.GT_camera_IDs])
        ID = self.extrinsic_optimization_IDs[0]
        assert ID in self.decomposed_cam_params
        # upstream converts only the INITIAL copy to axis-angle (:934); the live 3x3 matrix is what gets optimised
        self.decomposed_cam_params_initial[ID][1] = utils.rotation_conversion(self.decomposed_cam_params_initial[ID][1], to_vector=True)
        Rl, Tl = self.decomposed_cam_params[ID][1], self.decomposed_cam_params[ID][2]
        if tuple(Rl.shape) != (3, 3):
            Rl = self.decomposed_cam_params[ID][1] = utils.rotation_conversion(Rl.reshape(3), to_vector=False).to(self.torch_dtype)
        Rl[Rl == 0] = random.random() / 10 ** 6                                  # :937-938, one draw per tensor
        Tl[Tl == 0] = random.random() / 10 ** 6
        self.trajectory = self.initial_trajectory[t0:t1].clone().detach()
        self.batch_size = self.Time
        self.lambda_smooth, self.lambda_body_length = lambda_smooth, lambda_body_length
        device = self._pick_device()

        names = ['total_cost'] + (['smoothness_cost'] if lambda_smooth > 0 else []) + \
                (['body_length_cost'] if lambda_body_length > 0 else []) + ['extrinsic_param_sample_cost']
        consts = self._fixed_trajectory_costs(self.trajectory, device, lambda_smooth, lambda_body_length) \
            if (lambda_smooth > 0 or lambda_body_length > 0) else {}

        self.samples = self.sample_gaussians()
        cm1, R1, T1, d1 = self.decomposed_cam_params[self.GT_camera_IDs[0]]
        cm2, R2, T2, d2 = self.decomposed_cam_params[self.GT_camera_IDs[1]]
        self.samples_3d = torch.from_numpy(utils.triangulate_points(self.samples, cm1, d1, R1, T1, cm2, d2, R2, T2)).to(self.torch_dtype)

        dt, tag = self.torch_dtype, ('f32' if self.torch_dtype == torch.float32 else 'f64')
        lib = _lib.lib()
        n_iters_max = int(max_iter) + 1 if math.isfinite(max_iter) else 2 ** 31 - 2
        hist_cap = min(n_iters_max, 4_000_000)
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream().cuda_stream
            g_dev = self.gaussians_subset.to(device).contiguous()
            T_, C_, J_ = int(g_dev.shape[0]), int(g_dev.shape[1]), int(g_dev.shape[2])
            mean = torch.empty((T_, J_, 2), dtype=dt, device=device)
            S = torch.empty((T_, J_, 3), dtype=dt, device=device)
            scratch_mu = torch.empty_like(mean)
            scratch_S = torch.empty_like(S)
            prep = getattr(lib, f'mc3d_refine_prepare_{tag}')
            # means of camera INDEX 2 (hard-coded upstream, :803); inverse covariances of camera 0 (quirk Q1, :663-668)
            _lib.check(prep(g_dev.data_ptr(), T_, C_, J_, 2, 1e-6, mean.data_ptr(), scratch_S.data_ptr(), stream))
            _lib.check(prep(g_dev.data_ptr(), T_, C_, J_, 0, 1e-6, scratch_mu.data_ptr(), S.data_ptr(), stream))
            s3 = self.samples_3d.to(device).contiguous()
            params = torch.zeros(48, dtype=torch.float64, device=device)
            params[:9] = Rl.detach().to(torch.float64).reshape(9).to(device)
            params[9:12] = Tl.detach().to(torch.float64).reshape(3).to(device)
            params[36:48] = params[:12]
            ctrl = torch.zeros(64 + 2 * hist_cap, dtype=torch.float64, device=device)
            ctrl[32 + 3] = math.inf
            ctrl[48 + 3] = math.inf
            pb = _lib.ExtrinsicProblem()
            pb.n_frames, pb.n_joints, pb.n_samples = T_, J_, int(self.samples_3d.shape[2])
            pb.ignore_distortions = int(bool(ignore_distortions))
            pb.patience = int(min(patience, 2 ** 31 - 1)) if math.isfinite(patience) else 2 ** 31 - 1
            pb.max_iter = int(min(max_iter, 2 ** 31 - 2)) if math.isfinite(max_iter) else 2 ** 31 - 2
            pb.hist_capacity = hist_cap
            pb.lr, pb.beta1, pb.beta2, pb.eps, pb.tolerance = float(lr), float(betas[0]), float(betas[1]), 1e-8, float(tolerance)
            pb.const_cost = float(sum(consts.values()))
            Kl = torch.as_tensor(self.decomposed_cam_params[ID][0]).to(torch.float64).reshape(9)
            Dl = torch.zeros(5, dtype=torch.float64)
            dd = torch.as_tensor(self.decomposed_cam_params[ID][3]).to(torch.float64).reshape(-1)
            Dl[:min(5, dd.numel())] = dd[:5]
            for i in range(9):
                pb.K[i] = float(Kl[i])
            for i in range(5):
                pb.dist[i] = float(Dl[i])
            pb.samples3d, pb.mean, pb.S = s3.data_ptr(), mean.data_ptr(), S.data_ptr()
            pb.params, pb.ctrl = params.data_ptr(), ctrl.data_ptr()
            run = getattr(lib, f'mc3d_extrinsic_run_{tag}')

            def state(step):
                st = ctrl[32 + 16 * (step & 1):32 + 16 * (step & 1) + 8].cpu().numpy()
                return dict(adam_step=int(st[0]), best=float(st[3]), no_improve=int(st[4]), stopped=bool(st[5]), iterations=int(st[6]))

            chunk = max(1, int(print_frequency)) if print_frequency and math.isfinite(print_frequency) else 100
            chunk = min(chunk, 1000)
            step, printed, stopped_early = 0, 0, False
            st = state(0)

            def series(hist):
                cols = {'extrinsic_param_sample_cost': hist[:, 0], 'total_cost': hist[:, 1]}
                for k, v in consts.items():
                    cols[k] = np.full(len(hist), v)
                return cols

            while st['iterations'] < n_iters_max:
                n = min(chunk, n_iters_max - st['iterations'])
                _lib.check(run(ctypes.byref(pb), step, n, stream))
                step += n
                st = state(step)
                hist = ctrl[64:64 + 2 * st['adam_step']].cpu().numpy().reshape(-1, 2)
                cols = series(hist)
                if print_frequency and math.isfinite(print_frequency):
                    while printed < st['iterations']:
                        if printed % int(print_frequency) == 0 and not (st['stopped'] and printed == st['iterations'] - 1 and
                                                                        st['no_improve'] >= patience):
                            cur = {nm: _running_means(cols[nm], 1)[printed] for nm in names}
                            print(f'Iteration {printed}: ' + ', '.join(f'{k}: {v:.2e}' for k, v in cur.items()))
                        printed += 1
                if st['stopped']:
                    stopped_early = st['no_improve'] >= patience
                    break
            hist = ctrl[64:64 + 2 * st['adam_step']].cpu().numpy().reshape(-1, 2)
            cols = series(hist)
            if stopped_early:
                cur = {nm: _running_means(cols[nm], 1)[-1] for nm in names}
                print(f"Early stopping at iteration {st['iterations'] - 1}. " + ', '.join(f'{k}: {v:.2e}' for k, v in cur.items()))
            p_host = params.cpu()
        self._extrinsic_launch_plan = 'csrc/extrinsic.cu: 2 kernels per iteration, CUDA graph'
        self.decomposed_cam_params[ID][1] = p_host[:9].reshape(3, 3).to(dt)
        self.decomposed_cam_params[ID][2] = p_host[9:12].reshape(3, 1).to(dt)
        improved_once = math.isfinite(st['best'])
        self.best_trajectory = self.trajectory.clone().detach() if improved_once else None
        if improved_once:
            self.best_decomposed_cam_params = {k: [q.clone().detach() for q in self.decomposed_cam_params[k]]
                                               for k in self.decomposed_cam_params}
            self.best_decomposed_cam_params[ID][1] = p_host[36:45].reshape(3, 3).to(dt)
            self.best_decomposed_cam_params[ID][2] = p_host[45:48].reshape(3, 1).to(dt)
        else:
            self.best_decomposed_cam_params = None
        self.all_costs_total = {nm: _interleave_history(cols[nm], 1, dt) for nm in names}
        self.iterations = st['iterations']
        return self


class ExtrinsicParameterRefinement:
    """Upstream's first, superseded extrinsic learner (pose_refinement.py:233-362): R, T of a third camera from points
    sampled from two ground-truth cameras' heatmap Gaussians.  Kept call-compatible -- same constructor, ``sample_gaussians``,
    ``construct_loss``, ``optimize`` and attributes (``R``, ``T``, ``samples``, ``samples_3d``, ``best_params``,
    ``loss_function``) -- with its arithmetic on the GPU: the samples are triangulated by the batched 2-view kernel and every
    evaluation of the loss and its gradient w.r.t. the 9 entries of R and T is one launch of ``mc3d_extrinsic_costgrad_*``.
    The 12-parameter Adam step and the SVD re-orthogonalisation of R (:337-340) stay on the host, in float32 as upstream.

    Upstream's behaviour, kept on purpose:
      * the loss is the mean LOG-likelihood and it is *minimised* (:310-314, :325-333), i.e. the optimiser pushes the
        projections away from the Gaussians; use ``Optimized_3d_Pose_Estimation.sgd_optimize(extrinsic_optimization_IDs=...)``
        for the corrected form;
      * the covariances are reshaped to (T, 1, J, 2, 2) (:298), which broadcasts against the (T, J) residuals to a (T, T, J)
        table: every sample frame is scored with the inverse covariance of EVERY frame.  The mean over that table equals
        scoring each residual with the time-average of the inverse covariances, which is what the kernel is given;
      * R, T are float32 whatever ``torch_dtype`` is (:240-247), an explicit ``R_initial`` / ``T_initial`` is ignored
        (:245-247), means and covariances are those of camera index 2 (:297-298), and exactly three cameras are required.
    Non-finite samples are not supported (upstream's nan_mean would weight the normalisation term per joint)."""

    def __init__(self, gaussians, R_initial=None, T_initial=None, decomposed_cam_params=None, N_sample_points=100,
                 GT_camera_indicies=[0, 1], estimation_camera_index=2, torch_dtype=None, device=None):
        torch = _torch()
        torch_dtype = torch_dtype or torch.float32
        assert len(GT_camera_indicies) == 2
        assert min([idx in decomposed_cam_params.keys() for idx in GT_camera_indicies])
        if R_initial is None and T_initial is None:
            if estimation_camera_index in decomposed_cam_params:
                self.R = torch.tensor(np.asarray(_ref._as_numpy(decomposed_cam_params[estimation_camera_index][1])), dtype=torch.float32)
                self.T = torch.tensor(np.asarray(_ref._as_numpy(decomposed_cam_params[estimation_camera_index][2])), dtype=torch.float32)
            else:
                self.R = torch.eye(3)
                self.T = torch.zeros(3, 1)
        else:                                                     # upstream ignores the values it was given (:245-247)
            self.R = torch.eye(3)
            self.T = torch.zeros(3, 1)
        self.gaussians = torch.tensor(np.asarray(_ref._as_numpy(gaussians)), dtype=torch_dtype)
        self.decomposed_cam_params = {k: [torch.tensor(np.asarray(_ref._as_numpy(cp)), dtype=torch_dtype) for cp in decomposed_cam_params[k]]
                                      for k in decomposed_cam_params}
        self.Time = gaussians.shape[0]
        self.n_cams = gaussians.shape[1]
        assert self.n_cams == 3
        self.n_joints = gaussians.shape[2]
        self.N_sample_points = N_sample_points
        self.GT_camera_indicies = GT_camera_indicies
        self.estimation_camera_index = estimation_camera_index
        self.torch_dtype = torch_dtype
        self.device = device
        self.best_params = None

    def sample_gaussians(self, N=None):
        """N draws per (frame, ground-truth camera, joint) in upstream's order (:277-286): the same numpy stream."""
        if N is None:
            N = self.N_sample_points
        means = self.gaussians[:, self.GT_camera_indicies, :, :2].numpy()
        covs = self.gaussians[:, self.GT_camera_indicies, :, 2:].reshape(self.Time, 2, self.n_joints, 2, 2).numpy()
        samples = np.empty((self.Time, 2, self.n_joints, N, 2))
        for t in range(self.Time):
            for cam in range(2):
                for point in range(self.n_joints):
                    samples[t, cam, point] = np.random.multivariate_normal(means[t, cam, point], covs[t, cam, point], N)
        self.samples = np.transpose(samples, (0, 2, 3, 1, 4))
        return self.samples

    def construct_loss(self):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.Mc3dError('ExtrinsicParameterRefinement needs a CUDA device (no CPU fallback)')
        dev = torch.device(self.device) if self.device is not None else torch.device('cuda', torch.cuda.current_device())
        dt = self.torch_dtype
        tag = 'f32' if dt == torch.float32 else 'f64'
        cm1, R1, T1, d1 = self.decomposed_cam_params[self.GT_camera_indicies[0]]
        cm2, R2, T2, d2 = self.decomposed_cam_params[self.GT_camera_indicies[1]]
        self.samples_3d = torch.from_numpy(utils.triangulate_points(self.samples, cm1, d1, R1, T1, cm2, d2, R2, T2)).to(dt)
        if not bool(torch.isfinite(self.samples_3d).all()):
            raise NotImplementedError('non-finite samples are not supported by this class')
        g2 = self.gaussians[:, 2]                                             # camera index 2, hard-coded upstream (:297-298)
        cov = g2[..., 2:].reshape(self.Time, self.n_joints, 2, 2) + 1e-6 * torch.eye(2, dtype=dt)
        cov_inv = torch.linalg.inv(cov).to(torch.float32)                      # gaussian_likelihood(..., torch_dtype=torch.float32), :311
        s_bar = cov_inv.to(dt).mean(dim=0)                                     # the (T, T, J) broadcast, averaged over the covariance frame
        S = torch.stack([s_bar[:, 0, 0], 0.5 * (s_bar[:, 0, 1] + s_bar[:, 1, 0]), s_bar[:, 1, 1]], dim=-1)
        norm = 0.5 * torch.log((2 * math.pi) ** 2 * torch.det(cov) + 1e-6)     # (T, J)
        self._norm_const = float(norm.mean())
        lib = _lib.lib()
        with torch.cuda.device(dev):
            self._dev = dev
            self._s3 = self.samples_3d.to(dev).contiguous()
            self._mean = g2[..., :2].to(dt).to(dev).contiguous()
            self._S = S.to(dt).unsqueeze(0).repeat(self.Time, 1, 1).to(dev).contiguous()
            self._params = torch.zeros(48, dtype=torch.float64, device=dev)
            self._ctrl = torch.zeros(64, dtype=torch.float64, device=dev)
            pb = self._pb = _lib.ExtrinsicProblem()
            pb.n_frames, pb.n_joints, pb.n_samples = self.Time, self.n_joints, int(self.samples_3d.shape[2])
            pb.ignore_distortions, pb.patience, pb.max_iter, pb.hist_capacity = 0, 1, 1, 0
            K = self.decomposed_cam_params[self.estimation_camera_index][0].to(torch.float64).reshape(9)
            dd = self.decomposed_cam_params[self.estimation_camera_index][-1].to(torch.float64).reshape(-1)
            for i in range(9):
                pb.K[i] = float(K[i])
            for i in range(min(5, dd.numel())):
                pb.dist[i] = float(dd[i])
            pb.samples3d, pb.mean, pb.S = self._s3.data_ptr(), self._mean.data_ptr(), self._S.data_ptr()
            pb.params, pb.ctrl = self._params.data_ptr(), self._ctrl.data_ptr()
        self._costgrad = getattr(lib, f'mc3d_extrinsic_costgrad_{tag}')

        def loss(R, T, with_grad=False):
            """Mean log-likelihood of the projected samples (the quantity upstream minimises); with_grad: also d loss / d (R, T)."""
            with torch.cuda.device(self._dev):
                self._params[:9] = torch.as_tensor(R).detach().to(torch.float64).reshape(9).to(self._dev)
                self._params[9:12] = torch.as_tensor(T).detach().to(torch.float64).reshape(3).to(self._dev)
                self._ctrl.zero_()
                _lib.check(self._costgrad(ctypes.byref(self._pb), torch.cuda.current_stream().cuda_stream))
                acc = self._ctrl[:14].cpu().numpy()
            if acc[1] != self.Time * self.n_joints * self._pb.n_samples:
                raise NotImplementedError('non-finite projections are not supported by this class')
            value = -acc[0] / acc[1] - self._norm_const
            if not with_grad:
                return torch.tensor(value, dtype=self.torch_dtype)
            return value, (-acc[2:11] / acc[1]).reshape(3, 3), (-acc[11:14] / acc[1]).reshape(3, 1)

        self.loss_function = loss
        return loss

    def optimize(self, learning_rate=0.001, print_frequency=10, max_iter=np.inf, patience=10):
        torch = _torch()
        self.sample_gaussians()
        self.construct_loss()
        f32 = np.float32
        p = np.concatenate([self.R.detach().numpy().reshape(9), self.T.detach().numpy().reshape(3)]).astype(f32)
        m, v = np.zeros(12, f32), np.zeros(12, f32)
        b1, b2, eps = 0.9, 0.999, 1e-8
        best_cost, iteration, no_improvement_count, step = float('inf'), 0, 0, 0
        self.costs = []
        while no_improvement_count < patience and iteration <= max_iter:
            cost, gR, gT = self.loss_function(p[:9].reshape(3, 3), p[9:].reshape(3, 1), with_grad=True)
            g = np.concatenate([gR.reshape(9), gT.reshape(3)]).astype(f32)
            # torch.optim.Adam on two float32 tensors (no clipping here, :325-335)
            step += 1
            m = m + f32(1 - b1) * (g - m)
            v = v * f32(b2) + (f32(1 - b2) * g) * g
            step_size = learning_rate / (1 - b1 ** step)
            denom = np.sqrt(v) / f32(math.sqrt(1 - b2 ** step)) + f32(eps)
            p = (p + f32(-step_size) * (m / denom)).astype(f32)
            U, _, Vt = np.linalg.svd(p[:9].reshape(3, 3))                        # re-orthogonalise R (:337-340)
            p[:9] = (U @ Vt).astype(f32).reshape(9)
            self.R = torch.tensor(p[:9].reshape(3, 3).copy())
            self.T = torch.tensor(p[9:].reshape(3, 1).copy())
            current_cost = f32(cost)
            self.costs.append(float(current_cost))
            if current_cost < best_cost:
                best_cost = current_cost
                self.best_params = [self.R.clone().detach(), self.T.clone().detach()]
                no_improvement_count = 0
            else:
                no_improvement_count += 1
            if no_improvement_count >= patience:
                print(f'Early stopping at iteration {iteration}. Best cost = {best_cost:.2e}')
                break
            if iteration % print_frequency == 0:
                print(f'Iteration {iteration}: Cost = {current_cost:.2e}')
            iteration += 1
        return self.best_params


class _JointCameras:
    """Cameras learnt together with the trajectory (``extrinsic_optimization_IDs`` with ``optimize_trajectory=True``,
    pose_refinement.py:931-961): the trajectory steps through the engine's three phases; between its gradient phase and
    its Adam phase every learnt camera's gradient of the likelihood cost is accumulated over the trajectory points
    (csrc/extrinsic.cu) and one clip_grad_norm_ + Adam step covers everything (:1047-1050).

    Everything stays on the device: the cameras the trajectory's kernels project with live in device memory
    (``mc3d_refine_problem.cams_dev``), the joint step's new R, T are copied into them on the stream, the best cameras are
    kept by a masked copy driven by the ``improved`` flag of the control block, and two consecutive steps are captured once
    into a CUDA graph and replayed; the host reads the state back once per chunk of iterations."""

    def __init__(self, opt, engine, learn_ids, lr, betas):
        torch = _torch()
        self.opt, self.engine, self.ids = opt, engine, list(learn_ids)
        self.lr, self.betas = float(lr), (float(betas[0]), float(betas[1]))
        self.tag = engine.tag
        self.lib = _lib.lib()
        dev, dt = engine.device, opt.torch_dtype
        n = len(self.ids)
        self.cam_params = torch.zeros(36 * n, dtype=torch.float64, device=dev)
        self.cam_ctrl = torch.zeros(64 * n, dtype=torch.float64, device=dev)
        self.slots = [opt.camera_IDs.index(ID) if ID in opt.camera_IDs else None for ID in self.ids]
        self.problems = []
        J = engine.J
        x0 = engine.x_ext[2:]                                    # local frame 0 (no exchange block in this mode)
        for q, ID in enumerate(self.ids):
            K, Rm, Tv, dist = opt.decomposed_cam_params[ID]
            self.cam_params[36 * q:36 * q + 9] = Rm.detach().to(torch.float64).reshape(9).to(dev)
            self.cam_params[36 * q + 9:36 * q + 12] = Tv.detach().to(torch.float64).reshape(3).to(dev)
            pb = _lib.ExtrinsicProblem()
            pb.n_frames, pb.n_joints, pb.n_samples = engine.n_local, J, 1
            pb.ignore_distortions = engine.problem.ignore_distortions
            Kd = torch.as_tensor(K).to(dt).to(torch.float64).reshape(9)
            dd = torch.as_tensor(dist).to(dt).to(torch.float64).reshape(-1)
            for i in range(9):
                pb.K[i] = float(Kd[i])
            for i in range(min(5, dd.numel())):
                pb.dist[i] = float(dd[i])
            pb.samples3d = x0.data_ptr()
            pb.mean, pb.S = engine.mu0.data_ptr(), engine.S.data_ptr()
            pb.params = self.cam_params.data_ptr() + 8 * 36 * q
            pb.ctrl = self.cam_ctrl.data_ptr() + 8 * 64 * q
            self.problems.append(pb)
        # the cameras of the trajectory's phases, in device memory from here on
        n_cams = int(engine.problem.n_cams)
        rows = np.array([[engine.problem.cams[c][i] for i in range(26)] for c in range(n_cams)], dtype=np.float64)
        self.cams_dev = torch.tensor(rows, dtype=torch.float64, device=dev)
        engine.problem.cams_dev = self.cams_dev.data_ptr()
        self.cam_best = torch.zeros(12 * n, dtype=torch.float64, device=dev)
        self.any_improved = torch.zeros((), dtype=torch.float64, device=dev)
        self.host = self.cam_params.cpu().numpy().copy()
        self.best = {}
        self._graph = None
        self.graph_replays = 0
        self.use_graph = os.environ.get('MC3D_JOINT_GRAPH', '1') != '0'

    def _one_step(self):
        """One optimiser step, entirely on the current stream (no host synchronisation: capturable)."""
        torch = _torch()
        eng = self.engine
        st = eng._stream()
        costgrad = getattr(self.lib, f'mc3d_extrinsic_costgrad_{self.tag}')
        joint = getattr(self.lib, f'mc3d_extrinsic_joint_step_{self.tag}')
        eng.phases.phase(eng.problem, 0, eng.step, True, st)
        eng.phases.phase(eng.problem, 1, eng.step, True, st)
        for q, pb in enumerate(self.problems):
            if self.slots[q] is not None:                          # learnt but not part of the likelihood: no gradient reaches it
                _lib.check(costgrad(ctypes.byref(pb), st))
        _lib.check(joint(eng.ctrl.data_ptr(), self.cam_ctrl.data_ptr(), self.cam_params.data_ptr(), len(self.ids), eng.step,
                         self.lr, self.betas[0], self.betas[1], 1e-8, st))
        eng.phases.phase(eng.problem, 2, eng.step, True, st)      # the trajectory moves with the cameras of this step's gradient
        for q, slot in enumerate(self.slots):                      # the next step projects with the new R, T
            if slot is not None:
                self.cams_dev[slot, 9:21].copy_(self.cam_params[36 * q:36 * q + 12])
        eng.step += 1
        # best cameras: kept when the step just taken improved the running mean (the flag of the state entering the next step)
        improved = eng.ctrl[_lib.CT_STATE + 16 * (eng.step & 1) + 7] != 0
        rt = self.cam_params.view(-1, 36)[:, :12].reshape(-1)
        self.cam_best.copy_(torch.where(improved, rt, self.cam_best))
        self.any_improved.copy_(torch.maximum(self.any_improved, improved.to(torch.float64)))

    def steps(self, n):
        torch = _torch()
        eng = self.engine
        n = int(n)
        done = 0
        if self.use_graph and n >= 8:
            while done < 2 or (eng.step & 1):                      # eager warm-up, even parity for the captured pair
                self._one_step()
                done += 1
            if self._graph is None:
                first = eng.step
                try:
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                        self._one_step()
                        self._one_step()
                    self._graph = graph
                except Exception:                                  # capture unsupported here: stay eager
                    self.use_graph = False
                eng.step = first                                   # capturing did not execute anything
                torch.cuda.synchronize()
            if self._graph is not None:
                for _ in range((n - done) // 2):
                    self._graph.replay()
                    self.graph_replays += 1
                    eng.step += 2
                    done += 2
        for _ in range(n - done):
            self._one_step()
        # one read-back per chunk of iterations (a stopped run makes every further step a no-op on the device)
        self.host = self.cam_params.cpu().numpy().copy()
        if float(self.any_improved.item()) != 0.0:
            best = self.cam_best.cpu().numpy()
            dt = self.opt.torch_dtype
            self.best = {ID: (torch.tensor(best[12 * q:12 * q + 9].reshape(3, 3), dtype=dt),
                              torch.tensor(best[12 * q + 9:12 * q + 12].reshape(3, 1), dtype=dt)) for q, ID in enumerate(self.ids)}

    def _tensors(self, q):
        torch = _torch()
        dt = self.opt.torch_dtype
        return (torch.tensor(self.host[36 * q:36 * q + 9].reshape(3, 3), dtype=dt),
                torch.tensor(self.host[36 * q + 9:36 * q + 12].reshape(3, 1), dtype=dt))

    def write_back(self):
        self._graph = None
        for q, ID in enumerate(self.ids):
            self.opt.decomposed_cam_params[ID][1], self.opt.decomposed_cam_params[ID][2] = self._tensors(q)


def _running_means(costs, steps_per_iter):
    """The reference's per-iteration statistic (pose_refinement.py:1069-1071): after each iteration the mean of the
    list holding every step cost so far AND every earlier mean is appended to that same list (quirk Q5)."""
    means, s, n = [], 0.0, 0
    for i in range(len(costs) // steps_per_iter):
        for c in costs[i * steps_per_iter:(i + 1) * steps_per_iter]:
            s += float(c)
            n += 1
        mean = s / n
        means.append(mean)
        s += mean
        n += 1
    return means


def _interleave_history(costs, steps_per_iter, torch_dtype):
    """[cost..., running mean, cost..., running mean, ...] with upstream's element types: 0-d tensors for the step
    costs, numpy scalars for the means."""
    torch = _torch()
    np_dtype = np.float32 if torch_dtype == torch.float32 else np.float64
    out = []
    means = _running_means(costs, steps_per_iter)
    for i, mean in enumerate(means):
        for c in costs[i * steps_per_iter:(i + 1) * steps_per_iter]:
            out.append(torch.tensor(float(c), dtype=torch_dtype))
        out.append(np_dtype(mean))
    return out


# ---- command line (pose_refinement.py:1099-1256) ------------------------------------------------------------------------
def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--run_path', type=str)
    parser.add_argument('--refinement_types', nargs='+', default=['linear_interpolation'])
    parser.add_argument('--recording_log', type=str)
    parser.add_argument('--heatmaps_2d', type=str)
    parser.add_argument('--kpts_2d', type=str)
    parser.add_argument('--kpts_3d', type=str)
    parser.add_argument('--model', type=str)
    parser.add_argument('--save_path', type=str)
    parser.add_argument('--extrinsic_params_dir', type=str)
    parser.add_argument('--intrinsic_params_dir', type=str)
    parser.add_argument('--refinement_params_yaml', type=str)
    parser.add_argument('--body_part_lengths_yaml', type=str)
    parser.add_argument('--body_part_lengths_individual_name_yaml', default='my_lengths', type=str)
    parser.add_argument('--ignore_body_lengths', action='store_true')
    parser.add_argument('--interpolate_before_SGD', action='store_true')
    args = parser.parse_args(argv)
    torch = _torch()

    if args.run_path is None:
        args.run_path = os.getcwd()
    if args.save_path is None:
        args.save_path = args.run_path
    if args.extrinsic_params_dir is None:
        args.extrinsic_params_dir = Path(args.run_path).parent.parent / 'extrinsic_camera_parameters'
    if args.intrinsic_params_dir is None:
        args.intrinsic_params_dir = os.path.join(os.getcwd(), 'intrinsic_camera_parameters')

    log = {}
    if args.recording_log is not None:
        with open(args.recording_log) as fh:
            log = yaml.safe_load(fh)
    elif os.path.exists(os.path.join(args.run_path, 'recording_log.yaml')):
        with open(os.path.join(args.run_path, 'recording_log.yaml')) as fh:
            log = yaml.safe_load(fh)
    args.recording_log = log
    for name, value in vars(args).items():          # unset flags come from the recording log (:1142-1144)
        if value is None and name in log:
            setattr(args, name, log[name])

    kpts_3d = utils.load_if_exists(args.kpts_3d)
    utils.load_if_exists(args.kpts_2d) if args.kpts_2d else None
    heatmaps = utils.load_if_exists(args.heatmaps_2d) if args.heatmaps_2d else None
    refinement_types = set(args.refinement_types)
    params = utils.load_config(args.refinement_params_yaml)

    kpts_3d_interpolation = None
    if 'linear_interpolation' in refinement_types or args.interpolate_before_SGD:
        kwargs = utils.prepare_kwargs(linear_interpolation, params.get('linear_interpolation'))
        kpts_3d_interpolation = linear_interpolation(kpts_3d, **kwargs)
    if 'linear_interpolation' in refinement_types:
        out = os.path.join(args.save_path, 'kpts_3d_linear_interpolation.npy')
        print(f'saving linear interpolation at {out}')
        np.save(out, kpts_3d_interpolation)
        refinement_types.remove('linear_interpolation')

    if 'SGD' in refinement_types:
        with open(os.path.join(args.extrinsic_params_dir, 'camera_names.pkl'), 'rb') as fh:
            cameras, _origin_camera = pk.load(fh)
        decomposed = {}
        for i in cameras.keys():
            _, decomposed[i] = utils.get_params_from_name(cameras[i], intrinsic_params_dir=args.intrinsic_params_dir,
                                                          extrinsic_params_dir=args.extrinsic_params_dir)
        my_lengths = None
        if not args.ignore_body_lengths:
            if args.body_part_lengths_yaml is None and os.path.exists('./body_part_lengths.yaml'):
                args.body_part_lengths_yaml = './body_part_lengths.yaml'
            if args.body_part_lengths_yaml is not None:
                with open(args.body_part_lengths_yaml) as fh:
                    my_lengths = yaml.safe_load(fh)[args.body_part_lengths_individual_name_yaml]
        if args.interpolate_before_SGD:
            kpts_3d = kpts_3d_interpolation
        opt = Optimized_3d_Pose_Estimation(torch.tensor(heatmaps), kpts_3d, decomposed_cam_params_initial=decomposed,
                                           body_lengths=my_lengths)
        kwargs = utils.prepare_kwargs(opt.sgd_optimize, params.get('SGD'))
        opt.sgd_optimize(**kwargs)
        if my_lengths is not None:
            for title, traj in (("mean and standardized error of initial trajectory's body part lengths", opt.initial_trajectory),
                                ("mean and standardized error of the estimated trajectory's' body part lengths", opt.best_trajectory)):
                print(title)
                lengths = utils.get_body_part_lengths(traj)
                for bp in my_lengths:
                    print('; '.join([bp, str(torch.mean(lengths[bp])), str(torch.std(lengths[bp]))]))
        out = os.path.join(args.save_path, 'kpts_3d_SGD.npy')
        print(f'saving SGD at {out}')
        np.save(out, np.array(opt.best_trajectory))
        refinement_types.remove('SGD')


if __name__ == '__main__':
    main()
