"""Tiny pass over every kernel for compute-sanitizer (one tool per run):

    compute-sanitizer --tool memcheck  python tests/sanitize_small.py
    compute-sanitizer --tool racecheck python tests/sanitize_small.py

Sizes are chosen to hit the TMA ring (several tiles per CTA), the ragged last tile, both decode kernels, the
persistent refinement kernel (grid barriers) and the graph of separate kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import __graft_entry__ as g
    g.build()
    import mc3d_b200.pose_refinement as pr
    from mc3d_b200 import synthetic as syn
    from mc3d_b200.decode import decode_heatmaps
    from mc3d_b200.interpolation import linear_interpolation
    from mc3d_b200.pose_estimation import get_pose_3D
    from mc3d_b200.triangulation import triangulate_multiview
    dev = 'cuda:0'
    for V, n in ((8, 5000), (16, 1500), (2, 700), (5, 300)):
        kp, P, _, _ = syn.multiview_points(n, V, seed=V)
        for dt in (np.float32, np.float64):
            out = triangulate_multiview(torch.tensor(kp.astype(dt), device=dev), P)
            assert bool(torch.isfinite(out).all())
    rng = np.random.default_rng(0)
    cams = syn.stereo_rig(distortion=True)
    X = syn.smooth_trajectory(12, 17, rng, centre=(0, 0, 3000.0))
    kp = syn.keypoints_from_trajectory(X, cams, rng)
    assert np.isfinite(get_pose_3D(cams, list(kp))).all()
    hm, _ = syn.gaussian_blob_heatmaps(70, seed=1)
    k1, m1 = decode_heatmaps(torch.tensor(hm, device=dev))
    hm2, _ = syn.gaussian_blob_heatmaps(9, H=40, W=36, seed=2)
    k2, m2 = decode_heatmaps(torch.tensor(hm2, device=dev))
    assert bool(torch.isfinite(m1).all()) and bool(torch.isfinite(m2).all())
    gs, init, cams2, _ = syn.refinement_inputs(64, seed=7)
    for env in ({}, {'MC3D_REFINE_FUSED': '0'}, {'MC3D_REFINE_TWO_PHASE': '0'}, {'MC3D_REFINE_PEER': '0'}):
        os.environ.update(env)
        for dt in (torch.float32, torch.float64):
            opt = pr.Optimized_3d_Pose_Estimation(gs, init, decomposed_cam_params_initial={i: list(cams2[i]) for i in cams2},
                                                  body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=dt, device=dev)
            opt.sgd_optimize(print_frequency=np.inf, lr=0.01, lambda_smooth=1e-3, lambda_body_length=1.0, max_iter=9, time_interval=[0, 64])
            assert np.isfinite(np.array(opt.trajectory)).all()
        for k in env:
            del os.environ[k]
    out = linear_interpolation(np.asarray(init[:40]))
    assert np.isfinite(out).all()
    print('sanitize_small ok')


if __name__ == '__main__':
    main()
