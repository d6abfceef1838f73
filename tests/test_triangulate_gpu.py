"""GPU parity: csrc/triangulate.cu (through the C ABI) against the oracle and the reference goldens.

Tolerances (BASELINE.json north_star): fp64 path 1e-9 relative; fp32 path 1e-3 mm.
"relative" = per-joint ||X_gpu - X_ref|| / ||X_ref||.
"""
import numpy as np
import pytest

from conftest import cams_from_golden, load_golden, rel_err
from oracle import dlt as O

pytestmark = pytest.mark.gpu

FP64_RTOL = 1e-9
FP32_ATOL_MM = 1e-3


@pytest.fixture(scope='module')
def tri():
    import __graft_entry__ as g
    g.build()
    from mc3d_b200.triangulation import triangulate_multiview
    return triangulate_multiview


def _cuda(a):
    import torch
    return torch.tensor(a, device='cuda:0')


# ---- against outputs of the unmodified reference ------------------------------------------------------
def test_DLT_matches_reference_golden():
    import mc3d_b200.utils as u
    g = load_golden('dlt_stereo.npz')
    P, pts = g['P'], g['pts']
    for i in range(12):                                   # the per-point drop-in call
        got = u.DLT(P[0], P[1], pts[i][:, 0], pts[i][:, 1])
        assert got.shape == (3,) and got.dtype == np.float64
        assert rel_err(got[None], g['dlt'][i][None]).max() < FP64_RTOL


def test_batched_two_view_matches_reference_DLT(tri):
    g = load_golden('dlt_stereo.npz')
    kp = np.concatenate([np.transpose(g['pts'], (0, 2, 1)), np.ones((g['pts'].shape[0], 2, 1))], axis=2)
    got = tri(_cuda(kp), g['P']).cpu().numpy()
    assert rel_err(got, g['dlt']).max() < FP64_RTOL
    got_h = tri(kp, g['P'])                               # host pipeline
    assert np.array_equal(got, got_h)


def test_triangulate_points_matches_reference_golden():
    import torch
    import mc3d_b200.utils as u
    g = load_golden('dlt_stereo.npz')
    c = cams_from_golden(g, 2)
    args = (c[0][0], c[0][3], c[0][1], c[0][2], c[1][0], c[1][3], c[1][1], c[1][2])
    got = u.triangulate_points(g['pair'].reshape(8, 17, 2, 2), *args)
    assert got.shape == (8, 17, 3) and got.dtype == np.float64
    assert rel_err(got, g['tri']).max() < FP64_RTOL
    got_t = u.triangulate_points(torch.tensor(g['pair']), *[torch.tensor(a) for a in args])   # torch inputs (utils.py:1294)
    assert got_t.shape == (136, 3) and rel_err(got_t, g['tri']).max() < FP64_RTOL
    one = u.triangulate_points(g['pair'][5], *args)       # a single point, as get_pose_3D calls it upstream
    assert one.shape == (3,) and rel_err(one[None], g['tri'].reshape(-1, 3)[5][None]).max() < FP64_RTOL


@pytest.mark.parametrize('tag,n', [('c2', 2), ('c3', 3)])
def test_get_pose_3D_matches_reference_golden(tag, n):
    import mc3d_b200.pose_estimation as pe
    g = load_golden(f'pose3d_{tag}.npz')
    cams = cams_from_golden(g, n)
    kp = list(g['kpts'])
    got = pe.get_pose_3D(cams, kp)
    assert got.shape == g['p3d'].shape and got.dtype == np.float64
    assert rel_err(got, g['p3d']).max() < FP64_RTOL
    assert rel_err(pe.get_pose_3D(cams, kp, ignore_nonlinear_distortions=True), g['p3d_nodist']).max() < FP64_RTOL
    assert rel_err(pe.get_pose_3D(cams, kp, world_trans_rot=(g['Rw'], np.zeros(3))), g['p3d_world']).max() < FP64_RTOL
    # inputs are not modified
    assert np.array_equal(np.array(kp), g['kpts'])


def test_get_pose_3D_config1_full_size_matches_reference():
    """BASELINE.json configs[0] at its full size -- 2-camera COCO-17 triangulation of 400 frames -- against the output of
    the reference's own get_pose_3D on the same keypoints (tests/golden/pose3d_config1.npz); one kernel launch here."""
    import mc3d_b200
    import mc3d_b200.pose_estimation as pe
    g = load_golden('pose3d_config1.npz')
    cams = cams_from_golden(g, 2)
    kp = list(g['kpts'].astype(np.float64))
    before = mc3d_b200.launch_count()
    got = pe.get_pose_3D(cams, kp)
    assert mc3d_b200.launch_count() - before == 1
    assert got.shape == (400, 17, 3) and got.dtype == np.float64
    assert rel_err(got, g['p3d']).max() < FP64_RTOL


def test_get_pose_3D_camera_subset_and_no_scores(syn):
    import mc3d_b200.pose_estimation as pe
    rng = np.random.default_rng(3)
    cams = syn.ring_rig(4, distortion=True)
    X = syn.smooth_trajectory(5, 17, rng, centre=(0, 0, 3000.0))
    kp = syn.keypoints_from_trajectory(X, cams, rng)
    for ci in (None, [0, 1, 2], [0, 1]):
        ref = O.get_pose_3d(cams, list(kp), camera_indices=ci)
        assert rel_err(pe.get_pose_3D(cams, list(kp), camera_indices=ci), ref).max() < FP64_RTOL
    ref = O.get_pose_3d(cams, list(kp[:, :, :2, :]))
    assert rel_err(pe.get_pose_3D(cams, list(kp[:, :, :2, :])), ref).max() < FP64_RTOL


# ---- V-view weighted DLT against the oracle ---------------------------------------------------------------
@pytest.mark.parametrize('n_views', [3, 4, 5, 8, 11, 16])
def test_weighted_fp64(tri, syn, n_views):
    kp, P, _, _ = syn.multiview_points(20000, n_views, seed=n_views)
    got = tri(_cuda(kp), P).cpu().numpy()
    polished = O.dlt_weighted_polished(kp, P)
    assert rel_err(got, polished).max() < FP64_RTOL
    # against the raw LAPACK answer: its own ~1e-9 noise on mm-scale rigs is the only excess
    r = rel_err(got, O.dlt_weighted(kp, P))
    assert np.mean(r < FP64_RTOL) > 0.999 and r.max() < 1e-8
    assert rel_err(got, polished).max() <= rel_err(O.dlt_weighted(kp, P), polished).max() + 1e-13


def test_weighted_fp64_stereo_rig(tri, syn):
    rng = np.random.default_rng(8)
    cams = syn.stereo_rig(distortion=False)
    X = syn.SCENE_CENTRE + rng.normal(0, 400, size=(20000, 3))
    kp = np.empty((20000, 2, 3))
    for c in range(2):
        kp[:, c, :2] = syn.project(X, cams[c], distort=False) + rng.normal(0, 1, size=(20000, 2))
    kp[:, :, 2] = rng.uniform(0.2, 1, size=(20000, 2))
    P = syn.projection_matrices(cams)
    got = tri(_cuda(kp), P).cpu().numpy()
    assert rel_err(got, O.dlt_weighted(kp, P)).max() < FP64_RTOL
    assert rel_err(got, O.dlt_weighted_polished(kp, P)).max() < FP64_RTOL


@pytest.mark.parametrize('n_views', [2, 3, 8, 16])
def test_weighted_fp32(tri, syn, n_views):
    if n_views == 2:
        rng = np.random.default_rng(9)
        cams = syn.stereo_rig(distortion=False)
        X = syn.SCENE_CENTRE + rng.normal(0, 400, size=(20000, 3))
        kp = np.empty((20000, 2, 3))
        for c in range(2):
            kp[:, c, :2] = syn.project(X, cams[c], distort=False) + rng.normal(0, 1, size=(20000, 2))
        kp[:, :, 2] = rng.uniform(0.2, 1, size=(20000, 2))
        P = syn.projection_matrices(cams)
    else:
        kp, P, _, _ = syn.multiview_points(20000, n_views, seed=20 + n_views)
    kp32 = kp.astype(np.float32)
    got = tri(_cuda(kp32), P)
    assert str(got.dtype) == 'torch.float32'
    ref = O.dlt_weighted_polished(kp32.astype(np.float64), P)       # same (rounded) inputs, float64 maths
    err = np.linalg.norm(got.cpu().numpy().astype(np.float64) - ref, axis=1)
    assert err.max() < FP32_ATOL_MM, err.max()
    # float storage + double arithmetic: the error is one output rounding
    assert np.array_equal(got.cpu().numpy(), ref.astype(np.float32)) or err.max() < 5e-4


def test_fp32_mixed_solver_agrees_with_all_double_solver(tri, syn):
    """float storage: the mixed-precision solver (float A^T A + double residual Newton) against the all-double
    solver on the same inputs -- both are one output rounding away from the exact minimiser."""
    from mc3d_b200 import _lib
    for n_views in (2, 3, 4, 8, 16):
        kp, P, _, _ = syn.multiview_points(30000, n_views, seed=60 + n_views)
        if n_views == 2:
            cams = syn.stereo_rig(distortion=False)
            P = syn.projection_matrices(cams)
            rng = np.random.default_rng(3)
            X = syn.SCENE_CENTRE + rng.normal(0, 400, size=(30000, 3))
            for c in range(2):
                kp[:, c, :2] = syn.project(X, cams[c], distort=False) + rng.normal(0, 1, size=(30000, 2))
        kp32 = _cuda(kp.astype(np.float32))
        a = tri(kp32, P).cpu().numpy().astype(np.float64)
        b = tri(kp32, P, flags=_lib.TRI_FLAG_FP64).cpu().numpy().astype(np.float64)
        assert np.linalg.norm(a - b, axis=1).max() < 5e-4            # at most one float ulp per coordinate at ~3 m
        assert np.mean(np.all(a == b, axis=1)) > 0.85
        ref = O.dlt_weighted_polished(kp.astype(np.float32).astype(np.float64), P)
        assert np.linalg.norm(a - ref, axis=1).max() < FP32_ATOL_MM


def test_fp32_starting_subset_views_missing_or_wild(tri, syn):
    """The float kernel starts from a DLT over views {0, V/3, 2V/3} (triangulate.cu solve_merged).  Zero-weight
    views there (with wild pixels), outliers that put the start far from the answer, and a subset reduced to one
    usable view must all end at the same minimiser as the float64 oracle."""
    kp, P, _, _ = syn.multiview_points(20000, 8, seed=77)
    rng = np.random.default_rng(5)
    kp[0:4000, 0, 2] = 0.0                                   # view 0 unused, pixel left as is
    kp[4000:8000, 2, 2] = 0.0
    kp[4000:8000, 2, :2] = rng.uniform(-5000, 5000, size=(4000, 2))      # unused view with a wild pixel
    kp[8000:12000, 5, :2] += rng.normal(0, 60.0, size=(4000, 2))           # gross outlier in a starting view (weighted in)
    kp[12000:16000, [0, 2], 2] = 0.0                         # only one starting view usable -> 2 unknown-rank start
    kp[16000:18000, [0, 2, 5], 2] = 0.0                      # no starting view usable
    kp32 = kp.astype(np.float32)
    got = tri(_cuda(kp32), P).cpu().numpy().astype(np.float64)
    ref = O.dlt_weighted_polished(kp32.astype(np.float64), P)
    err = np.linalg.norm(got - ref, axis=1)
    assert np.isfinite(got).all()
    assert err.max() < FP32_ATOL_MM, (err.max(), int(err.argmax()))


@pytest.mark.parametrize('frac', [0.05, 0.2, 0.5])
def test_fp32_random_unusable_views_agree_with_all_double_solver(tri, syn, frac):
    """A fraction of ALL views unusable (weight 0, wild pixel), at random: joints go through the cascade of planned start pairs,
    the start from their own first two usable views (two or three views left) and, with nearly collinear rays, the double
    fallback.  Same NaN pattern as the all-double solver (fewer than two usable views -> NaN) and within the float32 bar."""
    import torch
    from mc3d_b200 import _lib
    n = 1_500_000
    kp, P, _, _ = syn.multiview_points(n, 8, seed=91)
    kp32 = torch.tensor(kp.astype(np.float32), device='cuda:0')
    gen = torch.Generator(device='cuda:0').manual_seed(int(frac * 100))
    bad = torch.rand((n, 8), device='cuda:0', generator=gen) < frac
    kp32[..., 2][bad] = 0.0
    kp32[..., 0][bad] = 5000.0 * torch.rand((int(bad.sum()),), device='cuda:0', generator=gen)
    kp32[..., 1][bad] = 5000.0 * torch.rand((int(bad.sum()),), device='cuda:0', generator=gen)
    got = tri(kp32, P)
    ref = tri(kp32, P, flags=_lib.TRI_FLAG_FP64)
    assert bool((torch.isnan(got) == torch.isnan(ref)).all())
    ok = torch.isfinite(ref).all(dim=1)
    n_views = (kp32[..., 2] > 0).sum(dim=1)
    assert bool((ok == (n_views >= 2))[n_views != 2].all())          # (two views can still be refused as collinear)
    err = (got[ok].double() - ref[ok].double()).norm(dim=1)
    assert float(err.max()) < FP32_ATOL_MM, float(err.max())


def test_fp32_points_near_world_origin(tri, syn):
    rng = np.random.default_rng(34)
    cams = syn.ring_rig(8, centre=(0.0, 0.0, 0.0))
    P = syn.projection_matrices(cams)
    X = rng.normal(0, 30.0, size=(20000, 3))
    kp = np.empty((20000, 8, 3))
    for c in range(8):
        kp[:, c, :2] = syn.project(X, cams[c], distort=False) + rng.normal(0, 1, size=(20000, 2))
    kp[:, :, 2] = rng.uniform(0.2, 1, size=(20000, 8))
    kp32 = kp.astype(np.float32)
    got = tri(_cuda(kp32), P).cpu().numpy().astype(np.float64)
    pol = O.dlt_weighted_polished(kp32.astype(np.float64), P, iters=8)
    assert np.linalg.norm(got - pol, axis=1).max() < FP32_ATOL_MM


def test_layouts_agree_bitwise(tri, syn):
    kp, P, _, _ = syn.multiview_points(3000, 8, seed=31)
    a = tri(_cuda(kp), P, layout='nv3')
    b = tri(_cuda(np.ascontiguousarray(np.transpose(kp, (0, 2, 1)))), P, layout='n3v')
    assert bool((a == b).all())
    a32 = tri(_cuda(kp.astype(np.float32)), P, layout='nv3')
    b32 = tri(_cuda(np.ascontiguousarray(np.transpose(kp, (0, 2, 1))).astype(np.float32)), P, layout='n3v')
    assert bool((a32 == b32).all())
    kp4 = kp.reshape(30, 100, 8, 3)                         # leading dims are kept
    assert tuple(tri(_cuda(kp4), P).shape) == (30, 100, 3)


def test_jacobi_solver_agrees_with_newton(tri, syn):
    from mc3d_b200 import _lib
    kp, P, _, _ = syn.multiview_points(20000, 8, seed=32)
    a = tri(_cuda(kp), P).cpu().numpy()
    j = tri(_cuda(kp), P, flags=_lib.TRI_FLAG_JACOBI).cpu().numpy()
    pol = O.dlt_weighted_polished(kp, P)
    assert rel_err(j, pol).max() < FP64_RTOL
    assert rel_err(a, j).max() < FP64_RTOL


def test_points_near_world_origin_need_the_eigen_shift(tri, syn):
    """Near the world origin lambda_min(B) is not negligible against M: the secular iteration must take
    extra steps (or fall back) and still agree with the oracle."""
    rng = np.random.default_rng(33)
    cams = syn.ring_rig(6, centre=(0.0, 0.0, 0.0))
    P = syn.projection_matrices(cams)
    X = rng.normal(0, 30.0, size=(20000, 3))
    kp = np.empty((20000, 6, 3))
    for c in range(6):
        kp[:, c, :2] = syn.project(X, cams[c], distort=False) + rng.normal(0, 1, size=(20000, 2))
    kp[:, :, 2] = rng.uniform(0.2, 1, size=(20000, 6))
    got = tri(_cuda(kp), P).cpu().numpy()
    pol = O.dlt_weighted_polished(kp, P, iters=8)
    # these joints are ill-conditioned by construction (two close eigenvalues); compare in absolute mm
    assert np.linalg.norm(got - pol, axis=1).max() < 1e-6


@pytest.mark.parametrize('n', [1, 2, 255, 256, 257, 1000, 70001])
def test_ragged_sizes(tri, syn, n):
    kp, P, _, _ = syn.multiview_points(n, 4, seed=40)
    got = tri(_cuda(kp), P).cpu().numpy()
    assert got.shape == (n, 3)
    assert rel_err(got, O.dlt_weighted_polished(kp, P)).max() < FP64_RTOL
    got32 = tri(_cuda(kp.astype(np.float32)), P).cpu().numpy()
    assert np.linalg.norm(got32 - got, axis=1).max() < FP32_ATOL_MM


def test_empty_input(tri, syn):
    import torch
    _, P, _, _ = syn.multiview_points(1, 4, seed=41)
    out = tri(torch.empty((0, 4, 3), dtype=torch.float64, device='cuda:0'), P)
    assert tuple(out.shape) == (0, 3)
    assert tri(np.empty((0, 4, 3)), P).shape == (0, 3)


def test_view_permutation_invariance(tri, syn):
    kp, P, _, _ = syn.multiview_points(5000, 8, seed=42)
    perm = np.random.default_rng(0).permutation(8)
    a = tri(_cuda(kp), P).cpu().numpy()
    b = tri(_cuda(np.ascontiguousarray(kp[:, perm])), P[perm]).cpu().numpy()
    assert rel_err(a, b).max() < 1e-11


def test_weight_scale_invariance(tri, syn):
    kp, P, _, _ = syn.multiview_points(5000, 8, seed=43)
    kp2 = kp.copy()
    kp2[:, :, 2] *= 4.0                                     # power of two: B scales exactly by 16
    a = tri(_cuda(kp), P).cpu().numpy()
    b = tri(_cuda(kp2), P).cpu().numpy()
    assert rel_err(a, b).max() < 1e-12


def test_degenerate_and_nonfinite_joints_give_nan_without_disturbing_neighbours(tri, syn):
    kp, P, _, _ = syn.multiview_points(600, 4, seed=44)
    ref = tri(_cuda(kp), P).cpu().numpy()
    bad = kp.copy()
    bad[10, :, 2] = 0.0                                     # no usable view
    bad[11, 1:, 2] = 0.0                                    # one usable view
    bad[300, 2, 0] = np.nan
    bad[301, 0, 1] = np.inf
    got = tri(_cuda(bad), P).cpu().numpy()
    for i in (10, 11, 300, 301):
        assert np.isnan(got[i]).all()
    keep = np.setdiff1d(np.arange(600), [10, 11, 300, 301])
    assert np.array_equal(got[keep], ref[keep])
    # a zero weight removes exactly that view
    z = kp.copy()
    z[:, 3, 2] = 0.0
    assert rel_err(tri(_cuda(z), P).cpu().numpy(), tri(_cuda(np.ascontiguousarray(kp[:, :3])), P[:3]).cpu().numpy()).max() < 1e-12


def test_top2_selection_semantics(tri, syn):
    """argsort(conf)[-2:] (pose_estimation.py:35-37): ties go to the higher index."""
    rng = np.random.default_rng(45)
    cams = syn.ring_rig(5, distortion=True)
    X = syn.smooth_trajectory(40, 17, rng, centre=(0, 0, 3000.0))
    kp = syn.keypoints_from_trajectory(X, cams, rng)        # (T, J, 3, C)
    kp[:, :, 2, :] = np.round(kp[:, :, 2, :] * 4) / 4       # many exact ties
    P = syn.projection_matrices(cams)
    K = np.stack([cams[i][0] for i in cams])
    D = np.stack([cams[i][3].reshape(-1) for i in cams])
    got = tri(_cuda(kp), P, K=K, dist=D, layout='n3v', mode='top2').cpu().numpy()
    ref = np.empty_like(got)
    for t in range(kp.shape[0]):
        for j in range(17):
            top = np.argsort(kp[t, j, 2, :], kind='stable')[-2:]
            pts = kp[t, j, :2, top]                         # (2 cams, 2 xy)
            c0, c1 = cams[top[0]], cams[top[1]]
            ref[t, j] = O.triangulate_points(pts, c0[0], c0[3], c0[1], c0[2], c1[0], c1[3], c1[1], c1[2])
    assert rel_err(got, ref).max() < FP64_RTOL


def test_host_pipeline_multi_chunk_equals_device_path(tri, syn):
    n = 3_000_000                                           # > one 64 MiB chunk for 2 views in fp32
    rng = np.random.default_rng(46)
    base, P, _, _ = syn.multiview_points(4096, 2, seed=46)
    cams = syn.stereo_rig(distortion=False)
    P = syn.projection_matrices(cams)
    X = syn.SCENE_CENTRE + rng.normal(0, 400, size=(4096, 3))
    for c in range(2):
        base[:, c, :2] = syn.project(X, cams[c], distort=False) + rng.normal(0, 1, size=(4096, 2))
    kp = np.tile(base.astype(np.float32), (n // 4096 + 1, 1, 1))[:n]
    kp[:, 0, 0] += (np.arange(n) % 97).astype(np.float32) * 0.01
    host = tri(kp, P)
    dev = tri(_cuda(kp), P).cpu().numpy()
    assert host.shape == (n, 3) and np.array_equal(host, dev)


def test_misaligned_and_cpu_tensors_are_rejected(tri, syn):
    import torch
    import mc3d_b200
    kp, P, _, _ = syn.multiview_points(16, 4, seed=47)
    with pytest.raises(mc3d_b200.Mc3dError):
        tri(torch.tensor(kp), P)                            # CPU tensor: no fallback
    with pytest.raises(ValueError):
        tri(_cuda(kp), P[:3])


def test_full_size_sample_check(tri, syn):
    """BASELINE config 2 shape at 1/10 scale (1.7e7 joints, 8 views, fp32): a random sample of joints is checked
    against the oracle, and the whole output is finite and inside the scene."""
    import torch
    n = 17_000_000
    g = torch.Generator(device='cuda:0').manual_seed(1)
    from bench import make_triangulation_workload
    kp, P = make_triangulation_workload(n, 8, torch.float32, 'cuda:0', seed=1)
    out = tri(kp, P)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all())
    idx = torch.randint(0, n, (50000,), device='cuda:0', generator=g)
    kps = kp[idx].cpu().numpy().astype(np.float64)
    ref = O.dlt_weighted_polished(kps, P)
    err = np.linalg.norm(out[idx].cpu().numpy().astype(np.float64) - ref, axis=1)
    assert err.max() < FP32_ATOL_MM
    # determinism: a second launch gives the same bits
    assert bool((tri(kp, P) == out).all())


def test_project_points_matches_reference_golden():
    """utils.project_points (the cv.projectPoints wrapper, utils.py:438-458) and compute_2d_coordinates (:558-567)."""
    import mc3d_b200.utils as u
    g = load_golden('dlt_stereo.npz')
    c = cams_from_golden(g, 2)
    for i in range(2):
        got = u.project_points(g['proj_pts'], c[i][0], c[i][1], c[i][2], c[i][3])
        assert got.shape == g[f'proj_cam{i}'].shape == (3, 17, 2)
        assert np.abs(got - g[f'proj_cam{i}']).max() < 1e-9
    flat = u.project_points(g['proj_pts'].reshape(-1, 3), c[1][0], c[1][1], c[1][2])
    assert flat.shape == (51, 2) and np.abs(flat - g['proj_flat_nodist']).max() < 1e-9
    assert np.abs(u.compute_2d_coordinates(g['P'][1], g['proj_pts'][0, 0]) - g['uv_c2d']).max() < 1e-9
