"""CPU: the host logic of the triangulation shims (argument marshalling, layouts, modes, camera ordering, result
containers) without a GPU.  ``mc3d_b200.triangulation.triangulate_multiview`` -- the one function that talks to the
library -- is replaced by a stand-in built on the ORACLE (test infrastructure; allowed in tests only), and the shims
around it are held to the reference goldens.  What runs on the GPU is tested in tests/test_triangulate_gpu.py."""
import numpy as np
import pytest

from conftest import cams_from_golden, load_golden, rel_err
from oracle import dlt as O


def _oracle_triangulate(kpts, P, K=None, dist=None, layout='nv3', mode='weighted', out=None, flags=0, device=0):
    """Same contract as triangulate_multiview for numpy input, arithmetic from oracle/dlt.py."""
    kp = np.asarray(kpts, dtype=np.float64)
    if layout == 'n3v':
        kp = np.swapaxes(kp, -1, -2)
    lead = kp.shape[:-2]
    kp = kp.reshape((-1,) + kp.shape[-2:]).copy()                      # (N, V, 3)
    P = np.asarray(P, dtype=np.float64).reshape(-1, 3, 4)
    assert P.shape[0] == kp.shape[1]
    if K is not None:
        K = np.asarray(K, dtype=np.float64).reshape(-1, 3, 3)
        dist = np.asarray(dist, dtype=np.float64).reshape(K.shape[0], -1)
        for v in range(kp.shape[1]):
            kp[:, v, :2] = O.undistort_points(kp[:, v, :2], K[v], dist[v])
    if mode == 'top2':
        score = np.where(np.isnan(kp[:, :, 2]), np.inf, kp[:, :, 2])
        order = np.argsort(score, axis=1, kind='stable')[:, -2:]     # ties -> higher index, NaN last (np.argsort semantics)
        res = np.empty((kp.shape[0], 3))
        for n in range(kp.shape[0]):
            a, b = order[n]
            pair = np.array([[[kp[n, a, 0], kp[n, a, 1], 1.0], [kp[n, b, 0], kp[n, b, 1], 1.0]]])
            res[n] = O.dlt_weighted_polished(pair, P[[a, b]])[0]
    else:
        res = O.dlt_weighted_polished(kp, P)
    res = res.reshape(lead + (3,))
    if out is not None:
        out[...] = res
        return out
    return res


@pytest.fixture()
def shims(monkeypatch):
    import mc3d_b200.pose_estimation as pe
    import mc3d_b200.triangulation as tri
    import mc3d_b200.utils as u
    monkeypatch.setattr(tri, 'triangulate_multiview', _oracle_triangulate)
    return u, pe


def test_DLT_and_triangulate_points_marshalling(shims):
    u, _ = shims
    g = load_golden('dlt_stereo.npz')
    P, pts = g['P'], g['pts']
    got = np.array([u.DLT(P[0], P[1], p[:, 0], p[:, 1]) for p in pts[:20]])
    assert got.shape == (20, 3) and got.dtype == np.float64 and rel_err(got, g['dlt'][:20]).max() < 1e-9
    c = cams_from_golden(g, 2)
    args = (c[0][0], c[0][3], c[0][1], c[0][2], c[1][0], c[1][3], c[1][1], c[1][2])
    tri = u.triangulate_points(g['pair'].reshape(8, 17, 2, 2), *args)
    assert tri.shape == (8, 17, 3) and tri.dtype == np.float64 and rel_err(tri, g['tri']).max() < 1e-9
    import torch
    tri_t = u.triangulate_points(torch.tensor(g['pair']), *[torch.tensor(a) for a in args])      # torch inputs (utils.py:1294)
    assert isinstance(tri_t, np.ndarray) and rel_err(tri_t, g['tri'].reshape(-1, 3)).max() < 1e-9
    one = u.triangulate_points(g['pair'][5], *args)                                               # a single point
    assert one.shape == (3,)


@pytest.mark.parametrize('tag,n', [('c2', 2), ('c3', 3)])
def test_get_pose_3D_marshalling(shims, tag, n):
    _, pe = shims
    g = load_golden(f'pose3d_{tag}.npz')
    cams = cams_from_golden(g, n)
    kp = list(g['kpts'])
    assert rel_err(pe.get_pose_3D(cams, kp), g['p3d']).max() < 1e-9
    assert rel_err(pe.get_pose_3D(cams, kp, ignore_nonlinear_distortions=True), g['p3d_nodist']).max() < 1e-9
    assert rel_err(pe.get_pose_3D(cams, kp, world_trans_rot=(g['Rw'], np.zeros(3))), g['p3d_world']).max() < 1e-9
    assert np.array_equal(np.array(kp), g['kpts'])                    # inputs untouched
    if n == 3:                                                        # a camera subset, and scoreless (J, 2, C) input
        ref = O.get_pose_3d(cams, kp, camera_indices=[0, 1])
        assert rel_err(pe.get_pose_3D(cams, kp, camera_indices=[0, 1]), ref).max() < 1e-9
        ref = O.get_pose_3d(cams, [k[:, :2, :] for k in kp])
        assert rel_err(pe.get_pose_3D(cams, [k[:, :2, :] for k in kp]), ref).max() < 1e-9
    with pytest.raises(ValueError):
        pe.get_pose_3D(cams, kp, camera_indices=[0])


def test_get_pose_3D_config1_full_size(shims):
    _, pe = shims
    g = load_golden('pose3d_config1.npz')
    got = pe.get_pose_3D(cams_from_golden(g, 2), list(g['kpts'].astype(np.float64)))
    assert got.shape == (400, 17, 3) and rel_err(got, g['p3d']).max() < 1e-9


# ---- PoseEstimator.get_heatmap_means_cov / _stds: recursion, in-place thresholding, containers --------------------------
def _oracle_decode(heatmaps, threshold=0.01, want_kpts=True, want_moments=True, **kw):
    from oracle import decode as D
    hm = np.asarray(heatmaps, dtype=np.float32)
    lead = hm.shape[:-2]
    flat = hm.reshape((-1,) + hm.shape[-2:])
    mom = D.heatmap_means_cov(flat.copy(), threshold=threshold, mutate=False).reshape(lead + (6,)) if want_moments else None
    kp = None
    if want_kpts:
        xy, sc = D.argmax_decode(flat)
        kp = np.concatenate([xy, sc[:, None]], axis=1).astype(np.float32).reshape(lead + (3,))
    return kp, mom


def test_heatmap_moment_shims(monkeypatch):
    import torch
    import mc3d_b200.decode as dec
    from mc3d_b200.mmpose_pose_estimation import PoseEstimator
    monkeypatch.setattr(dec, 'decode_heatmaps', _oracle_decode)
    g = load_golden('heatmap_moments.npz')
    expect = g['heatmaps'].copy()
    expect[expect < 0.01] = 0
    hm = g['heatmaps'].copy()
    got = PoseEstimator.get_heatmap_means_cov(None, hm)
    assert got.shape == (17, 6) and got.dtype == np.float64 and np.abs(got - g['moments']).max() < 2e-5
    assert np.array_equal(hm, expect)                                 # the caller's array is thresholded in place (Q7)
    t = torch.tensor(g['heatmaps'].copy())
    got_t = PoseEstimator.get_heatmap_means_cov(None, t)
    assert np.abs(got_t - g['moments_torch_in']).max() < 2e-5 and np.array_equal(t.numpy(), expect)
    lst = PoseEstimator.get_heatmap_means_cov(None, [g['heatmaps'].copy(), g['heatmaps'].copy()])
    assert lst.shape == (2, 17, 6) and np.array_equal(lst[0], lst[1])
    means, stds = PoseEstimator.get_heatmap_means_stds(expect.copy())
    assert np.abs(np.array(means) - g['means']).max() < 2e-5 and np.abs(np.array(stds) - g['stds']).max() < 2e-5
    with pytest.raises(NotImplementedError):
        PoseEstimator('det.py', 'det.pth', 'pose.py', 'pose.pth')


def test_default_device_resolution(monkeypatch):
    """The host-buffer pipelines run on the process's own GPU (`device=0` sent every torchrun rank to GPU 0):
    explicit argument > torch's current device (when CUDA is initialised) > LOCAL_RANK > 0."""
    import torch
    from mc3d_b200 import _lib
    assert _lib.default_device(3) == 3
    assert _lib.default_device(torch.device('cuda', 2)) == 2
    if not (torch.cuda.is_available() and torch.cuda.is_initialized()):
        monkeypatch.setenv('LOCAL_RANK', '5')
        assert _lib.default_device() == 5
        monkeypatch.delenv('LOCAL_RANK')
        assert _lib.default_device() == 0
