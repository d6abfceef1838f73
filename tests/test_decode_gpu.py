"""GPU parity: csrc/decode.cu against the reference golden (moments) and the oracle (moments + argmax)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import decode as D

pytestmark = pytest.mark.gpu

MOMENT_ATOL = 5e-5          # float32 heatmaps, values up to ~64: the reference's own numpy-vs-torch spread is 1e-5


@pytest.fixture(scope='module')
def dec():
    import __graft_entry__ as g
    g.build()
    from mc3d_b200.decode import decode_heatmaps
    return decode_heatmaps


def _cuda(a):
    import torch
    return torch.tensor(a, device='cuda:0')


def test_moments_match_reference_golden_and_mutate_input():
    from mc3d_b200.mmpose_pose_estimation import PoseEstimator
    g = load_golden('heatmap_moments.npz')
    hm = g['heatmaps'].copy()
    got = PoseEstimator.get_heatmap_means_cov(None, hm)
    assert got.shape == (17, 6) and got.dtype == np.float64
    assert np.abs(got - g['moments']).max() < MOMENT_ATOL
    assert np.array_equal(got[5], np.zeros(6))                     # sub-threshold map -> six zeros
    assert np.allclose(got[6], [7, 10, 0, 0, 0, 0], atol=1e-6)     # single pixel
    expect = g['heatmaps'].copy()
    expect[expect < 0.01] = 0
    assert np.array_equal(hm, expect)                              # upstream's in-place thresholding (Q7)
    # the float64 statement of the same moments is even closer
    assert np.abs(got - D.heatmap_means_cov_f64(g['heatmaps'])).max() < 2e-5


def test_moments_torch_inputs_cpu_and_cuda():
    import torch
    from mc3d_b200.mmpose_pose_estimation import PoseEstimator
    g = load_golden('heatmap_moments.npz')
    expect = g['heatmaps'].copy()
    expect[expect < 0.01] = 0
    t_cpu = torch.tensor(g['heatmaps'].copy())
    got = PoseEstimator.get_heatmap_means_cov(None, t_cpu)
    assert np.abs(got - g['moments_torch_in']).max() < MOMENT_ATOL
    assert np.array_equal(t_cpu.numpy(), expect)
    t_gpu = _cuda(g['heatmaps'].copy())
    got = PoseEstimator.get_heatmap_means_cov(None, t_gpu)
    assert np.abs(got - g['moments']).max() < MOMENT_ATOL
    assert np.array_equal(t_gpu.cpu().numpy(), expect)             # written back on the device
    lst = PoseEstimator.get_heatmap_means_cov(None, [g['heatmaps'].copy(), g['heatmaps'].copy()])
    assert lst.shape == (2, 17, 6)


def test_means_stds_match_reference_golden():
    from mc3d_b200.mmpose_pose_estimation import PoseEstimator
    g = load_golden('heatmap_moments.npz')
    hm = np.where(g['heatmaps'] < 0.01, 0, g['heatmaps']).astype(np.float32)
    means, stds = PoseEstimator.get_heatmap_means_stds(hm)
    assert np.abs(np.array(means) - g['means']).max() < MOMENT_ATOL
    assert np.abs(np.array(stds) - g['stds']).max() < MOMENT_ATOL


@pytest.mark.parametrize('shape', [(64, 48), (96, 72), (32, 32), (17, 23), (128, 96), (5, 4)])
@pytest.mark.parametrize('generic', [False, True, 'tma', 'no_stage'])
def test_decode_vs_oracle(dec, syn, shape, generic):
    H, W = shape
    hm, _ = syn.gaussian_blob_heatmaps(300, H=H, W=W, seed=H * W) if min(H, W) > 16 else \
        (np.random.default_rng(H).uniform(0, 1, size=(300, H, W)).astype(np.float32), None)
    hm[3] = 0.0
    hm[4] = -1.0                                                   # maximum <= 0 -> (-1, -1)
    hm[7, 0, 0] = 5.0                                              # maximum on the border: no sub-pixel shift
    kp, mom = dec(_cuda(hm), generic=(generic is True), force_tma=(generic == 'tma'), no_stage=(generic == 'no_stage'))
    kp, mom = kp.cpu().numpy(), mom.cpu().numpy()
    ref_kp, ref_sc = D.argmax_decode(hm)
    assert np.array_equal(kp[:, :2], ref_kp)
    assert np.array_equal(kp[:, 2], ref_sc)
    ref_m = D.heatmap_means_cov_f64(hm)
    scale = max(H, W) ** 2 / 4096.0
    assert np.abs(mom - ref_m).max() < MOMENT_ATOL * max(1.0, scale)
    ref32 = D.heatmap_means_cov(hm.copy())
    assert np.abs(mom - ref32).max() < 4 * MOMENT_ATOL * max(1.0, scale)


def test_host_pipeline_and_outputs_optional(dec, syn):
    hm, _ = syn.gaussian_blob_heatmaps(6000, seed=5)               # 73 MB: two chunks
    kp_d, mom_d = dec(_cuda(hm))
    kp_h, mom_h = dec(hm)
    assert np.array_equal(kp_d.cpu().numpy(), kp_h) and np.array_equal(mom_d.cpu().numpy(), mom_h)
    only_k, none_m = dec(hm, want_moments=False)
    assert none_m is None and np.array_equal(only_k, kp_h)
    none_k, only_m = dec(_cuda(hm), want_kpts=False)
    assert none_k is None and np.array_equal(only_m.cpu().numpy(), mom_h)
    e_k, e_m = dec(np.empty((0, 64, 48), dtype=np.float32))
    assert e_k.shape == (0, 3) and e_m.shape == (0, 6)


def test_transposed_layouts_feed_triangulation(dec, syn):
    """Config 3 shape: heatmaps (T, C, J, H, W) -> keypoints (T, J, C, 3) / (T, J, 3, C) with an affine per (t, c)."""
    T_, C, J = 5, 4, 17
    hm, _ = syn.gaussian_blob_heatmaps(T_ * C * J, seed=6)
    hm = hm.reshape(T_, C, J, 64, 48)
    aff = np.random.default_rng(1).uniform(0.5, 20, size=(T_ * C, 4)).astype(np.float32)
    plain, _ = dec(_cuda(hm), want_moments=False)
    nv3, _ = dec(_cuda(hm), want_moments=False, kpt_layout='nv3', affine=aff, affine_group=J)
    n3v, _ = dec(_cuda(hm), want_moments=False, kpt_layout='n3v', affine=aff, affine_group=J)
    plain = plain.cpu().numpy()                                    # (T, C, J, 3)
    a = aff.reshape(T_, C, 1, 4)
    expect = plain.copy()
    expect[..., 0] = np.float32(plain[..., 0]) * a[..., 0] + a[..., 2]
    expect[..., 1] = np.float32(plain[..., 1]) * a[..., 1] + a[..., 3]
    expect = np.transpose(expect, (0, 2, 1, 3))                    # (T, J, C, 3)
    assert tuple(nv3.shape) == (T_, J, C, 3) and tuple(n3v.shape) == (T_, J, 3, C)
    assert np.allclose(nv3.cpu().numpy(), expect, rtol=1e-6, atol=1e-5)
    assert np.array_equal(np.transpose(n3v.cpu().numpy(), (0, 1, 3, 2)), nv3.cpu().numpy())


def test_write_back_on_device_keeps_subpixel_decode(dec, syn):
    """In-place thresholding (upstream quirk Q7) must not disturb the raw-neighbour read of the quarter-pixel shift."""
    hm, _ = syn.gaussian_blob_heatmaps(500, seed=8, noise=0.02)
    for kw in ({}, {'force_tma': True}, {'generic': True}, {'no_stage': True}):
        t = _cuda(hm.copy())
        kp, mom = dec(t, write_back=True, **kw)
        ref_kp, ref_sc = D.argmax_decode(hm)
        assert np.array_equal(kp.cpu().numpy()[:, :2], ref_kp) and np.array_equal(kp.cpu().numpy()[:, 2], ref_sc)
        expect = hm.copy()
        expect[expect < 0.01] = 0
        assert np.array_equal(t.cpu().numpy(), expect)
        assert np.abs(mom.cpu().numpy() - D.heatmap_means_cov_f64(hm)).max() < MOMENT_ATOL


def test_nan_map_gives_nan_moments_only_there(dec, syn):
    hm, _ = syn.gaussian_blob_heatmaps(40, seed=7)
    ref = dec(_cuda(hm))[1].cpu().numpy()
    hm2 = hm.copy()
    hm2[9, 3, 3] = np.nan
    got = dec(_cuda(hm2))[1].cpu().numpy()
    assert np.isnan(got[9]).all()
    keep = np.arange(40) != 9
    assert np.array_equal(got[keep], ref[keep])


def test_argmax_half_matches_transformers_port_of_mmpose(dec):
    """The kernel against tests/golden/argmax_vitpose.npz (HF transformers' port of mmpose's ``_get_max_preds``):
    integer pixel and score of every map, through all three kernel variants."""
    g = load_golden('argmax_vitpose.npz')
    for kw in ({}, {'force_tma': True}, {'generic': True}, {'no_stage': True}):
        kp, _ = dec(_cuda(g['heatmaps'].copy()), **kw)
        kp = kp.cpu().numpy()
        assert np.array_equal(kp[:, 2], g['scores'])
        assert np.array_equal(np.rint(kp[:, :2]), g['coords'])
        assert np.all(np.isin(np.abs(kp[:, :2] - g['coords']), (0.0, 0.25)))
