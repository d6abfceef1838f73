"""linear_interpolation: oracle vs the reference golden (CPU) and the CUDA kernel vs both (GPU)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import interpolation as I

CASES = [('default', {}), ('k9', dict(k=9, k_std=1.5, median_std=3)), ('rolling', dict(k=7, use_rolling_average=True)),
         ('nomedian', dict(k=4, filter_distance_from_median=False))]


@pytest.mark.parametrize('name,kw', CASES)
def test_oracle_matches_reference(name, kw):
    g = load_golden('interp.npz')
    assert np.abs(I.linear_interpolation(g['points'], **kw) - g[name]).max() < 1e-9
    if name == 'default':
        assert np.abs(I.linear_interpolation(g['points'][:, :, 0]) - g['two_dim']).max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize('name,kw', CASES)
def test_kernel_matches_reference(name, kw):
    import __graft_entry__ as ge
    ge.build()
    from mc3d_b200.interpolation import linear_interpolation
    g = load_golden('interp.npz')
    got = linear_interpolation(g['points'], **kw)
    assert got.shape == g[name].shape and got.dtype == np.float64
    assert np.abs(got - g[name]).max() < 1e-9                      # ~3e-13 relative on ~3000 mm coordinates
    if name == 'default':
        assert np.abs(linear_interpolation(g['points'][:, :, 0]) - g['two_dim']).max() < 1e-9


@pytest.mark.gpu
def test_kernel_edge_cases_vs_oracle():
    import torch
    from mc3d_b200.interpolation import linear_interpolation
    rng = np.random.default_rng(2)
    for T in (1, 2, 3, 7):
        X = rng.normal(0, 10, size=(T, 3, 3))
        assert np.abs(linear_interpolation(X, k=5) - I.linear_interpolation(X, k=5)).max() < 1e-10
    X = rng.normal(0, 10, size=(40, 5, 3))
    X[10] = 1e6                                                    # an outlier frame
    ref = I.linear_interpolation(X, k=3, k_std=0.5)                # harsh filter: some windows keep < 2 samples -> 0
    got = linear_interpolation(X, k=3, k_std=0.5)
    assert (ref == 0).any() and np.abs(got - ref).max() < 1e-9
    gpu = linear_interpolation(torch.tensor(X, device='cuda:0'), k=3, k_std=0.5)
    assert gpu.is_cuda and np.abs(gpu.cpu().numpy() - ref).max() < 1e-9
    assert linear_interpolation(np.zeros((0, 4, 3))).shape == (0, 4, 3)
