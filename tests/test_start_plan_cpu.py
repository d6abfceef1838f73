"""CPU: the starting-point plan of the float triangulation kernel (host code of libmc3d.so, no device work).

`mc3d_triangulate_start_plan` picks up to four pairs of views and the constants of the closed-form two-view point the
kernel starts from (csrc/triangulate.cu: fill_start_pairs / pair_start).  The kernel's RESULT does not depend on the
plan (its fixed point is set by double residuals over all views); what has to hold is that the start lands within the
kernel's acceptance radius (|e|^2 <= 1e-3 (|X|^2 + rig^2), a few per cent of the range) for ordinary input, otherwise
every joint pays a second pass.  The closed form is evaluated here in float32 numpy exactly as the kernel does and
compared with the oracle's DLT (reference utils.py:19-34 generalised).
"""
import ctypes

import numpy as np
import pytest

from oracle import dlt as O


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    import mc3d_b200
    return mc3d_b200.lib()


def start_plan(lib, P):
    from mc3d_b200 import _lib
    rig, keep = _lib.make_rig(P)
    pairs = (_lib.TriStartPair * 4)()
    n = ctypes.c_int32(-1)
    assert lib.mc3d_triangulate_start_plan(ctypes.byref(rig), pairs, ctypes.byref(n)) == 0
    return [pairs[i] for i in range(n.value)]


def eval_start(pc, kp):
    """pair_start of csrc/triangulate.cu in float32."""
    f = np.float32
    H = np.array(pc.H, dtype=f).reshape(3, 3)
    C = np.array(pc.C, dtype=f)
    ua, ub = np.array(pc.ua, dtype=f), np.array(pc.ub, dtype=f)
    xa, ya = kp[:, pc.view_a, 0], kp[:, pc.view_a, 1]
    xb, yb = kp[:, pc.view_b, 0], kp[:, pc.view_b, 1]
    z = f(pc.alpha) * xb + f(pc.beta) * yb
    ta = ua[0] * xa + ua[1] * ya + ua[2]
    tb = ub[0] * xa + ub[1] * ya + ub[2]
    s = (z * f(pc.kb) - f(pc.ka)) / (ta - z * tb)
    D = np.stack([H[r, 0] * xa + H[r, 1] * ya + H[r, 2] for r in range(3)], axis=1)
    return C[None, :] + s[:, None] * D


@pytest.mark.parametrize('n_views', [3, 4, 8, 16])
def test_ring_rig_pairs_are_disjoint_and_start_near_the_dlt_point(lib, syn, n_views):
    kp, P, X, _ = syn.multiview_points(20000, n_views, seed=3)
    kp = kp.astype(np.float32)
    pairs = start_plan(lib, P)
    assert len(pairs) == min(4, n_views * (n_views - 1) // 2)
    views = [(p.view_a, p.view_b) for p in pairs]
    assert all(0 <= v < n_views for pr in views for v in pr) and all(a != b for a, b in views)
    assert len({frozenset(pr) for pr in views}) == len(views)            # no pair twice
    n_disjoint = min(len(pairs), n_views // 2)                           # a zero-weight view spoils one pair only, while views last
    assert len({v for pr in views[:n_disjoint] for v in pr}) == 2 * n_disjoint, views
    ref = O.dlt_weighted(kp.astype(np.float64), P)
    for k, pc in enumerate(pairs):
        assert abs(pc.alpha ** 2 + pc.beta ** 2 - 1.0) < 1e-5
        d = np.linalg.norm(eval_start(pc, kp) - ref, axis=1)
        # 1 px of noise at 3 m is a few millimetres; the acceptance radius is ~3 % of the range (~130 mm).  The later pairs
        # have narrower angles: farther starts (another pass at worst), still inside the radius for nearly every joint
        assert np.median(d) < (10.0 if k < 2 else 40.0), (k, np.median(d))
        assert np.mean(d < 100.0) > (0.999 if k < 2 else 0.9), (k, np.mean(d < 100.0))


def test_stereo_rig_with_diverging_axes(lib, syn):
    """Config-1 rig: the optical axes meet BEHIND the cameras; the plan must not aim at that point."""
    cams = syn.stereo_rig(distortion=False)
    P = syn.projection_matrices(cams)
    rng = np.random.default_rng(0)
    X = np.array([0.0, 0.0, 3000.0]) + rng.normal(0.0, 400.0, size=(5000, 3))
    kp = np.ones((5000, 2, 3))
    for c, cam in enumerate(cams.values()):
        kp[:, c, :2] = syn.project(X, cam, distort=False) + rng.normal(0.0, 1.0, size=(5000, 2))
    kp = kp.astype(np.float32)
    pairs = start_plan(lib, P)
    assert sorted((p.view_a, p.view_b) for p in pairs) == [(0, 1), (1, 0)]
    ref = O.dlt_weighted(kp.astype(np.float64), P)
    for pc in pairs:
        d = np.linalg.norm(eval_start(pc, kp) - ref, axis=1)
        assert np.median(d) < 10.0 and np.mean(d < 100.0) > 0.999


def test_degenerate_rigs_give_no_plan_instead_of_garbage(lib):
    """Affine / singular cameras have no centre: zero pairs (the kernel then starts every joint from the origin)."""
    P = np.zeros((3, 3, 4))
    P[:, 0, 0] = P[:, 1, 1] = 1.0
    P[:, 2, 3] = 1.0                                  # affine cameras: third row (0, 0, 0, 1)
    assert start_plan(lib, P) == []
    # one finite camera among them is still not a pair
    P[0] = np.array([[1000.0, 0, 640, 0], [0, 1000.0, 360, 0], [0, 0, 1, 0]])
    assert start_plan(lib, P) == []


def test_bad_arguments(lib):
    from mc3d_b200 import _lib
    rig, keep = _lib.make_rig(np.zeros((1, 3, 4)))
    pairs = (_lib.TriStartPair * 4)()
    n = ctypes.c_int32(0)
    assert lib.mc3d_triangulate_start_plan(ctypes.byref(rig), pairs, ctypes.byref(n)) == 1
    assert lib.mc3d_triangulate_start_plan(None, pairs, ctypes.byref(n)) == 1
