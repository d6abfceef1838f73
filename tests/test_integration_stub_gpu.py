"""GPU: the raw-ctypes binding shown in INTEGRATION.md section 2 (no mc3d_b200 package involved) reproduces the
reference's utils.triangulate_points golden."""
import ctypes

import numpy as np
import pytest

from conftest import cams_from_golden, load_golden, rel_err

pytestmark = pytest.mark.gpu


class _Rig(ctypes.Structure):
    _fields_ = [('n_views', ctypes.c_int32), ('P', ctypes.POINTER(ctypes.c_double)),
                ('K', ctypes.POINTER(ctypes.c_double)), ('dist', ctypes.POINTER(ctypes.c_double))]


def test_raw_ctypes_stub_matches_reference():
    import __graft_entry__ as g
    lib = ctypes.CDLL(g.build())
    lib.mc3d_triangulate_host_f64.restype = ctypes.c_int
    lib.mc3d_triangulate_host_f64.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(_Rig), ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    lib.mc3d_last_error.restype = ctypes.c_char_p
    gold = load_golden('dlt_stereo.npz')
    c = cams_from_golden(gold, 2)
    pts = gold['pair'].reshape(-1, 2, 2)
    kp = np.ascontiguousarray(np.concatenate([pts, np.ones((len(pts), 2, 1))], axis=2))
    P = np.ascontiguousarray(np.stack([c[i][0] @ np.hstack((c[i][1], c[i][2].reshape(3, 1))) for i in range(2)]))
    K = np.ascontiguousarray(np.stack([c[0][0], c[1][0]]).astype(np.float64))
    D = np.ascontiguousarray(np.stack([np.ravel(c[0][3])[:5], np.ravel(c[1][3])[:5]]).astype(np.float64))
    as_p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    rig = _Rig(2, as_p(P), as_p(K), as_p(D))
    out = np.empty((len(pts), 3))
    status = lib.mc3d_triangulate_host_f64(kp.ctypes.data, len(pts), ctypes.byref(rig), 0, 0, 0, out.ctypes.data, 0)
    assert status == 0, lib.mc3d_last_error().decode()
    assert rel_err(out, gold['tri'].reshape(-1, 3)).max() < 1e-9
    # error path: a bad view count is reported through the status / last-error channel, not a crash
    bad = _Rig(1, as_p(P), None, None)
    assert lib.mc3d_triangulate_host_f64(kp.ctypes.data, 4, ctypes.byref(bad), 0, 0, 0, out.ctypes.data, 0) == 1
    assert b'n_views' in lib.mc3d_last_error()
