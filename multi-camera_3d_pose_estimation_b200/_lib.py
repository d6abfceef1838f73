"""ctypes binding of libmc3d.so (include/mc3d.h).  There is no CPU fallback: if the library is missing
or a call fails, an exception is raised."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libmc3d.so')

MAX_VIEWS = 16
LAYOUT_V3 = 0      # (N, V, 3)
LAYOUT_3V = 1      # (N, 3, V)  -- the reference's kpts_2d (T, J, 3, C)
TRI_WEIGHTED = 0
TRI_TOP2 = 1
TRI_FLAG_JACOBI = 1
KPT_PLAIN, KPT_NV3, KPT_N3V = 0, 1, 2
DECODE_FLAG_WRITE_BACK = 1
DECODE_FLAG_GENERIC = 2


class Mc3dError(RuntimeError):
    pass


class Rig(ctypes.Structure):
    _fields_ = [('n_views', ctypes.c_int32),
                ('P', ctypes.POINTER(ctypes.c_double)),
                ('K', ctypes.POINTER(ctypes.c_double)),
                ('dist', ctypes.POINTER(ctypes.c_double))]


class RefineCamera(ctypes.Structure):
    _fields_ = [('K', ctypes.c_double * 9), ('R', ctypes.c_double * 9), ('T', ctypes.c_double * 3),
                ('dist', ctypes.c_double * 5)]


_lib = None

_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int
_c_vp = ctypes.c_void_p
_c_dbl = ctypes.c_double

# name -> (restype, argtypes); must list every symbol include/mc3d.h declares.
SIGNATURES = {
    'mc3d_version': (_c_int, []),
    'mc3d_last_error': (ctypes.c_char_p, []),
    'mc3d_status_string': (ctypes.c_char_p, [_c_int]),
    'mc3d_launch_count': (_c_i64, []),
    'mc3d_device_info': (_c_int, [ctypes.c_char_p, _c_int, ctypes.POINTER(_c_int), ctypes.POINTER(_c_int),
                                  ctypes.POINTER(_c_int)]),
    'mc3d_triangulate_f32': (_c_int, [_c_vp, _c_i64, ctypes.POINTER(Rig), _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    'mc3d_triangulate_f64': (_c_int, [_c_vp, _c_i64, ctypes.POINTER(Rig), _c_int, _c_int, _c_int, _c_vp, _c_vp]),
    'mc3d_triangulate_host_f32': (_c_int, [_c_vp, _c_i64, ctypes.POINTER(Rig), _c_int, _c_int, _c_int, _c_vp, _c_int]),
    'mc3d_triangulate_host_f64': (_c_int, [_c_vp, _c_i64, ctypes.POINTER(Rig), _c_int, _c_int, _c_int, _c_vp, _c_int]),
    'mc3d_decode_heatmaps_f32': (_c_int, [_c_vp, _c_i64, _c_int, _c_int, ctypes.c_float, _c_int, _c_int, _c_int, _c_int,
                                          _c_vp, _c_int, _c_vp, _c_vp, _c_vp]),
    'mc3d_decode_heatmaps_host_f32': (_c_int, [_c_vp, _c_i64, _c_int, _c_int, ctypes.c_float, _c_vp, _c_vp, _c_int]),
}


def lib():
    """The loaded library; raises Mc3dError when libmc3d.so has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Mc3dError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                            '(nvcc, sm_100a).  This package has no CPU fallback.')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status):
    if status != 0:
        l = lib()
        raise Mc3dError(f'{l.mc3d_status_string(status).decode()}: {l.mc3d_last_error().decode()}')


def launch_count():
    return int(lib().mc3d_launch_count())


def make_rig(P, K=None, dist=None):
    """Rig struct + the numpy arrays that keep its pointers alive."""
    P = np.ascontiguousarray(np.asarray(P, dtype=np.float64).reshape(-1, 12))
    keep = [P]
    rig = Rig()
    rig.n_views = P.shape[0]
    rig.P = P.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    if K is not None:
        K = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(-1, 9))
        dist = np.zeros((P.shape[0], 5)) if dist is None else np.asarray(dist, dtype=np.float64).reshape(P.shape[0], -1)
        d5 = np.zeros((P.shape[0], 5))
        d5[:, :min(5, dist.shape[1])] = dist[:, :5]
        if K.shape[0] != P.shape[0]:
            raise ValueError('K and P must describe the same number of views')
        keep += [K, d5]
        rig.K = K.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        rig.dist = d5.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    return rig, keep
