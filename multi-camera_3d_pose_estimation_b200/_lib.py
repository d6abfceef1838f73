"""ctypes binding of libmc3d.so (include/mc3d.h).  There is no CPU fallback: if the library is missing
or a call fails, an exception is raised."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MC3D_LIB') or os.path.join(HERE, 'libmc3d.so')     # MC3D_LIB: tuning builds only

MAX_VIEWS = 16
LAYOUT_V3 = 0      # (N, V, 3)
LAYOUT_3V = 1      # (N, 3, V)  -- the reference's kpts_2d (T, J, 3, C)
TRI_WEIGHTED = 0
TRI_TOP2 = 1
TRI_FLAG_JACOBI = 1
TRI_FLAG_FP64 = 2
KPT_PLAIN, KPT_NV3, KPT_N3V = 0, 1, 2
DECODE_FLAG_WRITE_BACK = 1
DECODE_FLAG_GENERIC = 2
DECODE_FLAG_TMA = 4
DECODE_FLAG_NO_STAGE = 8


class Mc3dError(RuntimeError):
    pass


class Rig(ctypes.Structure):
    _fields_ = [('n_views', ctypes.c_int32),
                ('P', ctypes.POINTER(ctypes.c_double)),
                ('K', ctypes.POINTER(ctypes.c_double)),
                ('dist', ctypes.POINTER(ctypes.c_double))]


class TriStartPair(ctypes.Structure):
    """mc3d_tri_start_pair (include/mc3d.h): constants of one closed-form two-view starting point."""
    _fields_ = [('H', ctypes.c_float * 9), ('C', ctypes.c_float * 3), ('ua', ctypes.c_float * 3), ('ub', ctypes.c_float * 3),
                ('alpha', ctypes.c_float), ('beta', ctypes.c_float), ('ka', ctypes.c_float), ('kb', ctypes.c_float),
                ('view_a', ctypes.c_int32), ('view_b', ctypes.c_int32)]


MAX_JOINTS = 133
MAX_BONES = 64
MAX_PEERS = 16
IPC_HANDLE_BYTES = 64
XCHG_X_OFFSET = 32768
CT_ACC, CT_STATE, CT_HIST = 0, 32, 64          # control-block layout (include/mc3d.h)


class RefineProblem(ctypes.Structure):
    """mc3d_refine_problem (include/mc3d.h)."""
    _fields_ = [('n_joints', ctypes.c_int32), ('n_cams', ctypes.c_int32), ('n_bones', ctypes.c_int32),
                ('ignore_distortions', ctypes.c_int32), ('patience', ctypes.c_int32), ('max_iter', ctypes.c_int32),
                ('n_frames', ctypes.c_int64), ('frame_offset', ctypes.c_int64), ('win_begin', ctypes.c_int64),
                ('win_end', ctypes.c_int64), ('hist_capacity', ctypes.c_int64), ('total_frames', ctypes.c_int64),
                ('lr', ctypes.c_double), ('beta1', ctypes.c_double), ('beta2', ctypes.c_double),
                ('eps', ctypes.c_double), ('lambda_smooth', ctypes.c_double), ('lambda_body', ctypes.c_double),
                ('tolerance', ctypes.c_double), ('aa', ctypes.c_double),
                ('cams', (ctypes.c_double * 26) * MAX_VIEWS),
                ('bone_len', ctypes.c_double * MAX_BONES),
                ('bone_start', ctypes.c_int32 * MAX_BONES), ('bone_end', ctypes.c_int32 * MAX_BONES),
                ('adj_start', ctypes.c_int32 * (MAX_JOINTS + 3)),
                ('adj_bone', ctypes.c_int32 * (2 * MAX_BONES)), ('adj_sign', ctypes.c_int32 * (2 * MAX_BONES)),
                ('x', ctypes.c_void_p), ('m', ctypes.c_void_p), ('v', ctypes.c_void_p), ('best', ctypes.c_void_p),
                ('g', ctypes.c_void_p), ('mu0', ctypes.c_void_p), ('S', ctypes.c_void_p),
                ('term_ok', ctypes.c_void_p), ('ctrl', ctypes.c_void_p), ('gc', ctypes.c_void_p),
                ('rank', ctypes.c_int32), ('world', ctypes.c_int32), ('n_frames_left', ctypes.c_int64),
                ('spin_timeout_ns', ctypes.c_int64), ('xchg', ctypes.c_void_p * MAX_PEERS),
                ('gauss_cam_stride', ctypes.c_int64), ('cams_dev', ctypes.c_void_p), ('test_flags', ctypes.c_int64)]


class RefineXchg(ctypes.Structure):
    """mc3d_refine_xchg (include/mc3d.h): head of a rank's peer allocation."""
    _fields_ = [('sums', ((ctypes.c_double * 8) * MAX_PEERS) * 2),
                ('seq_costs', (ctypes.c_int64 * MAX_PEERS) * 2), ('seq_grad', (ctypes.c_int64 * MAX_PEERS) * 2),
                ('halo_seq', ctypes.c_int64 * 2), ('ticket', ctypes.c_int64 * 4), ('gen', ctypes.c_int64 * 4),
                ('error', ctypes.c_int64),
                ('acc2', (ctypes.c_double * 24) * 2), ('sums2', ((ctypes.c_double * 24) * MAX_PEERS) * 2),
                ('seq2', (ctypes.c_int64 * MAX_PEERS) * 2), ('ll', ((ctypes.c_int64 * 40) * MAX_PEERS) * 2),
                ('ll_retry', (ctypes.c_int64 * 40) * MAX_PEERS), ('blk_seq', ctypes.c_int64 * 960)]


class ExtrinsicProblem(ctypes.Structure):
    """mc3d_extrinsic_problem (include/mc3d.h)."""
    _fields_ = [('n_frames', ctypes.c_int64), ('n_joints', ctypes.c_int32), ('n_samples', ctypes.c_int32),
                ('ignore_distortions', ctypes.c_int32), ('patience', ctypes.c_int32), ('max_iter', ctypes.c_int32),
                ('reserved', ctypes.c_int32), ('hist_capacity', ctypes.c_int64),
                ('lr', ctypes.c_double), ('beta1', ctypes.c_double), ('beta2', ctypes.c_double), ('eps', ctypes.c_double),
                ('tolerance', ctypes.c_double), ('const_cost', ctypes.c_double),
                ('K', ctypes.c_double * 9), ('dist', ctypes.c_double * 5),
                ('samples3d', ctypes.c_void_p), ('mean', ctypes.c_void_p), ('S', ctypes.c_void_p),
                ('params', ctypes.c_void_p), ('ctrl', ctypes.c_void_p)]


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
    'mc3d_triangulate_start_plan': (_c_int, [ctypes.POINTER(Rig), ctypes.POINTER(TriStartPair), ctypes.POINTER(ctypes.c_int32)]),
    'mc3d_triangulate_host_f32': (_c_int, [_c_vp, _c_i64, ctypes.POINTER(Rig), _c_int, _c_int, _c_int, _c_vp, _c_int]),
    'mc3d_triangulate_host_f64': (_c_int, [_c_vp, _c_i64, ctypes.POINTER(Rig), _c_int, _c_int, _c_int, _c_vp, _c_int]),
    'mc3d_decode_heatmaps_f32': (_c_int, [_c_vp, _c_i64, _c_int, _c_int, ctypes.c_float, _c_int, _c_int, _c_int, _c_int,
                                          _c_vp, _c_int, _c_vp, _c_vp, _c_vp]),
    'mc3d_decode_heatmaps_host_f32': (_c_int, [_c_vp, _c_i64, _c_int, _c_int, ctypes.c_float, _c_vp, _c_vp, _c_int]),
    'mc3d_project_points_f32': (_c_int, [_c_vp, _c_i64, _c_vp, _c_int, _c_vp, _c_vp]),
    'mc3d_project_points_f64': (_c_int, [_c_vp, _c_i64, _c_vp, _c_int, _c_vp, _c_vp]),
    'mc3d_refine_prepare_f32': (_c_int, [_c_vp, _c_i64, _c_int, _c_int, _c_int, _c_dbl, _c_vp, _c_vp, _c_vp]),
    'mc3d_refine_prepare_f64': (_c_int, [_c_vp, _c_i64, _c_int, _c_int, _c_int, _c_dbl, _c_vp, _c_vp, _c_vp]),
    'mc3d_refine_problem_size': (_c_int, []),
    'mc3d_refine_plan': (ctypes.c_char_p, [ctypes.POINTER(RefineProblem)]),
    'mc3d_refine_sweep_range': (_c_int, [_c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, ctypes.POINTER(_c_i64), ctypes.POINTER(_c_i64)]),
    'mc3d_refine_flags_f32': (_c_int, [ctypes.POINTER(RefineProblem), _c_vp]),
    'mc3d_refine_flags_f64': (_c_int, [ctypes.POINTER(RefineProblem), _c_vp]),
    'mc3d_refine_phase_f32': (_c_int, [ctypes.POINTER(RefineProblem), _c_int, _c_i64, _c_int, _c_vp]),
    'mc3d_refine_phase_f64': (_c_int, [ctypes.POINTER(RefineProblem), _c_int, _c_i64, _c_int, _c_vp]),
    'mc3d_refine_run_f32': (_c_int, [ctypes.POINTER(RefineProblem), _c_i64, _c_i64, _c_vp]),
    'mc3d_refine_run_f64': (_c_int, [ctypes.POINTER(RefineProblem), _c_i64, _c_i64, _c_vp]),
    'mc3d_extrinsic_problem_size': (_c_int, []),
    'mc3d_extrinsic_run_f32': (_c_int, [ctypes.POINTER(ExtrinsicProblem), _c_i64, _c_i64, _c_vp]),
    'mc3d_extrinsic_run_f64': (_c_int, [ctypes.POINTER(ExtrinsicProblem), _c_i64, _c_i64, _c_vp]),
    'mc3d_extrinsic_costgrad_f32': (_c_int, [ctypes.POINTER(ExtrinsicProblem), _c_vp]),
    'mc3d_extrinsic_costgrad_f64': (_c_int, [ctypes.POINTER(ExtrinsicProblem), _c_vp]),
    'mc3d_extrinsic_joint_step_f32': (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_i64, _c_dbl, _c_dbl, _c_dbl, _c_dbl, _c_vp]),
    'mc3d_extrinsic_joint_step_f64': (_c_int, [_c_vp, _c_vp, _c_vp, _c_int, _c_i64, _c_dbl, _c_dbl, _c_dbl, _c_dbl, _c_vp]),
    'mc3d_peer_alloc': (_c_int, [_c_i64, ctypes.POINTER(_c_vp), _c_vp]),
    'mc3d_peer_open': (_c_int, [_c_vp, ctypes.POINTER(_c_vp)]),
    'mc3d_peer_close': (_c_int, [_c_vp]),
    'mc3d_peer_free': (_c_int, [_c_vp]),
    'mc3d_linear_interpolation_f64': (_c_int, [_c_vp, _c_i64, _c_i64, _c_int, _c_dbl, _c_dbl, _c_int, _c_int, _c_vp, _c_vp]),
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
        if handle.mc3d_refine_problem_size() != ctypes.sizeof(RefineProblem):
            raise Mc3dError('mc3d_refine_problem layout mismatch between include/mc3d.h and _lib.py')
        if handle.mc3d_extrinsic_problem_size() != ctypes.sizeof(ExtrinsicProblem):
            raise Mc3dError('mc3d_extrinsic_problem layout mismatch between include/mc3d.h and _lib.py')
        _lib = handle
    return _lib


def check(status):
    if status != 0:
        l = lib()
        raise Mc3dError(f'{l.mc3d_status_string(status).decode()}: {l.mc3d_last_error().decode()}')


def default_device(device=None):
    """CUDA ordinal for the host-buffer pipelines: an explicit ``device`` wins; otherwise the process's current CUDA
    device when torch has one selected (under torchrun: the rank's own GPU), else LOCAL_RANK, else 0.  The library
    itself restores the calling thread's current device before it returns."""
    if device is not None:
        if hasattr(device, 'index'):                      # torch.device
            return int(device.index or 0)
        return int(device)
    try:
        import torch
        if torch.cuda.is_available() and torch.cuda.is_initialized():
            return int(torch.cuda.current_device())
    except Exception:
        pass
    return int(os.environ.get('LOCAL_RANK', '0') or 0)


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
