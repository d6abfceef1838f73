"""Batched multi-view DLT triangulation on B200 (host side of csrc/triangulate.cu).

``triangulate_multiview`` is the V-view, confidence-weighted generalisation of the reference's
``utils.DLT`` (utils.py:19-34); ``mode='top2'`` is the per-joint camera selection + 2-view solve of
``pose_estimation.get_pose_3D`` (pose_estimation.py:27-54).  torch tensors on a CUDA device are
processed in place on the current stream; numpy arrays go through the library's chunked
H2D -> kernel -> D2H pipeline.  There is no CPU implementation.
"""
import ctypes

import numpy as np

from . import _lib

_LAYOUTS = {'nv3': _lib.LAYOUT_V3, 'n3v': _lib.LAYOUT_3V}
_MODES = {'weighted': _lib.TRI_WEIGHTED, 'top2': _lib.TRI_TOP2}


def _views_of(shape, layout):
    if len(shape) < 2:
        raise ValueError(f'keypoints must have shape (..., V, 3) or (..., 3, V), got {tuple(shape)}')
    if layout == 'nv3':
        if shape[-1] != 3:
            raise ValueError(f"layout 'nv3' expects (..., V, 3), got {tuple(shape)}")
        return shape[-2]
    if shape[-2] != 3:
        raise ValueError(f"layout 'n3v' expects (..., 3, V), got {tuple(shape)}")
    return shape[-1]


def triangulate_multiview(kpts, P, K=None, dist=None, layout='nv3', mode='weighted', out=None, flags=0, device=None):
    """Triangulate every joint of ``kpts`` from V views.

    kpts   (..., V, 3) [x, y, w] (layout 'nv3') or (..., 3, V) (layout 'n3v', the reference's
           kpts_2d (T, J, 3, C) of pose_estimation.py:135); float32 or float64;
           a CUDA torch tensor (stays on the device) or a numpy array (host pipeline).
    P      (V, 3, 4) projection matrices K[R|T] (host, float64).
    K,dist optional (V,3,3), (V,5): undistort the pixels first like utils.py:1314-1315.
    mode   'weighted' (all views, rows scaled by w) or 'top2' (two highest w, unweighted).
    Returns (..., 3) in the dtype / container of the input.
    """
    if layout not in _LAYOUTS:
        raise ValueError(f'unknown layout {layout!r}')
    if mode not in _MODES:
        raise ValueError(f'unknown mode {mode!r}')
    lib = _lib.lib()
    n_views = _views_of(kpts.shape, layout)
    rig, keep = _lib.make_rig(P, K, dist)
    if rig.n_views != n_views:
        raise ValueError(f'P describes {rig.n_views} views but keypoints have {n_views}')
    lead = tuple(kpts.shape[:-2])
    n = int(np.prod(lead)) if lead else 1

    if isinstance(kpts, np.ndarray):
        if kpts.dtype not in (np.float32, np.float64):
            kpts = kpts.astype(np.float64)
        kpts = np.ascontiguousarray(kpts)
        res = np.empty(lead + (3,), dtype=kpts.dtype) if out is None else out
        fn = lib.mc3d_triangulate_host_f32 if kpts.dtype == np.float32 else lib.mc3d_triangulate_host_f64
        _lib.check(fn(kpts.ctypes.data, n, ctypes.byref(rig), _LAYOUTS[layout], _MODES[mode], flags,
                      res.ctypes.data, _lib.default_device(device)))
        return res

    import torch
    if not isinstance(kpts, torch.Tensor):
        raise TypeError('kpts must be a numpy array or a torch tensor')
    if not kpts.is_cuda:
        raise _lib.Mc3dError('torch keypoints must live on a CUDA device (no CPU fallback); '
                             'pass a numpy array to use the host pipeline')
    if kpts.dtype not in (torch.float32, torch.float64):
        kpts = kpts.to(torch.float64)
    kpts = kpts.contiguous()
    res = torch.empty(lead + (3,), dtype=kpts.dtype, device=kpts.device) if out is None else out
    fn = lib.mc3d_triangulate_f32 if kpts.dtype == torch.float32 else lib.mc3d_triangulate_f64
    with torch.cuda.device(kpts.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(fn(kpts.data_ptr(), n, ctypes.byref(rig), _LAYOUTS[layout], _MODES[mode], flags,
                      res.data_ptr(), stream))
    return res
