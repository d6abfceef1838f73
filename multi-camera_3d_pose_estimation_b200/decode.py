"""Heatmap decode on B200 (host side of csrc/decode.cu): argmax keypoints and Gaussian moments in one
pass over the heatmaps.  CUDA torch tensors are processed on the current stream; numpy arrays go
through the library's chunked host pipeline.  No CPU implementation."""
import numpy as np

from . import _lib

_KPT_LAYOUTS = {'plain': _lib.KPT_PLAIN, 'nv3': _lib.KPT_NV3, 'n3v': _lib.KPT_N3V}


def decode_heatmaps(heatmaps, threshold=0.01, want_kpts=True, want_moments=True, write_back=False,
                    kpt_layout='plain', affine=None, affine_group=None, generic=False, device=None, force_tma=False,
                    no_stage=False):
    """heatmaps (..., H, W) float32 -> (kpts (..., 3) float32 [x, y, score] | None,
                                         moments (..., 6) float64 | None).

    kpt_layout 'nv3' / 'n3v' expect heatmaps (T, C, J, H, W) (the reference's order, pose_estimation.py:110,190)
    and return keypoints as (T, J, C, 3) / (T, J, 3, C), ready for ``triangulate_multiview``.
    write_back=True also zeroes values < threshold in the input, as upstream does (mmpose_pose_estimation.py:166).
    affine (G, 4) float32 [sx, sy, ox, oy] per ``affine_group`` consecutive maps maps heatmap pixels to image pixels.
    generic / force_tma / no_stage select a kernel variant (test hooks: plain loads; the shared-memory kernel for 64x48
    maps; the 64x48 moments kernel without its per-warp TMA stage).
    """
    lib = _lib.lib()
    if heatmaps.ndim < 2:
        raise ValueError('heatmaps must have shape (..., H, W)')
    H, W = int(heatmaps.shape[-2]), int(heatmaps.shape[-1])
    lead = tuple(heatmaps.shape[:-2])
    n = int(np.prod(lead)) if lead else 1
    if kpt_layout not in _KPT_LAYOUTS:
        raise ValueError(f'unknown kpt_layout {kpt_layout!r}')
    views = joints = 0
    kshape = lead + (3,)
    if kpt_layout != 'plain':
        if len(lead) != 3:
            raise ValueError("kpt_layout 'nv3'/'n3v' needs heatmaps of shape (T, C, J, H, W)")
        T_, views, joints = lead
        kshape = (T_, joints, views, 3) if kpt_layout == 'nv3' else (T_, joints, 3, views)

    if isinstance(heatmaps, np.ndarray):
        if kpt_layout != 'plain' or affine is not None or write_back or generic:
            raise ValueError('the host pipeline supports the plain layout only; pass a CUDA tensor')
        hm = np.ascontiguousarray(heatmaps, dtype=np.float32)
        kpts = np.empty(kshape, dtype=np.float32) if want_kpts else None
        mom = np.empty(lead + (6,), dtype=np.float64) if want_moments else None
        _lib.check(lib.mc3d_decode_heatmaps_host_f32(hm.ctypes.data, n, H, W, float(threshold),
                                                     kpts.ctypes.data if want_kpts else None,
                                                     mom.ctypes.data if want_moments else None, _lib.default_device(device)))
        return kpts, mom

    import torch
    if not isinstance(heatmaps, torch.Tensor):
        raise TypeError('heatmaps must be a numpy array or a torch tensor')
    if not heatmaps.is_cuda:
        raise _lib.Mc3dError('torch heatmaps must live on a CUDA device (no CPU fallback)')
    if heatmaps.dtype != torch.float32:
        raise TypeError('heatmaps must be float32')
    if not heatmaps.is_contiguous():
        if write_back:
            raise ValueError('write_back needs a contiguous tensor')
        heatmaps = heatmaps.contiguous()
    dev = heatmaps.device
    kpts = torch.empty(kshape, dtype=torch.float32, device=dev) if want_kpts else None
    mom = torch.empty(lead + (6,), dtype=torch.float64, device=dev) if want_moments else None
    aff_ptr, group = None, 0
    if affine is not None:
        affine = torch.as_tensor(affine, dtype=torch.float32, device=dev).contiguous().reshape(-1, 4)
        group = int(affine_group) if affine_group else max(1, n // affine.shape[0])
        if affine.shape[0] * group < n:
            raise ValueError('affine table too short for the number of heatmaps')
        aff_ptr = affine.data_ptr()
    flags = (_lib.DECODE_FLAG_WRITE_BACK if write_back else 0) | (_lib.DECODE_FLAG_GENERIC if generic else 0) | \
        (_lib.DECODE_FLAG_TMA if force_tma else 0) | (_lib.DECODE_FLAG_NO_STAGE if no_stage else 0)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.mc3d_decode_heatmaps_f32(heatmaps.data_ptr(), n, H, W, float(threshold), flags,
                                                _KPT_LAYOUTS[kpt_layout], views, joints, aff_ptr, group,
                                                kpts.data_ptr() if want_kpts else None,
                                                mom.data_ptr() if want_moments else None, stream))
    return kpts, mom
