"""Drop-in for ``pose_refinement.linear_interpolation`` (pose_refinement.py:15-84) on the GPU (csrc/interp.cu)."""
import numpy as np

from . import _lib


def linear_interpolation(points, k=5, k_std=2, median_std=2, use_rolling_average=False, filter_distance_from_median=True,
                         device=None):
    """Smooth ``points`` ([time, n_points, dim] or [time, n_points]) by windowed outlier rejection and a local line
    fit; same arguments and result shape as upstream.  numpy in -> numpy out; a CUDA tensor in -> CUDA tensor out."""
    import torch
    lib = _lib.lib()
    is_tensor = isinstance(points, torch.Tensor)
    if is_tensor and points.is_cuda:
        p = points.to(torch.float64).contiguous()
        dev = points.device
    else:
        if not torch.cuda.is_available():
            raise _lib.Mc3dError('linear_interpolation needs a CUDA device (no CPU fallback)')
        arr = points.numpy() if is_tensor else np.array(points)
        dev = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
        p = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64)).to(dev)
    if p.dim() not in (2, 3):
        raise ValueError('points must have shape [time, n_points, dim] or [time, n_points]')
    T = int(p.shape[0])
    pd = int(np.prod(p.shape[1:])) if p.dim() > 1 else 1
    out = torch.empty_like(p)
    with torch.cuda.device(dev):
        _lib.check(lib.mc3d_linear_interpolation_f64(p.data_ptr(), T, pd, int(k), float(k_std), float(median_std),
                                                     int(bool(use_rolling_average)), int(bool(filter_distance_from_median)),
                                                     out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    if is_tensor and points.is_cuda:
        return out.to(points.dtype)
    res = out.cpu().numpy()
    src_dtype = (points.numpy() if is_tensor else np.asarray(points)).dtype
    return res.astype(src_dtype) if np.issubdtype(src_dtype, np.floating) else res
