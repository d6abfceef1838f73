"""A/B check of tuning builds of the triangulation kernels (profiles/build_variants.py) on one B200.

    python profiles/build_variants.py && python profiles/tri_variant_check.py [--dtype f32|f64] [--joints 17000000] [--views 8]

For every profiles/variants/libmc3d_*.so: the outputs on the bench workload must be BIT-IDENTICAL to the shipped
library's (the variants only re-pack instructions), then both are timed with CUDA events (inputs >> L2).  One line per
library; nothing here is a bench value of record -- a variant that wins becomes the default and is then measured by
bench.py.
"""
import argparse
import ctypes
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--joints', type=int, default=17_000_000)
    ap.add_argument('--views', type=int, default=8)
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--dtype', choices=['f32', 'f64'], default='f32', help='storage type (f64: the lean64 build matters)')
    ap.add_argument('--unusable', type=float, default=0.0, help='fraction of views given weight 0 and a wild pixel')
    args = ap.parse_args()
    import torch
    import bench
    from mc3d_b200 import _lib
    dev = torch.device('cuda:0')
    tdt = torch.float32 if args.dtype == 'f32' else torch.float64
    idt = torch.int32 if args.dtype == 'f32' else torch.int64
    kp, P = bench.make_triangulation_workload(args.joints, args.views, tdt, dev, seed=0)
    if args.unusable > 0:
        bad = torch.rand(kp.shape[:2], device=dev) < args.unusable
        kp[..., 2][bad] = 0.0
        kp[..., 0][bad] = 1.0e4
    rig, keep = _lib.make_rig(P)
    libs = [('shipped', _lib.LIB_PATH)] + [(os.path.basename(p)[8:-3], p)
                                           for p in sorted(glob.glob(os.path.join(ROOT, 'profiles', 'variants', 'libmc3d_*.so')))]
    ref = None
    stream = torch.cuda.current_stream().cuda_stream
    for name, path in libs:
        h = ctypes.CDLL(path)
        fn = getattr(h, f'mc3d_triangulate_{args.dtype}')
        fn.restype = ctypes.c_int
        fn.argtypes = _lib.SIGNATURES[f'mc3d_triangulate_{args.dtype}'][1]
        out = torch.empty((args.joints, 3), dtype=tdt, device=dev)

        def run():
            st = fn(kp.data_ptr(), args.joints, ctypes.byref(rig), _lib.LAYOUT_V3, _lib.TRI_WEIGHTED, 0, out.data_ptr(), stream)
            assert st == 0, st
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        if ref is None:
            ref = out.clone()
            same = 'reference'
        else:
            if torch.equal(out.view(idt), ref.view(idt)):
                same = 'bit-identical'
            else:
                n_diff = (out.view(idt) != ref.view(idt)).any(dim=1).sum().item()
                ok = torch.isfinite(out).all(dim=1) & torch.isfinite(ref).all(dim=1)
                worst = (out[ok].double() - ref[ok].double()).abs().max().item() if ok.any() else float('nan')
                nan_mismatch = (torch.isnan(out) != torch.isnan(ref)).any().item()
                same = f'DIFFERENT in {n_diff} joints (max |diff| {worst:.3e} mm, NaN pattern {"differs" if nan_mismatch else "equal"})'
        print(f'{name:14s} {ms:8.4f} ms  {args.joints / ms * 1e3:.4e} joints/s  {same}', flush=True)


if __name__ == '__main__':
    main()
