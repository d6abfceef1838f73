"""Triangulation throughput on one B200 (CUDA events, inputs >> L2) for every dispatched kernel variant, plus the accuracy of
the float-storage kernel against the all-double solver on the same input.

    python profiles/tri_rates.py [--joints 17000000]

Rows: float storage V = 8 on clean input and with 1 / 5 / 20 % of the views unusable (weight 0, wild pixel: a joint whose
first starting pair holds such a view takes the second pair; a warp only pays another pass when both pairs fail), the
reference's (N, 3, V) layout, V = 16 / 4 / 2, and double storage.  Nothing here is a bench value of record (bench.py is).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--joints', type=int, default=17_000_000)
    ap.add_argument('--reps', type=int, default=10)
    args = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    import torch
    from bench import make_triangulation_workload
    from mc3d_b200 import _lib
    from mc3d_b200.triangulation import triangulate_multiview
    dev = 'cuda:0'
    peak = 6451.8

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.reps

    def report(name, kp, P, layout='nv3', check=False):
        n = kp.shape[0]
        V = kp.shape[1] if layout == 'nv3' else kp.shape[2]
        out = torch.empty((n, 3), dtype=kp.dtype, device=dev)
        ms = timed(lambda: triangulate_multiview(kp, P, layout=layout, out=out))
        gbs = n * (3 * V + 3) * kp.element_size() / ms / 1e6
        extra = ''
        if check:
            ref = triangulate_multiview(kp, P, layout=layout, flags=_lib.TRI_FLAG_FP64)
            ok = torch.isfinite(out).all(dim=1) & torch.isfinite(ref).all(dim=1)
            d = (out[ok].double() - ref[ok].double()).norm(dim=1)
            same_nan = bool((torch.isnan(out) == torch.isnan(ref)).all())
            extra = (f'  vs all-double solver: max {d.max().item():.2e} mm, bit-identical {float((d == 0).float().mean()):.3f}, '
                     f'NaN pattern {"equal" if same_nan else "DIFFERS"}')
        print(f'{name:34s} {ms:8.4f} ms {n / ms * 1e3:.4e} joints/s {gbs:7.1f} GB/s = {gbs / peak:.3f} of {peak} GB/s{extra}', flush=True)

    n = args.joints
    kp, P = make_triangulation_workload(n, 8, torch.float32, dev, seed=1)
    report('f32 V=8 clean', kp, P, check=True)
    for frac in (0.01, 0.05, 0.2):
        k2 = kp.clone()
        gen = torch.Generator(device=dev).manual_seed(3)
        bad = torch.rand((n, 8), device=dev, generator=gen) < frac
        k2[..., 2][bad] = 0.0
        k2[..., 0][bad] = 5000.0 * torch.rand((int(bad.sum()),), device=dev, generator=gen)
        k2[..., 1][bad] = 5000.0 * torch.rand((int(bad.sum()),), device=dev, generator=gen)
        report(f'f32 V=8 {frac:4.0%} views unusable', k2, P, check=True)
        del k2
    k3 = kp.transpose(1, 2).contiguous()
    report('f32 V=8 layout (N,3,V)', k3, P, layout='n3v', check=True)
    del k3
    report('f32 V=8 ragged (n - 37)', kp[:n - 37], P)
    del kp
    for V in (16, 4, 2):
        kp, P = make_triangulation_workload(n * 8 // V if V > 8 else n, V, torch.float32, dev, seed=2)
        report(f'f32 V={V}', kp, P, check=True)
        del kp
    for V in (8, 16):
        kp, P = make_triangulation_workload(n * 8 // V // 2 * 2 if V > 8 else n, V, torch.float64, dev, seed=1)
        report(f'f64 V={V}', kp, P)
        del kp


if __name__ == '__main__':
    main()
