"""A/B timing of tuning builds of the float triangulation kernel on one B200.

    python profiles/tri_ab.py build NAME=-DMACRO=1,-DOTHER=2 ...   # here (no GPU): profiles/variants/libmc3d_NAME.so
    python profiles/tri_ab.py run [--views 8] [--joints 17000000]   # on the GPU box: default library + every variant

`run` prints CUDA-event times and whether a variant's output is bit-identical to the default library's.  Nothing printed
here is a bench value of record: a variant that wins becomes the default in the source and bench.py measures it.
"""
import ctypes
import glob
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'multi-camera_3d_pose_estimation_b200')
OUT = os.path.join(ROOT, 'profiles', 'variants')
sys.path.insert(0, ROOT)


def build(specs):
    spec = importlib.util.spec_from_file_location('mc3d_build', os.path.join(PKG, 'build.py'))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()
    os.makedirs(OUT, exist_ok=True)
    nvcc = b._nvcc()
    others = [os.path.join(PKG, 'build', os.path.basename(s)[:-3] + '.o') for s in b.sources() if not s.endswith('triangulate.cu')]
    procs = []
    for sp in specs:
        name, flags = sp.split('=', 1)
        obj = os.path.join(OUT, f'triangulate_{name}.o')
        cmd = [nvcc] + b.ARCH + b.FLAGS + flags.split(',') + ['-c', os.path.join(PKG, 'csrc', 'triangulate.cu'), '-o', obj]
        procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, obj, p in procs:
        out, _ = p.communicate()
        with open(os.path.join(OUT, f'ptxas_{name}.log'), 'w') as fh:
            fh.write(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise SystemExit(f'nvcc failed for {name}')
        lib = os.path.join(OUT, f'libmc3d_{name}.so')
        subprocess.run([nvcc] + b.ARCH + ['-shared', '-Xcompiler', '-fPIC', '-o', lib, obj] + others + ['-lcudart'], check=True)
        regs = [l for l in out.splitlines() if 'mixed_kernelILi8ELi0' in l or 'Used' in l]
        print(lib)


def run(argv):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--joints', type=int, default=17_000_000)
    ap.add_argument('--views', type=int, default=8)
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--dtype', choices=['f32', 'f64'], default='f32')
    args = ap.parse_args(argv)
    import torch
    import bench
    from mc3d_b200 import _lib
    dev = torch.device('cuda:0')
    tdt = torch.float32 if args.dtype == 'f32' else torch.float64
    idt = torch.int32 if args.dtype == 'f32' else torch.int64
    kp, P = bench.make_triangulation_workload(args.joints, args.views, tdt, dev, seed=0)
    rig, keep = _lib.make_rig(P)
    libs = [('default', _lib.LIB_PATH)] + [(os.path.basename(p)[8:-3], p) for p in sorted(glob.glob(os.path.join(OUT, 'libmc3d_*.so')))]
    ref = None
    stream = torch.cuda.current_stream().cuda_stream
    for name, path in libs:
        h = ctypes.CDLL(path)
        fn = getattr(h, f'mc3d_triangulate_{args.dtype}')
        fn.restype = ctypes.c_int
        fn.argtypes = _lib.SIGNATURES[f'mc3d_triangulate_{args.dtype}'][1]
        out = torch.empty((args.joints, 3), dtype=tdt, device=dev)

        def go():
            st = fn(kp.data_ptr(), args.joints, ctypes.byref(rig), _lib.LAYOUT_V3, _lib.TRI_WEIGHTED, 0, out.data_ptr(), stream)
            assert st == 0, st
        for _ in range(3):
            go()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            go()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        if ref is None:
            ref, same = out.clone(), 'reference'
        else:
            ok = torch.isfinite(out).all(dim=1) & torch.isfinite(ref).all(dim=1)
            worst = (out[ok].double() - ref[ok].double()).abs().max().item()
            same = 'bit-identical' if torch.equal(out.view(idt), ref.view(idt)) else f'max |diff| {worst:.2e} mm'
        print(f'{name:24s} {ms:8.4f} ms  {args.joints / ms * 1e3:.4e} joints/s  {same}', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'build':
        build(sys.argv[2:])
    else:
        run(sys.argv[2:] if len(sys.argv) > 1 and sys.argv[1] == 'run' else sys.argv[1:])
