"""A/B timing of builds of the 64x48 decode kernel on one B200: the default library and every
profiles/variants/libmc3d_*.so (python profiles/decode_ab.py build NAME=-DMACRO=V ... builds them from csrc/decode.cu)."""
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
    others = [os.path.join(PKG, 'build', os.path.basename(s)[:-3] + '.o') for s in b.sources() if not s.endswith('decode.cu')]
    for sp in specs:
        name, flags = sp.split('=', 1)
        obj = os.path.join(OUT, f'decode_{name}.o')
        r = subprocess.run([nvcc] + b.ARCH + b.FLAGS + flags.split(',') + ['-c', os.path.join(PKG, 'csrc', 'decode.cu'), '-o', obj],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        open(os.path.join(OUT, f'ptxas_{name}.log'), 'w').write(r.stdout)
        if r.returncode:
            raise SystemExit(r.stdout)
        lib = os.path.join(OUT, f'libmc3d_{name}.so')
        subprocess.run([nvcc] + b.ARCH + ['-shared', '-Xcompiler', '-fPIC', '-o', lib, obj] + others + ['-lcudart'], check=True)
        print(lib)


def run():
    import torch
    from mc3d_b200 import _lib
    dev = 'cuda:0'
    n = 200_000
    hm = torch.rand((n, 64, 48), device=dev) * 0.05
    hm[:, 30:34, 20:24] += 0.8
    kpt = torch.empty((n, 3), dtype=torch.float32, device=dev)
    mom = torch.empty((n, 6), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    libs = [('default', _lib.LIB_PATH)] + [(os.path.basename(p)[8:-3], p) for p in sorted(glob.glob(os.path.join(OUT, 'libmc3d_*.so')))]
    ref = None
    for name, path in libs:
        h = ctypes.CDLL(path)
        fn = h.mc3d_decode_heatmaps_f32
        fn.restype = ctypes.c_int
        fn.argtypes = _lib.SIGNATURES['mc3d_decode_heatmaps_f32'][1]
        for label, mptr in (('kpts+moments', mom.data_ptr()), ('kpts only', None)):
            def go():
                st = fn(hm.data_ptr(), n, 64, 48, 0.01, 0, 0, 0, 0, None, 0, kpt.data_ptr(), mptr, stream)
                assert st == 0, st
            for _ in range(3):
                go()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                go()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            extra = ''
            if mptr:
                if ref is None:
                    ref = mom.clone()
                else:
                    extra = f'  max |moment diff| vs default {float((mom - ref).abs().max()):.2e}'
            print(f'{name:16s} {label:13s} {ms:7.4f} ms  {n / ms * 1e3:.4e} maps/s  {n * 12288 / ms / 1e6:7.1f} GB/s = {n * 12288 / ms / 1e6 / 6451.8:.3f}{extra}', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'build':
        build(sys.argv[2:])
    else:
        run()
