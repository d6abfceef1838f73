"""Tuning builds of libmc3d.so: the default objects with csrc/triangulate.cu recompiled under extra -D flags.

    python profiles/build_variants.py                  # builds every variant listed below into profiles/variants/
    MC3D_LIB=profiles/variants/libmc3d_packed.so python profiles/tri_variant_check.py

The shipped library is untouched (its build is multi-camera_3d_pose_estimation_b200/build.py); a variant becomes the
default only by changing the macro's default in the source after it has been measured AND checked bit-for-bit on a B200.
"""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'multi-camera_3d_pose_estimation_b200')
OUT = os.path.join(ROOT, 'profiles', 'variants')

VARIANTS = {
    # name: extra nvcc flags for triangulate.cu
    'packed': ['-DMC3D_TRI_PACKED_SOLVE=1'],             # both solve phases of the float kernel on (joint 0, joint 1) pairs
    'lean': ['-DMC3D_TRI_LEAN=1'],                       # full-tile loop with running pointers, one barrier per tile
    'lean_packed': ['-DMC3D_TRI_LEAN=1', '-DMC3D_TRI_PACKED_SOLVE=1'],
    # rows formed as w (yx P2 + p10): two instructions fewer per joint-view, rounding differs (NOT bit-identical)
    'lean_packed_rows': ['-DMC3D_TRI_LEAN=1', '-DMC3D_TRI_PACKED_SOLVE=1', '-DMC3D_TRI_ROWS_E=1'],
    # + double residuals from the rows as given (no re-centring subtraction): equal up to float-rounding ties
    'lean_packed_rows_raw': ['-DMC3D_TRI_LEAN=1', '-DMC3D_TRI_PACKED_SOLVE=1', '-DMC3D_TRI_ROWS_E=1', '-DMC3D_TRI_RAW_RESID=1'],
    'lean_packed_raw': ['-DMC3D_TRI_LEAN=1', '-DMC3D_TRI_PACKED_SOLVE=1', '-DMC3D_TRI_RAW_RESID=1'],
    # + accepted results leave the solver as floats (9 F2F + 3 DADD per joint only for joints that need another pass)
    'lean_packed_raw_ftail': ['-DMC3D_TRI_LEAN=1', '-DMC3D_TRI_PACKED_SOLVE=1', '-DMC3D_TRI_RAW_RESID=1', '-DMC3D_TRI_FLOAT_TAIL=1'],
    # double storage: lean full-tile loop (bit-identical)
    'lean64': ['-DMC3D_TRI_LEAN64=1'],
    'all': ['-DMC3D_TRI_LEAN64=1', '-DMC3D_TRI_LEAN=1', '-DMC3D_TRI_PACKED_SOLVE=1', '-DMC3D_TRI_ROWS_E=1', '-DMC3D_TRI_RAW_RESID=1', '-DMC3D_TRI_FLOAT_TAIL=1'],
}


def main(names):
    spec = importlib.util.spec_from_file_location('mc3d_build', os.path.join(PKG, 'build.py'))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()                                             # default objects in PKG/build/
    os.makedirs(OUT, exist_ok=True)
    nvcc = b._nvcc()
    others = [os.path.join(PKG, 'build', os.path.basename(s)[:-3] + '.o') for s in b.sources()
              if not s.endswith('triangulate.cu')]
    for name in names:
        obj = os.path.join(OUT, f'triangulate_{name}.o')
        cmd = [nvcc] + b.ARCH + b.FLAGS + VARIANTS[name] + ['-c', os.path.join(PKG, 'csrc', 'triangulate.cu'), '-o', obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        with open(os.path.join(OUT, f'ptxas_{name}.log'), 'w') as fh:
            fh.write(r.stdout)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise SystemExit(f'nvcc failed for variant {name}')
        lib = os.path.join(OUT, f'libmc3d_{name}.so')
        subprocess.run([nvcc] + b.ARCH + ['-shared', '-Xcompiler', '-fPIC', '-o', lib, obj] + others + ['-lcudart'], check=True)
        print(lib)


if __name__ == '__main__':
    main(sys.argv[1:] or list(VARIANTS))
