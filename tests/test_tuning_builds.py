"""CPU: the tuning builds of the float triangulation kernel (profiles/build_variants.py) still compile for sm_100a, and
their macros are off in the shipped build.  One nvcc run with every macro on (~30 s); nothing is executed."""
import importlib.util
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'multi-camera_3d_pose_estimation_b200', 'csrc', 'triangulate.cu')
MACROS = ['MC3D_TRI_LEAN', 'MC3D_TRI_PACKED_SOLVE', 'MC3D_TRI_ROWS_E', 'MC3D_TRI_RAW_RESID', 'MC3D_TRI_FLOAT_TAIL', 'MC3D_TRI_LEAN64']


def _module(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_macros_default_to_off_and_variants_only_use_known_macros():
    src = open(SRC).read()
    for m in MACROS:
        assert re.search(rf'#ifndef {m}\n#define {m} 0\n#endif', src), m
    variants = _module(os.path.join(ROOT, 'profiles', 'build_variants.py'), 'mc3d_variants').VARIANTS
    used = {f[2:].split('=')[0] for flags in variants.values() for f in flags}
    assert used <= set(MACROS), used
    build = _module(os.path.join(ROOT, 'multi-camera_3d_pose_estimation_b200', 'build.py'), 'mc3d_build_check')
    assert not any(m in ' '.join(build.FLAGS) for m in MACROS)        # the shipped library is built without them


@pytest.mark.skipif(shutil.which('nvcc') is None and not os.path.exists('/usr/local/cuda/bin/nvcc'), reason='nvcc not available')
def test_all_tuning_macros_compile(tmp_path):
    build = _module(os.path.join(ROOT, 'multi-camera_3d_pose_estimation_b200', 'build.py'), 'mc3d_build_check2')
    cmd = [build._nvcc()] + build.ARCH + build.FLAGS + [f'-D{m}=1' for m in MACROS] + ['-c', SRC, '-o', str(tmp_path / 'tri.o')]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    assert 'triangulate_mixed_lean_kernel' in r.stdout               # ptxas -v lists the lean kernel's instantiations
