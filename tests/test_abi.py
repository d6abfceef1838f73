"""CPU: libmc3d.so loads and exports exactly what include/mc3d.h declares (no compute calls)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _header_symbols():
    with open(os.path.join(ROOT, 'include', 'mc3d.h')) as fh:
        text = re.sub(r'/\*.*?\*/', '', fh.read(), flags=re.S)
    return sorted(set(re.findall(r'\b(mc3d_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def built_lib():
    import __graft_entry__ as g
    g.build()
    import mc3d_b200
    return mc3d_b200.lib()


def test_header_declares_symbols():
    syms = _header_symbols()
    assert 'mc3d_triangulate_f32' in syms and 'mc3d_last_error' in syms and len(syms) >= 9


def test_library_exports_every_declared_symbol(built_lib):
    from mc3d_b200 import _lib
    for name in _header_symbols():
        assert hasattr(built_lib, name), f'{name} declared in mc3d.h but missing from libmc3d.so'
        assert name in _lib.SIGNATURES, f'{name} has no ctypes signature in _lib.py'
    assert sorted(_lib.SIGNATURES) == _header_symbols()


def test_library_is_sm100a_only(built_lib):
    from mc3d_b200 import _lib
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


def test_version_and_error_strings(built_lib):
    assert built_lib.mc3d_version() == 100
    assert built_lib.mc3d_status_string(0) == b'ok'
    assert built_lib.mc3d_status_string(2) == b'misaligned pointer'


def test_argument_validation_happens_before_any_cuda_call(built_lib):
    """n_views out of range / NULL rig are rejected with INVALID_ARGUMENT even without a GPU."""
    import ctypes
    import numpy as np
    from mc3d_b200 import _lib
    rig, keep = _lib.make_rig(np.zeros((1, 3, 4)))
    st = built_lib.mc3d_triangulate_f32(None, 10, ctypes.byref(rig), 0, 0, 0, None, None)
    assert st == 1 and b'n_views' in built_lib.mc3d_last_error()
    st = built_lib.mc3d_triangulate_f64(None, 10, None, 0, 0, 0, None, None)
    assert st == 1
    rig, keep = _lib.make_rig(np.zeros((2, 3, 4)))
    assert built_lib.mc3d_triangulate_f32(None, 10, ctypes.byref(rig), 7, 0, 0, None, None) == 1
    assert built_lib.mc3d_triangulate_f32(None, 0, ctypes.byref(rig), 0, 0, 0, None, None) == 0     # empty is fine
    assert built_lib.mc3d_triangulate_f32(ctypes.c_void_p(8), 5, ctypes.byref(rig), 0, 0, 0, ctypes.c_void_p(16), None) == 2


def test_no_cpu_fallback():
    """Without a CUDA device the product raises instead of computing on the CPU."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    import mc3d_b200
    import mc3d_b200.utils as u
    with pytest.raises(mc3d_b200.Mc3dError):
        u.DLT(np.eye(3, 4), np.eye(3, 4), [1.0, 2.0], [3.0, 4.0])
    with pytest.raises(mc3d_b200.Mc3dError):
        from mc3d_b200.triangulation import triangulate_multiview
        triangulate_multiview(torch.zeros(4, 2, 3), np.zeros((2, 3, 4)))
    # every other entry point of the path: heatmap decode, projection, interpolation, refinement (class and engine)
    from mc3d_b200.mmpose_pose_estimation import PoseEstimator
    with pytest.raises(mc3d_b200.Mc3dError):
        PoseEstimator.get_heatmap_means_cov(None, np.zeros((2, 64, 48), dtype=np.float32))
    import mc3d_b200.pose_refinement as pr
    with pytest.raises(mc3d_b200.Mc3dError):
        pr.project_points_torch(np.zeros((2, 17, 3)), np.eye(3), np.eye(3), np.zeros((3, 1)), np.zeros((1, 5)))
    with pytest.raises(mc3d_b200.Mc3dError):
        pr.linear_interpolation(np.zeros((12, 17, 3)))
    import mc3d_b200.synthetic as syn
    gs, init, cams, _ = syn.refinement_inputs(8, seed=1)
    opt = pr.Optimized_3d_Pose_Estimation(gs, init, decomposed_cam_params_initial={i: list(cams[i]) for i in cams},
                                          body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS))
    with pytest.raises(mc3d_b200.Mc3dError):
        opt.sgd_optimize(max_iter=1, print_frequency=np.inf)
    from mc3d_b200.refinement import RefineEngine, camera_rows
    with pytest.raises(mc3d_b200.Mc3dError):
        RefineEngine(init, gs, camera_rows(cams, [0, 1]), dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=torch.float64,
                     device='cpu', lr=1e-3, betas=(0.9, 0.999), lambda_smooth=1.0, lambda_body_length=1.0, patience=10,
                     tolerance=1e-5, max_iter=1, ignore_distortions=False, window=(0, 8), n_window_frames=8,
                     hist_capacity=4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'multi-camera_3d_pose_estimation_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), os.path.join(dirpath, f)


def test_header_is_valid_c_and_links_from_c(built_lib, tmp_path):
    """tests/c/abi_check.c (C99, -pedantic): includes mc3d.h, takes the address of all 35 entry points, compares the struct
    sizes a C compiler sees with the library's, and checks one argument-validation path -- no compute, no GPU."""
    from mc3d_b200 import _lib
    exe = str(tmp_path / 'abi_check')
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', os.path.join(ROOT, 'include'),
           os.path.join(ROOT, 'tests', 'c', 'abi_check.c'), '-o', exe, '-L', libdir, '-lmc3d', f'-Wl,-rpath,{libdir}']
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.startswith(f'ok {len(_header_symbols())} entry points')
