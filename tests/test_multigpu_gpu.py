"""GPU, >= 2 devices: frame-sharded paths under torchrun (skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.timeout(900)
def test_sharded_refinement_equals_single_gpu():
    if _n_gpus() < 2:
        pytest.skip('needs two GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29541', os.path.join(ROOT, 'tests', 'mgpu_refine_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=800)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'ok=True' in r.stdout


@pytest.mark.timeout(900)
def test_sharded_refinement_with_repeated_pass1_equals_single_gpu():
    """The repetition of pass 1 (counts of finite terms declared changed at every fourth step: mc3d_refine_problem.test_flags)
    exchanges its sums through the ranks' retry slots: sharded run == single-GPU run with the same hook."""
    if _n_gpus() < 2:
        pytest.skip('needs two GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29545', os.path.join(ROOT, 'tests', 'mgpu_refine_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=800, env=dict(os.environ, MC3D_REFINE_TEST_FLAGS='1'))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'ok=True' in r.stdout and 'peer_exchange=True' in r.stdout


@pytest.mark.timeout(900)
def test_sharded_refinement_falls_back_to_host_driven_exchange():
    """When a peer's exchange block cannot be mapped (MC3D_PEER_FAIL=1 simulates it) every rank switches to the NCCL path."""
    if _n_gpus() < 2:
        pytest.skip('needs two GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29543', os.path.join(ROOT, 'tests', 'mgpu_refine_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=800, env=dict(os.environ, MC3D_PEER_FAIL='1'))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'ok=True' in r.stdout and 'peer_exchange=False' in r.stdout


@pytest.mark.timeout(900)
def test_sharded_triangulation_bench_line():
    if _n_gpus() < 2:
        pytest.skip('needs two GPUs')
    import json
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29542', os.path.join(ROOT, 'bench.py'), '--gpus', '2', '--steps', '3', '--warmup', '3',
           '--workload', 'tri8_coco17_1Mframes_f32']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=800)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1])
    assert line['n_gpus'] == 2 and line['scaling'] == 'weak' and line['value'] > 1e10
    assert line['e2e']['matches_device_path']


def test_host_pipeline_leaves_the_current_device_alone():
    """mc3d_triangulate_host_* / mc3d_decode_heatmaps_host_f32 work on the device they are given and restore the calling
    thread's current device; with device=None they follow torch's current device."""
    if _n_gpus() < 2:
        pytest.skip('needs two GPUs')
    import numpy as np
    import torch
    from mc3d_b200 import synthetic as syn
    from mc3d_b200.decode import decode_heatmaps
    from mc3d_b200.triangulation import triangulate_multiview
    kp, P, _, _ = syn.multiview_points(3000, 4, seed=2)
    torch.cuda.set_device(1)
    ref = triangulate_multiview(kp, P, device=0)
    assert torch.cuda.current_device() == 1
    got = triangulate_multiview(kp, P)                       # device=None -> cuda:1
    assert torch.cuda.current_device() == 1 and np.array_equal(ref, got)
    hm, _ = syn.gaussian_blob_heatmaps(34, seed=3)
    k0, m0 = decode_heatmaps(hm, device=0)
    assert torch.cuda.current_device() == 1
    k1, m1 = decode_heatmaps(hm)
    assert np.array_equal(k0, k1) and np.array_equal(m0, m1)
    torch.cuda.set_device(0)
