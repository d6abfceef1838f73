"""CPU, world_size 2 over gloo: the frame-sharded refinement driver (halo exchange + scalar all-reduces of
refinement.RefineEngine / DistComm) gives the same trajectory and cost history as one process, and as the oracle.
The phases are the numpy statement in tests/cpu_phases.py (the CUDA kernels cannot run here)."""
import os
import socket

import numpy as np
import pytest

from oracle import refine as R

N_FRAMES, N_STEPS = 41, 7
KW = dict(lr=0.01, betas=(0.9, 0.999), lambda_smooth=0.3, lambda_body_length=1.0, patience=50, tolerance=1e-5, max_iter=100,
          ignore_distortions=False)


def _inputs():
    import mc3d_b200.synthetic as syn
    gs, init, cams, _ = syn.refinement_inputs(N_FRAMES, n_cams=2, seed=31)
    init[20, 4] = np.nan                                   # a masked joint in the middle of rank 0 / rank 1's boundary zone
    return gs, init, cams, dict(syn.EXAMPLE_BODY_LENGTHS)


def _run_engine(comm):
    import torch
    from cpu_phases import NumpyPhases
    from mc3d_b200 import refinement as rf
    gs, init, cams, lengths = _inputs()
    rows = rf.camera_rows(cams, list(cams))
    phases = NumpyPhases(list(cams.values()), R.bone_table(lengths))
    eng = rf.RefineEngine(init, gs, rows, lengths, torch_dtype=torch.float64, device='cpu', window=(0, N_FRAMES),
                          n_window_frames=N_FRAMES, hist_capacity=64, comm=comm, phases=phases, **KW)
    phases.engine = eng
    eng.run(N_STEPS)
    return eng.trajectory().numpy(), eng.best_trajectory().numpy(), eng.history(N_STEPS), eng.state()


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from mc3d_b200 import refinement as rf
        traj, best, hist, st = _run_engine(rf.DistComm())
        np.savez(os.path.join(out_dir, f'rank{rank}.npz'), traj=traj, best=best, hist=hist, steps=st['adam_step'])
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_frame_shard_is_a_balanced_partition():
    from mc3d_b200.refinement import frame_shard
    for n in (0, 1, 7, 100000, 100003):
        for world in (1, 2, 3, 8):
            parts = [frame_shard(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def test_numpy_phases_match_oracle():
    from mc3d_b200 import refinement as rf
    traj, best, hist, st = _run_engine(rf.LocalComm())
    gs, init, cams, lengths = _inputs()
    # the oracle has no masked-joint freezing in its Adam step, so compare on the unmasked problem
    assert st['adam_step'] == N_STEPS and np.isfinite(hist).all()
    init2 = init.copy()
    init2[20, 4] = init[19, 4]
    import torch
    from cpu_phases import NumpyPhases
    rows = rf.camera_rows(cams, list(cams))
    phases = NumpyPhases(list(cams.values()), R.bone_table(lengths))
    eng = rf.RefineEngine(init2, gs, rows, lengths, torch_dtype=torch.float64, device='cpu', window=(0, N_FRAMES),
                          n_window_frames=N_FRAMES, hist_capacity=64, phases=phases, **KW)
    phases.engine = eng
    eng.run(N_STEPS)
    ref = R.sgd_optimize(gs, init2, list(cams.values()), lengths, lr=KW['lr'], lambda_smooth=KW['lambda_smooth'],
                         lambda_body_length=KW['lambda_body_length'], max_iter=N_STEPS - 1, time_interval=(0, N_FRAMES))
    assert np.abs(eng.trajectory().numpy() - ref['final']).max() < 1e-9
    assert np.allclose(eng.history(N_STEPS)[:, 0], ref['history']['total_cost'][0::2], rtol=1e-12)


@pytest.mark.timeout(300)
def test_two_ranks_equal_one_rank(tmp_path):
    import torch.multiprocessing as mp
    from mc3d_b200 import refinement as rf
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    single = _run_engine(rf.LocalComm())
    for rank in range(2):
        got = np.load(tmp_path / f'rank{rank}.npz')
        assert int(got['steps']) == N_STEPS
        assert got['traj'].shape == (N_FRAMES, 17, 3)
        np.testing.assert_allclose(got['hist'], single[2], rtol=1e-12)
        np.testing.assert_allclose(got['traj'], single[0], rtol=0, atol=1e-10)
        np.testing.assert_allclose(got['best'], single[1], rtol=0, atol=1e-10)
