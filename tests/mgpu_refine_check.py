"""Run under torchrun (one rank per GPU): the NCCL frame-sharded refinement equals the single-GPU run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_refine_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import mc3d_b200.pose_refinement as pr
    import mc3d_b200.synthetic as syn
    from mc3d_b200 import refinement as rf
    rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    n_frames, iters = 3001, 40
    gs, init, cams, _ = syn.refinement_inputs(n_frames, n_cams=2, seed=41)
    init[1500, 7] = np.nan                      # a masked joint exactly at the 2-rank shard boundary
    kw = dict(lr=0.01, lambda_smooth=1e-3, lambda_body_length=1.0, max_iter=iters - 1, time_interval=[0, n_frames],
              patience=10 ** 6, print_frequency=np.inf)
    mk = lambda: pr.Optimized_3d_Pose_Estimation(gs.copy(), init.copy(), decomposed_cam_params_initial={i: list(cams[i]) for i in cams},
                                                 body_lengths=dict(syn.EXAMPLE_BODY_LENGTHS), torch_dtype=torch.float64,
                                                 device=f'cuda:{local}')
    single = mk()
    single.sgd_optimize(**kw)                   # torch.distributed not initialised yet: one GPU, CUDA graph path
    dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    sharded = mk()
    sharded.sgd_optimize(**kw)
    h1 = np.array([float(v) for v in single.all_costs_total['total_cost']])
    h2 = np.array([float(v) for v in sharded.all_costs_total['total_cost']])
    ok = len(h1) == len(h2) == 2 * iters and np.allclose(h1, h2, rtol=1e-11)
    dtraj = np.nanmax(np.abs(single.trajectory.numpy() - sharded.trajectory.numpy()))
    dbest = np.nanmax(np.abs(single.best_trajectory.numpy() - sharded.best_trajectory.numpy()))
    ok = ok and dtraj < 1e-9 and dbest < 1e-9 and tuple(sharded.trajectory.shape) == (n_frames, 17, 3)
    ok = ok and bool(np.isnan(sharded.trajectory.numpy()[1500, 7]).all())
    ok = ok and sharded._engine.peer is None                 # sgd_optimize closed the exchange
    # iterations per second of the sharded path: in-kernel exchange over NVLink peer memory against the host-driven
    # NCCL exchange (CUDA graph and eager)
    import time
    rows = rf.camera_rows(cams, list(cams))

    def engine():
        return rf.RefineEngine(init, gs, rows, syn.EXAMPLE_BODY_LENGTHS, torch_dtype=torch.float32, device=f'cuda:{local}', lr=0.01,
                               betas=(0.9, 0.999), lambda_smooth=1e-3, lambda_body_length=1.0, patience=10 ** 9, tolerance=1e-5,
                               max_iter=10 ** 9, ignore_distortions=False, window=(0, n_frames), n_window_frames=n_frames,
                               hist_capacity=64, comm=rf.DistComm())

    def rate_of(eng, n):
        eng.run(20)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        eng.run(n)
        torch.cuda.synchronize()
        dist.barrier()
        return n / (time.perf_counter() - t0)

    eng = engine()
    peer_on = eng.peer is not None
    rate_peer = rate_of(eng, 400)
    eng.close()
    os.environ['MC3D_REFINE_PEER'] = '0'
    eng = engine()
    rate = rate_of(eng, 200)
    graph_on = eng._graph is not None
    eng.use_graph = False
    rate_eager = rate_of(eng, 50)
    eng.close()
    del os.environ['MC3D_REFINE_PEER']
    flag = torch.tensor([1.0 if ok else 0.0], device=f'cuda:{local}')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f'MGPU_REFINE world={dist.get_world_size()} ok={bool(flag.item())} dtraj={dtraj:.2e} dbest={dbest:.2e} '
              f'hist_rel={np.max(np.abs(h1 - h2) / np.abs(h1)):.2e} peer_exchange={peer_on} iters_per_s: in-kernel={rate_peer:.0f} '
              f'nccl_graph={rate:.0f} (graph={graph_on}) nccl_eager={rate_eager:.0f}')
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == '__main__':
    main()
