#!/usr/bin/env python
"""bench.py -- headline benchmark of the mc3d hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workload (BASELINE.json configs[1]): 8-camera COCO-17 confidence-weighted DLT triangulation of 10 M
synthetic frames (1.7e8 joints), float storage.  A "step" is one pass of the triangulation kernel over the
whole batch.  Metric: 3D joints triangulated per second (whole job, all ranks).

  value      inputs resident in HBM, CUDA events on the launch stream, K steps, max over ranks
  e2e        the same batch through the host-buffer C ABI (mc3d_triangulate_host_f32): pinned host
             keypoints -> H2D -> kernel -> D2H -> pinned host result, every step
  roofline   algorithmic bytes per launch (108 B/joint: 8 views x 3 floats in + 3 floats out, SURVEY.md
             section 8d) / average launch time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the numpy oracle port (batched LAPACK SVD of A^T A, utils.py:19-34 generalised) on the
             host cores, bounded sample; reported, not the target

One JSON line on stdout (rank 0).  Multi-GPU: frames are sharded across ranks, no data-path collective
(weak scaling: every rank triangulates its own 10 M frames).

Beside the headline the line carries, at EVERY --gpus N (each rank works on its shard, times are the max over ranks):
  other_configs  config 2 in float64 at its full 10 M frames, config 3 (decode 17x64x48 x 4 views + triangulation) and
                 config 5 (16 views, point-count sweep 1e6..1e9 sharded over the ranks), each with its own roofline block
  refine         config 4: iterations/s of the 100 000-frame refinement (sharded over the ranks when N > 1) with a roofline
                 block, and -- N > 1 -- `vs_single_gpu`: the same problem run on ONE GPU by rank 0 and compared with the
                 sharded run (cost history and trajectory); the process exits non-zero when they differ by more than 1e-6
  e2e.host_link  H2D / D2H / both-ways pinned-copy bandwidth of this box measured beside the e2e number (its ceiling)
and, at N = 1 only: `occlusion` (throughput at 0 / 1 / 5 / 20 % unusable views) and the CPU baselines of all three kernels.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

JOINTS = 17
WORKLOADS = {
    # name: (frames per GPU, views, io dtype)
    'tri8_coco17_10Mframes_f32': (10_000_000, 8, 'f32'),
    'tri8_coco17_10Mframes_f64': (10_000_000, 8, 'f64'),
    'tri8_coco17_1Mframes_f32': (1_000_000, 8, 'f32'),
}
DEFAULT_WORKLOAD = 'tri8_coco17_10Mframes_f32'


def make_triangulation_workload(n_joints, n_views, dtype, device, seed=0):
    """Synthetic config-2 input generated on the device: ring rig of ``n_views`` cameras, X ~ N(centre, 400 mm),
    pixels = projection + N(0, 1 px), w ~ U(0.2, 1).  Returns (kpts (n, V, 3) on ``device``, P (V,3,4) numpy)."""
    import torch
    from mc3d_b200 import synthetic as syn
    cams = syn.ring_rig(n_views)
    P = syn.projection_matrices(cams)
    gen = torch.Generator(device=device).manual_seed(seed)
    kp = torch.empty((n_joints, n_views, 3), dtype=dtype, device=device)
    centre = torch.tensor(syn.SCENE_CENTRE, dtype=torch.float64, device=device)
    Pt = torch.tensor(P, dtype=torch.float64, device=device)
    chunk = 1 << 21
    for lo in range(0, n_joints, chunk):
        m = min(chunk, n_joints - lo)
        X = centre + 400.0 * torch.randn((m, 3), dtype=torch.float64, device=device, generator=gen)
        Xh = torch.cat([X, torch.ones((m, 1), dtype=torch.float64, device=device)], dim=1)
        proj = torch.einsum('vij,nj->nvi', Pt, Xh)                     # (m, V, 3)
        uv = proj[..., :2] / proj[..., 2:3] + torch.randn((m, n_views, 2), dtype=torch.float64, device=device,
                                                          generator=gen)
        w = 0.2 + 0.8 * torch.rand((m, n_views, 1), dtype=torch.float64, device=device, generator=gen)
        kp[lo:lo + m] = torch.cat([uv, w], dim=2).to(dtype)
    return kp, P


# ---- clocks ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.gpu_index}', f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                clk, mx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.1 <= ts <= t1 + 0.3:
                sm.append(clk)
                for nm, val in zip(names, parts[3:7]):
                    if val.lower().startswith('active'):
                        reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax, 'samples': len(sm),
                'reasons': sorted(reasons)}


# ---- CPU baseline (oracle port) ----------------------------------------------------------------------------
def _cpu_worker(args):
    n, n_views, seed, reps = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from mc3d_b200 import synthetic as syn
    from oracle import dlt as O
    kp, P, _, _ = syn.multiview_points(n, n_views, seed=seed)
    kp = kp.astype(np.float32).astype(np.float64)
    t0 = time.perf_counter()
    for _ in range(reps):
        O.dlt_weighted(kp, P)
    return n * reps, time.perf_counter() - t0


def cpu_baseline_run(n_views, per_worker=100_000, reps=32, workers=None):
    """Oracle port over all host cores (one process per core, frames split in blocks)."""
    import multiprocessing as mp
    workers = workers or (os.cpu_count() or 1)
    ctx = mp.get_context('spawn')
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        # first map warms the workers (imports) and is not timed
        pool.map(_cpu_worker, [(256, n_views, 1, 1)] * workers)
        res = pool.map(_cpu_worker, [(per_worker, n_views, 100 + i, reps) for i in range(workers)])
        wall = max(r[1] for r in res)                        # the workers time the DLT only (input generation excluded)
    joints = sum(r[0] for r in res)
    return joints / wall, workers, joints, wall


def cpu_loop_rate(n_views, n=4000):
    """The reference's own style: one scipy SVD per joint in a Python loop (utils.py:19-34), one core."""
    from mc3d_b200 import synthetic as syn
    from oracle import dlt as O
    kp, P, _, _ = syn.multiview_points(n, n_views, seed=7)
    t0 = time.perf_counter()
    O.dlt_weighted_loop(kp, P)
    return n / (time.perf_counter() - t0)


def _cpu_decode_worker(args):
    n_maps, seed, reps = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from mc3d_b200 import synthetic as syn
    from oracle import decode as D
    hm, _ = syn.gaussian_blob_heatmaps(n_maps, seed=seed)
    t0 = time.perf_counter()
    for _ in range(reps):
        D.heatmap_means_cov(hm.copy())                   # upstream thresholds its input in place (Q7): a fresh copy per call
        D.argmax_decode(hm)
    return n_maps * reps, time.perf_counter() - t0


def cpu_decode_baseline(per_worker=1700, reps=100, workers=None):
    """Oracle port of get_heatmap_means_cov (mmpose_pose_estimation.py:163-215) + the argmax decode, all host cores."""
    import multiprocessing as mp
    workers = workers or (os.cpu_count() or 1)
    with mp.get_context('spawn').Pool(workers) as pool:
        pool.map(_cpu_decode_worker, [(17, 1, 1)] * workers)          # imports, untimed
        res = pool.map(_cpu_decode_worker, [(per_worker, 50 + i, reps) for i in range(workers)])
    wall = max(r[1] for r in res)
    maps = sum(r[0] for r in res)
    return {'value': maps / wall, 'unit': 'heatmaps/s', 'cores': workers, 'kind': 'port',
            'sample': f'{maps} maps of 64x48 ({maps // workers} per worker x {workers} workers), {wall:.1f} s; numpy oracle of '
                      'get_heatmap_means_cov + argmax decode', 'frames_per_s_4views_17joints': maps / wall / 68.0}


def cpu_refine_baseline():
    """Oracle port of sgd_optimize (closed-form numpy loss + gradient + Adam, pose_refinement.py:894-1096) on one host
    core at the sizes BASELINE.md section 3.1 names; the reference itself (torch autograd + a Python loop over frames)
    ran ~10 it/s at T = 400 in the survey's probe."""
    from mc3d_b200 import synthetic as syn
    from oracle import refine as R
    out = {}
    for T_, iters in ((400, 600), (4000, 120)):
        gs, init, cams, _ = syn.refinement_inputs(T_, n_cams=2, seed=0)
        kw = dict(lr=0.01, lambda_smooth=1e-6, lambda_body_length=1.0, time_interval=[0, T_], patience=10 ** 9, dtype=np.float32)
        R.sgd_optimize(gs, init, list(cams.values()), syn.EXAMPLE_BODY_LENGTHS, max_iter=1, **kw)      # warm-up
        t0 = time.perf_counter()
        R.sgd_optimize(gs, init, list(cams.values()), syn.EXAMPLE_BODY_LENGTHS, max_iter=iters - 1, **kw)
        dt = time.perf_counter() - t0
        out[f'T{T_}'] = {'value': iters / dt, 'unit': 'iterations/s', 'cores': 1, 'kind': 'port',
                         'sample': f'{iters} iterations of {T_} frames x 17 joints x 2 cameras, {dt:.1f} s; numpy oracle of sgd_optimize'}
    return out


def workload_config(name):
    """`config` of the JSON line: the same dict in both arms (the driver compares them)."""
    frames, n_views, io = WORKLOADS[name]
    esize = 4 if io == 'f32' else 8
    n = frames * JOINTS
    return {'workload': name, 'views': n_views, 'joints_per_frame': JOINTS, 'frames_per_gpu': frames,
            'joints_per_gpu': n, 'io_dtype': io, 'layout': '(N, V, 3) [x, y, w]', 'mode': 'weighted',
            'l2': f'inputs {n * 3 * n_views * esize / 1e9:.1f} GB per GPU >> 126 MB L2 (no flush needed)',
            'sharding': 'frames across ranks, no collective'}


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            with open(path) as fh:
                return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def recorded_traffic(workload):
    """dram__bytes_read+write per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(path):
        try:
            with open(path) as fh:
                return json.load(fh).get(workload)
        except Exception:
            return None
    return None


# ---- reference arm ------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference DLT on all host cores; every step is a bounded sample of the
    workload (per_worker joints per core).  One worker pool for the whole run."""
    if rank != 0:
        return
    import multiprocessing as mp
    frames, n_views, io = WORKLOADS[args.workload]
    per_worker = 200_000
    workers = os.cpu_count() or 1
    t_all = time.perf_counter()
    total, wall = 0, 0.0
    with mp.get_context('spawn').Pool(workers) as pool:
        pool.map(_cpu_worker, [(256, n_views, 1, 1)] * workers)            # imports, untimed
        for step in range(args.warmup + args.steps):
            res = pool.map(_cpu_worker, [(per_worker, n_views, 1000 * step + i, 1) for i in range(workers)])
            dt = max(r[1] for r in res)                     # the workers time the DLT only (input generation excluded)
            if step >= args.warmup:
                total += sum(r[0] for r in res)
                wall += dt
    value = total / wall
    line = {
        'impl': 'reference', 'metric': 'joints_triangulated_per_sec', 'value': value, 'unit': 'joints/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * wall / max(1, args.steps),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.workload),
        'note': 'CPU arm: numpy oracle port of the reference DLT (batched LAPACK SVD of A^T A, utils.py:19-34 generalised to V '
                f'weighted views, float64 arithmetic on the float32-rounded input) on a bounded sample of {per_worker} joints per '
                'worker per step of the same workload; the reference itself is pure Python without a V-view path and is not '
                'pip-installable (no setup.py)',
        'cpu_baseline': {'value': value, 'unit': 'joints/s', 'cores': workers, 'kind': 'port',
                         'sample': f'{per_worker} joints x {workers} workers per step, {args.steps} steps'},
        'e2e': {'value': value, 'unit': 'joints/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


# ---- our arm ------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import mc3d_b200
    from mc3d_b200.triangulation import triangulate_multiview

    torch.cuda.set_device(local_rank)
    device = f'cuda:{local_rank}'
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(device))
    frames, n_views, io = WORKLOADS[args.workload]
    tdtype = torch.float32 if io == 'f32' else torch.float64
    esize = 4 if io == 'f32' else 8
    n = frames * JOINTS
    kp, P = make_triangulation_workload(n, n_views, tdtype, device, seed=1000 + rank)
    out = torch.empty((n, 3), dtype=tdtype, device=device)
    algo_bytes_per_joint = (3 * n_views + 3) * esize

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        triangulate_multiview(kp, P, out=out)

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = mc3d_b200.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = mc3d_b200.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * n * args.steps / (ms * 1e-3)
    ms_per_step = ms / args.steps
    achieved = algo_bytes_per_joint * n / (ms_per_step * 1e-3) / 1e9       # GB/s per GPU (per launch)

    # ---- e2e: host buffers through the C ABI, every step --------------------------------------------------
    e2e = None
    try:
        avail = _mem_available_bytes()
        need = n * (3 * n_views + 3) * esize
        local_world = int(os.environ.get('LOCAL_WORLD_SIZE', world))
        n_e2e = n
        while n_e2e > (1 << 20) and need * local_world * 2.5 > avail:
            n_e2e //= 2
            need = n_e2e * (3 * n_views + 3) * esize
        h_kp = torch.empty((n_e2e, n_views, 3), dtype=tdtype, pin_memory=True)
        h_out = torch.empty((n_e2e, 3), dtype=tdtype, pin_memory=True)
        h_kp.copy_(kp[:n_e2e])
        torch.cuda.synchronize()
        kp_np, out_np = h_kp.numpy(), h_out.numpy()
        e2e_steps = max(2, min(args.steps, 5))
        triangulate_multiview(kp_np, P, out=out_np, device=local_rank)            # warm-up (allocates the ring)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            triangulate_multiview(kp_np, P, out=out_np, device=local_rank)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        check = bool(np.array_equal(out_np[:4096], out[:4096].cpu().numpy()))
        e2e = {'value': world * n_e2e * e2e_steps / dt, 'unit': 'joints/s',
               'h2d_bytes_per_step': int(n_e2e * 3 * n_views * esize), 'd2h_bytes_per_step': int(n_e2e * 3 * esize),
               'steps': e2e_steps, 'joints_per_step_per_gpu': int(n_e2e), 'ms_per_step': 1e3 * dt / e2e_steps,
               'matches_device_path': check,
               'api': 'mc3d_triangulate_host_f32 (pinned host in/out, 3-deep H2D/kernel/D2H pipeline)'
               if io == 'f32' else 'mc3d_triangulate_host_f64'}
        del h_kp, h_out
    except Exception as exc:            # report, do not hide
        e2e = {'value': None, 'unit': 'joints/s', 'error': repr(exc)}

    # ---- the box's host <-> device link, measured beside the e2e number (all ranks at once) ------------------------
    if e2e is not None and e2e.get('value'):
        try:
            link = host_link_bandwidth(device, world)
            ceiling = link['h2d_GBs'] * 1e9 / (3 * n_views * esize)
            e2e['host_link'] = link
            e2e['ceiling_joints_per_s'] = ceiling
            e2e['frac_of_ceiling'] = e2e['value'] / ceiling
            e2e['ceiling_note'] = ('every joint needs 3V scalars host -> device and 3 back; the ceiling is the measured aggregate pinned '
                                   'H2D bandwidth of this box with all ranks copying at once, divided by the input bytes per joint')
        except Exception as exc:
            e2e['host_link'] = {'error': repr(exc)}

    del kp, out
    torch.cuda.empty_cache()
    peak, peak_src = measured_peak()
    refine = None
    mgpu_ok = True
    if not args.no_refine:
        try:
            refine = {}
            key = 'T100k_f32' if world == 1 else f'T100k_f32_sharded_over_{world}_gpus'
            refine[key] = refine_benchmark(100_000, 400, 'f32', device, world=world, peak=peak)
            if world == 1:
                refine['T400_f32'] = refine_benchmark(400, 2000, 'f32', device, peak=peak)
                for nf in (50_000, 25_000, 12_500):     # the per-GPU shard sizes of config 4 at 2 / 4 / 8 GPUs, alone on one GPU
                    refine[f'T{nf}_f32'] = refine_benchmark(nf, 400, 'f32', device, peak=peak)
            else:
                refine['vs_single_gpu'] = refine_sharded_vs_single(device, rank, world)
                mgpu_ok = bool(refine['vs_single_gpu'].get('ok', False))
        except Exception as exc:
            refine = {'error': repr(exc)}
            mgpu_ok = world == 1
    extras = None
    if not args.no_extras:
        try:
            extras = extra_measurements(device, peak, rank, world)
        except Exception as exc:
            extras = {'error': repr(exc)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if not mgpu_ok:
            sys.exit(3)
        return

    line = {
        'metric': 'joints_triangulated_per_sec', 'value': value, 'unit': 'joints/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(3, args.warmup), 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32+f64' if io == 'f32' else 'f64',     # f32 storage: float normal equations, double residuals
        'data': 'synthetic',
        'config': workload_config(args.workload),
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': recorded_traffic(args.workload), 'peak_source': peak_src,
                     'algorithmic_bytes_per_joint': algo_bytes_per_joint,
                     'kernel': f'mc3d::triangulate_mixed_kernel<{n_views}, layout>' if io == 'f32' else f'mc3d::triangulate_lean64_kernel<{n_views}, layout>',
                     'frac_of_nominal_8TBs': achieved / 8000.0},
        'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks,
    }
    if refine is not None:
        line['refine'] = refine
    if extras is not None:
        line['other_configs'] = extras
    if world == 1 and not args.no_extras:
        try:
            line['occlusion'] = occlusion_sweep(device, n_views, tdtype, peak, algo_bytes_per_joint)
        except Exception as exc:
            line['occlusion'] = {'error': repr(exc)}
    if world == 1 and not args.no_cpu:
        rate, workers, joints, wall = cpu_baseline_run(n_views)
        line['cpu_baseline'] = {'value': rate, 'unit': 'joints/s', 'cores': workers, 'kind': 'port',
                                'sample': f'{joints} joints ({joints // workers} per worker x {workers} workers), '
                                          f'{wall:.1f} s; numpy oracle: batched LAPACK SVD of A^T A',
                                'reference_style_loop_joints_per_s_1core': cpu_loop_rate(n_views)}
        try:                                             # the other two kernels' CPU figures, same box
            dec = cpu_decode_baseline()
            ref = cpu_refine_baseline()
            line['cpu_baseline']['decode'] = dec
            line['cpu_baseline']['refine'] = ref
            if extras and 'decode_tri4_coco17' in extras:
                extras['decode_tri4_coco17']['cpu_baseline'] = dec
            if refine and 'T400_f32' in refine:
                refine['T400_f32']['cpu_baseline'] = ref['T400']
                refine['T100k_f32']['cpu_baseline'] = dict(ref['T4000'], note='largest size the CPU port is timed at (T = 4 000); its time '
                                                           'per iteration grows linearly with T')
        except Exception as exc:
            line['cpu_baseline']['others_error'] = repr(exc)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not mgpu_ok:
        sys.exit(3)


def _max_over_ranks(value, device, world):
    if world <= 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def _sync_all(world):
    import torch
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()


def host_link_bandwidth(device, world, nbytes=256 << 20, reps=4):
    """Aggregate pinned-memory copy bandwidth of this box, every rank copying at the same time: H2D alone, D2H alone, both."""
    import torch
    h_a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=device)
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)

    def timed(fn):
        fn()
        _sync_all(world)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = _max_over_ranks(time.perf_counter() - t0, device, world)
        _sync_all(world)
        return world * nbytes * reps / dt / 1e9

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    out = {'h2d_GBs': timed(h2d), 'd2h_GBs': timed(d2h), 'bidirectional_GBs_each_way': timed(both), 'ranks': world,
           'bytes_per_copy': nbytes, 'note': 'aggregate over the ranks, pinned host memory, cudaMemcpyAsync'}
    return out


def occlusion_sweep(device, n_views, tdtype, peak, bytes_per_joint, n=17_000_000):
    """Throughput when a fraction of all views is unusable (weight 0, wild pixel): a joint whose first starting pair holds such
    a view takes the second pair; a warp pays another pass only when both pairs of one of its joints fail."""
    import torch
    from mc3d_b200.triangulation import triangulate_multiview
    kp, P = make_triangulation_workload(n, n_views, tdtype, device, seed=77)
    res = torch.empty((n, 3), dtype=tdtype, device=device)
    out = {}
    for frac in (0.0, 0.01, 0.05, 0.2):
        k2 = kp.clone()
        if frac:
            gen = torch.Generator(device=device).manual_seed(3)
            bad = torch.rand((n, n_views), device=device, generator=gen) < frac
            k2[..., 2][bad] = 0.0
            k2[..., 0][bad] = 5000.0 * torch.rand((int(bad.sum()),), device=device, generator=gen).to(tdtype)
            k2[..., 1][bad] = 5000.0 * torch.rand((int(bad.sum()),), device=device, generator=gen).to(tdtype)
        ms = _time_launches(lambda: triangulate_multiview(k2, P, out=res), 10)
        out[f'{frac:.2f}'] = {'joints_per_s': n / ms * 1e3, 'roofline_frac': n * bytes_per_joint / ms / 1e6 / peak,
                              'finite_outputs': float(torch.isfinite(res).all(dim=1).float().mean())}
        del k2
    out['joints'] = n
    return out


def refine_sharded_vs_single(device, rank, world, n_frames=100_000, iters=60):
    """Config 4 sharded over the ranks against the SAME problem on one GPU (rank 0): cost history and final trajectory.
    Float state: the two runs add their 17 global sums in different orders, so they agree to rounding, not bit for bit."""
    import torch
    from mc3d_b200 import refinement as rf
    from mc3d_b200 import synthetic as syn
    gs, init, cams, _ = syn.refinement_inputs(n_frames, n_cams=2, seed=11)
    rows = rf.camera_rows(cams, list(cams))

    def run(comm):
        eng = rf.RefineEngine(init, gs, rows, syn.EXAMPLE_BODY_LENGTHS, torch_dtype=torch.float32, device=device, lr=0.01,
                              betas=(0.9, 0.999), lambda_smooth=1e-6, lambda_body_length=1.0, patience=10 ** 9, tolerance=1e-5,
                              max_iter=10 ** 9, ignore_distortions=False, window=(0, n_frames), n_window_frames=n_frames,
                              hist_capacity=iters + 8, comm=comm)
        eng.run(iters)
        torch.cuda.synchronize()
        hist = eng.history(iters)[:, 0].copy()
        traj = eng.trajectory()                       # gathers the shards (collective when sharded)
        traj = traj.cpu().numpy() if hasattr(traj, 'cpu') else np.asarray(traj)
        peer = eng.peer is not None
        eng.close()
        return hist, traj, peer

    h_sh, x_sh, peer = run(rf.DistComm())
    out = {'frames': n_frames, 'iters': iters, 'dtype': 'f32', 'world': world, 'in_kernel_exchange': peer}
    if rank == 0:
        h_1, x_1, _ = run(rf.LocalComm())
        hist_rel = float(np.max(np.abs(h_sh - h_1) / np.abs(h_1)))
        moved = float(np.max(np.abs(x_1 - np.asarray(init, dtype=np.float32))))
        traj_abs = float(np.max(np.abs(x_sh - x_1)))
        traj_rel = traj_abs / float(np.max(np.abs(x_1)))
        worst = max(hist_rel, traj_rel)
        out.update({'max_rel_diff_vs_single_gpu': worst, 'cost_history_rel_diff': hist_rel, 'trajectory_abs_diff_mm': traj_abs,
                    'trajectory_rel_diff': traj_rel, 'largest_move_mm': moved, 'tolerance': 1e-6, 'ok': bool(worst <= 1e-6)})
    flag = torch.tensor([1.0 if out.get('ok', True) else 0.0], device=device)
    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    out['ok'] = bool(flag.item() == 1.0)
    return out


def refine_benchmark(n_frames, iters, dtype_name, device, world=1, seed=0, peak=None):
    """Config 4 shape: refinement iterations per second on `n_frames` frames x 17 joints x 2 cameras
    (lambda_smooth=1e-6, lambda_body_length=1, lr=0.01).  Returns a dict for the JSON line."""
    import torch
    from mc3d_b200 import refinement as rf
    from mc3d_b200 import synthetic as syn
    gs, init, cams, _ = syn.refinement_inputs(n_frames, n_cams=2, seed=seed)
    dt = torch.float32 if dtype_name == 'f32' else torch.float64
    rows = rf.camera_rows(cams, list(cams))
    comm = rf.DistComm() if world > 1 else rf.LocalComm()
    eng = rf.RefineEngine(init, gs, rows, syn.EXAMPLE_BODY_LENGTHS, torch_dtype=dt, device=device, lr=0.01,
                          betas=(0.9, 0.999), lambda_smooth=1e-6, lambda_body_length=1.0, patience=10 ** 9, tolerance=1e-5,
                          max_iter=10 ** 9, ignore_distortions=False, window=(0, n_frames), n_window_frames=n_frames,
                          hist_capacity=iters * 2 + 64, comm=comm)
    warm = 32
    eng.run(warm)                                  # several ranks: also captures the 2-step CUDA graph
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    eng.run(iters)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    graphed = getattr(eng, '_graph', None) is not None
    exchange = 'none (one rank)' if world == 1 else ('in-kernel stores + sequence flags over NVLink peer memory (CUDA IPC), no NCCL'
                                                     if eng.peer is not None else 'NCCL all_reduce x2 + all_gather halo')
    hist = eng.history(warm + iters)
    plan = eng.plan()
    eng.close()
    esize = 4 if dtype_name == 'f32' else 8
    # per joint-frame and iteration read x, m, v, mu0, S (15 scalars) and write x, m, v (9), plus the gradient
    # written and read once when the step is split in passes (6): 30 scalars = 120 B in float.  The shipped two-phase step
    # stores THREE gradient components: pass 1 reads x, mu0, S (8) and writes 9, Adam reads 9 + m, v, x (9)
    # and writes m, v, x (9) = 44 scalars = 176 B (the best-trajectory snapshot is written when an improving streak ends).
    us = 1e3 * ms / iters
    algo = n_frames * 17 * 30 * esize
    moved = n_frames * 17 * 44 * esize
    per_gpu = algo / world
    out = {'frames': n_frames, 'iters': iters, 'dtype': dtype_name, 'iters_per_s': iters / (ms * 1e-3), 'us_per_iter': us,
           'algorithmic_bytes_per_iter': algo, 'moved_bytes_per_iter_two_phase': moved,
           'cost_first': float(hist[0, 0]), 'cost_last': float(hist[-1, 0]), 'step': plan, 'world': world,
           'exchange': exchange, 'nccl_collectives_per_iter': 3 if exchange.startswith('NCCL') else 0,
           'multi_rank_cuda_graph': graphed}
    if peak:
        gbs = per_gpu / (us * 1e-6) / 1e9
        out['roofline'] = {'bound': 'hbm', 'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak,
                           'algorithmic_bytes_per_joint_frame': 30 * esize, 'moved_bytes_per_joint_frame': 44 * esize,
                           'achieved_moved': gbs * 44 / 30, 'frac_moved': gbs * 44 / 30 / peak,
                           'per': 'GPU', 'kernel': 'mc3d::refine_fused2_kernel (float state: fused sweep, Adam of step s beside pass 1 of '
                                                   'step s + 1; double state: pass 1, then clip + Adam)',
                           'note': 'a shard of 12 500 .. 25 000 frames (25 .. 50 MB of state) is L2-resident and the step is bound by '
                                   'latency (one grid-wide meeting with the cross-rank exchange, the per-thread chain of pass 1), not by HBM'}
    return out


def _time_launches(fn, reps):
    import torch
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _time_max(fn, reps, device, world):
    """Average ms per call of fn over `reps` calls (CUDA events), max over the ranks."""
    import torch
    fn()
    _sync_all(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return _max_over_ranks(e0.elapsed_time(e1) / reps, device, world)


def _roof(bytes_per_gpu, ms, peak, **more):
    gbs = bytes_per_gpu / ms / 1e6
    return dict({'bound': 'hbm', 'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak, 'per': 'GPU'}, **more)


def extra_measurements(device, peak, rank=0, world=1):
    """The other BASELINE.json configs (inputs resident, CUDA events, every rank on its shard, times = max over ranks):
    config 2 in float64 at its full 10 M frames per GPU, config 5 (COCO-WholeBody 133 joints x 16 views: a total of 1e6 .. 1e9
    points sharded over the ranks) and config 3 (heatmap decode 17x64x48 x 4 views feeding the triangulation)."""
    import torch
    from mc3d_b200.decode import decode_heatmaps
    from mc3d_b200.triangulation import triangulate_multiview
    out = {}
    # config 2, float64 storage, 10 M frames per GPU (weak scaling like the headline)
    n = 10_000_000 * JOINTS
    kp, P = make_triangulation_workload(n, 8, torch.float64, device, seed=5 + rank)
    res = torch.empty((n, 3), dtype=torch.float64, device=device)
    ms = _time_max(lambda: triangulate_multiview(kp, P, out=res), 5, device, world)
    out['tri8_coco17_10Mframes_f64'] = {'joints_per_gpu': n, 'joints_per_s': world * n / ms * 1e3, 'ms_per_launch': ms,
                                        'roofline': _roof(n * 216, ms, peak, algorithmic_bytes_per_joint=216,
                                                          kernel='mc3d::triangulate_lean64_kernel<8, layout>'), 'n_gpus': world}
    del kp, res
    torch.cuda.empty_cache()
    # config 5: 16 views, WholeBody; a TOTAL of `total` points, sharded over the ranks (strong scaling)
    sweep = {}
    for total in (1_000_000, 10_000_000, 100_000_000, 1_000_000_000):
        for io, dt, es in (('f32', torch.float32, 4), ('f64', torch.float64, 8)):
            shard = total // world
            key = f'{total:.0e}_{io}'
            if shard * 54 * es > 150e9:
                sweep[key] = {'skipped': f'{shard * 54 * es / 1e9:.0f} GB per GPU does not fit in 180 GB of HBM beside the workspace: '
                                         'needs more ranks'}
                continue
            kp, P = make_triangulation_workload(shard, 16, dt, device, seed=6 + rank)
            res = torch.empty((shard, 3), dtype=dt, device=device)
            ms = _time_max(lambda: triangulate_multiview(kp, P, out=res), 3, device, world)
            sweep[key] = {'points_per_s': world * shard / ms * 1e3, 'points_per_gpu': shard, 'ms_per_launch': ms,
                          'roofline': _roof(shard * 51 * es, ms, peak, algorithmic_bytes_per_point=51 * es)}
            del kp, res
            torch.cuda.empty_cache()
    out['tri16_wholebody133'] = {'scaling': 'strong (total points fixed, sharded over the ranks, no collective)', 'n_gpus': world, 'sweep': sweep,
                                 'kernels': 'mc3d::triangulate_mixed_kernel<16, layout> (f32), mc3d::triangulate_kernel<double, 16, weighted> (f64)'}
    # config 3: decode (17 x 64 x 48 per view, 4 views) + triangulation; 16 384 frames resident per GPU, re-used as a stream
    T_, C = 16_384, 4
    hm = torch.rand((T_, C, JOINTS, 64, 48), device=device) * 0.02
    cy = torch.randint(8, 56, (T_, C, JOINTS), device=device)
    cx = torch.randint(8, 40, (T_, C, JOINTS), device=device)
    ti, ci, ji = torch.meshgrid(torch.arange(T_, device=device), torch.arange(C, device=device),
                                torch.arange(JOINTS, device=device), indexing='ij')
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            hm[ti, ci, ji, cy + dy, cx + dx] += 0.9 if (dx == 0 and dy == 0) else 0.4
    from mc3d_b200 import synthetic as syn
    P4 = syn.projection_matrices(syn.ring_rig(4))
    aff = torch.tensor([[1280 / 48.0, 720 / 64.0, 0.0, 0.0]], dtype=torch.float32, device=device).repeat(T_ * C, 1)
    res = torch.empty((T_, JOINTS, 3), dtype=torch.float32, device=device)

    def step():
        kpts, _ = decode_heatmaps(hm, want_moments=False, kpt_layout='nv3', affine=aff, affine_group=JOINTS)
        triangulate_multiview(kpts, P4, out=res)

    ms = _time_max(step, 5, device, world)
    in_bytes = hm.numel() * 4
    both = _time_max(lambda: decode_heatmaps(hm, kpt_layout='nv3', affine=aff, affine_group=JOINTS), 5, device, world)
    bytes_per_frame = 835_584 + 204
    out['decode_tri4_coco17'] = {'frames_resident_per_gpu': T_, 'frames_per_s': world * T_ / ms * 1e3,
                                 'joints_per_s': world * T_ * JOINTS / ms * 1e3, 'n_gpus': world,
                                 'roofline': _roof(T_ * bytes_per_frame, ms, peak, algorithmic_bytes_per_frame=bytes_per_frame,
                                                   kernel='mc3d::decode_reg6448_kernel<false> + mc3d::triangulate_mixed_kernel<4, layout>'),
                                 'decode_kpts_plus_moments': {'frames_per_s': world * T_ / both * 1e3,
                                                              'roofline': _roof(T_ * (835_584 + 4 * 17 * (12 + 48)), both, peak,
                                                                                kernel='mc3d::decode_reg6448_kernel<true>')},
                                 'time_for_1M_frames_s': 1e6 / (world * T_ / ms * 1e3),
                                 'note': '1 M frames = 835.6 GB of heatmaps, sharded over the ranks and streamed through HBM in 16 384-frame '
                                         'chunks; the resident chunk (13.7 GB >> L2) is re-used for timing'}
    del hm, res
    torch.cuda.empty_cache()
    return out


def _mem_available_bytes():
    try:
        with open('/proc/meminfo') as fh:
            for ln in fh:
                if ln.startswith('MemAvailable:'):
                    return int(ln.split()[1]) * 1024
    except OSError:
        pass
    return 64 << 30


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-refine', action='store_true', help='skip the refinement iters/sec extra')
    ap.add_argument('--no-extras', action='store_true', help='skip the other BASELINE configs (f64, 16-view WholeBody, decode+triangulate)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', '29531', os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
