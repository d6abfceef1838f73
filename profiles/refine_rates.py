"""Refinement iterations per second on one GPU for a few shard sizes (CUDA events); the environment selects the path:
MC3D_REFINE_PEER=0 plain graph of three kernels; MC3D_REFINE_TWO_PHASE=0/1, MC3D_REFINE_FUSED=0/1 force a variant,
MC3D_REFINE_SWEEP=0 the two-pass persistent kernel, MC3D_REFINE_BLOCKS=2/3 its register build; MC3D_RATES_DTYPE=f64."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402
g.build()
import bench  # noqa: E402
sizes = [int(a) for a in sys.argv[1:]] or [400, 12500, 100000]
for n in sizes:
    r = bench.refine_benchmark(n, 400 if n > 1000 else 2000, os.environ.get('MC3D_RATES_DTYPE', 'f32'), 'cuda:0')
    print('peer', os.environ.get('MC3D_REFINE_PEER', ''), 'fused', os.environ.get('MC3D_REFINE_FUSED', ''), 'two_phase',
          os.environ.get('MC3D_REFINE_TWO_PHASE', ''), n, round(r['iters_per_s']), round(r['us_per_iter'], 1), r['cost_first'], r['cost_last'])
