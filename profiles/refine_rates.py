import sys, os, json
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g; g.build()
import bench
for n in (400, 12500, 100000):
    r = bench.refine_benchmark(n, 400 if n > 1000 else 2000, 'f32', 'cuda:0')
    print(os.environ.get('MC3D_REFINE_PEER',''), os.environ.get('MC3D_REFINE_FUSED',''), n, round(r['iters_per_s']), round(r['us_per_iter'],1), r['cost_first'], r['cost_last'])
