"""One launch of the float V = 8 kernel on 1.7e7 joints with a given fraction of unusable views (weight 0, wild pixel), for
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum -k regex:triangulate_mixed:   python profiles/tri_occl_probe.py 0.01"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from bench import make_triangulation_workload
from mc3d_b200.triangulation import triangulate_multiview
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
dev = 'cuda:0'
n = 17_000_000
kp, P = make_triangulation_workload(n, 8, torch.float32, dev, seed=1)
if frac > 0:
    gen = torch.Generator(device=dev).manual_seed(3)
    bad = torch.rand((n, 8), device=dev, generator=gen) < frac
    kp[..., 2][bad] = 0.0
    kp[..., 0][bad] = 5000.0 * torch.rand((int(bad.sum()),), device=dev, generator=gen)
    kp[..., 1][bad] = 5000.0 * torch.rand((int(bad.sum()),), device=dev, generator=gen)
out = torch.empty((n, 3), dtype=torch.float32, device=dev)
triangulate_multiview(kp, P, out=out)
torch.cuda.synchronize()
print('finite', float(torch.isfinite(out).all(dim=1).float().mean()))
