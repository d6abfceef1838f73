import os, sys
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import __graft_entry__ as g
g.build()
from bench import make_triangulation_workload
from mc3d_b200.triangulation import triangulate_multiview
from mc3d_b200 import _lib
dev="cuda:0"; n=17_000_000
kp, P = make_triangulation_workload(n, 8, torch.float32, dev, seed=1)
gen = torch.Generator(device=dev).manual_seed(3)
bad = torch.rand((n, 8), device=dev, generator=gen) < 0.2
kp[..., 2][bad] = 0.0
kp[..., 0][bad] = 5000.0 * torch.rand((int(bad.sum()),), device=dev, generator=gen)
kp[..., 1][bad] = 5000.0 * torch.rand((int(bad.sum()),), device=dev, generator=gen)
out = triangulate_multiview(kp, P)
ref = triangulate_multiview(kp, P, flags=_lib.TRI_FLAG_FP64)
ref64 = triangulate_multiview(kp.double(), P)          # double storage kernel on the same values
ok = torch.isfinite(out).all(1) & torch.isfinite(ref).all(1)
d = (out.double()-ref.double()).norm(dim=1); d[~ok]=0
d2 = (out.double()-ref64).norm(dim=1); d2[~ok]=0
d3 = (ref.double()-ref64).norm(dim=1); d3[~ok]=0
nv = (kp[...,2]>0).sum(1)
print('max vs FP64 flag', d.max().item(), 'max vs double-storage kernel', d2.max().item(), 'FP64flag vs double-storage', d3.max().item())
idx = torch.argsort(d, descending=True)[:12]
for i in idx.tolist():
    print(i, 'err', f'{d[i].item():.2e}', 'err vs dbl-storage', f'{d2[i].item():.2e}', 'flag-vs-dbl', f'{d3[i].item():.2e}', 'views', int(nv[i]), (kp[i,:,2]>0).int().tolist(), 'X', ref[i].tolist())
for k in range(2,9):
    m = (nv==k)&ok
    if m.any(): print('views',k,'count',int(m.sum()),'max err',f'{d[m].max().item():.2e}', 'p99.9', f'{torch.quantile(d[m][:1000000].float(),0.999).item():.2e}')
