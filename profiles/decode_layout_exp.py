import sys, os
sys.path.insert(0, os.getcwd())
import torch
import __graft_entry__ as g
g.build()
from mc3d_b200.decode import decode_heatmaps
dev='cuda:0'
T_, C, J = 16384, 4, 17
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
hm = torch.rand((T_, C, J, 64, 48), device=dev) * 0.02
cy = torch.randint(8, 56, (T_, C, J), device=dev); cx = torch.randint(8, 40, (T_, C, J), device=dev)
ti, ci, ji = torch.meshgrid(torch.arange(T_, device=dev), torch.arange(C, device=dev), torch.arange(J, device=dev), indexing='ij')
for dy in (-1,0,1):
    for dx in (-1,0,1):
        hm[ti, ci, ji, cy+dy, cx+dx] += 0.9 if (dx==0 and dy==0) else 0.4
aff = torch.tensor([[1280/48.0, 720/64.0, 0.0, 0.0]], dtype=torch.float32, device=dev).repeat(T_*C, 1)
nb = hm.numel()*4
for name, fn in [('nv3+affine', lambda: decode_heatmaps(hm, kpt_layout='nv3', affine=aff, affine_group=J)),
                 ('nv3', lambda: decode_heatmaps(hm, kpt_layout='nv3')),
                 ('plain', lambda: decode_heatmaps(hm.view(-1,64,48))),
                 ('plain moments only', lambda: decode_heatmaps(hm.view(-1,64,48), want_kpts=False)),
                 ('plain kpts only', lambda: decode_heatmaps(hm.view(-1,64,48), want_moments=False))]:
    ms = timeit(fn); print(f'{name:22s} {ms:8.3f} ms {nb/ms/1e6:8.1f} GB/s {nb/ms/1e6/6451.8:.3f}')
hm2 = torch.rand((T_*C*J, 64, 48), device=dev) * 0.05
hm2[:, 30:34, 20:24] += 0.8
ms = timeit(lambda: decode_heatmaps(hm2)); print(f'ab-style data          {ms:8.3f} ms {nb/ms/1e6:8.1f} GB/s {nb/ms/1e6/6451.8:.3f}')
hm3 = hm2[:200000].contiguous()
ms = timeit(lambda: decode_heatmaps(hm3)); print(f'ab-style data 200k     {ms:8.3f} ms {hm3.numel()*4/ms/1e6:8.1f} GB/s {hm3.numel()*4/ms/1e6/6451.8:.3f}')
