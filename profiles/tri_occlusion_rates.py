"""Float triangulation throughput when some views are unusable (weight 0, wild pixel), V = 8: the starting point of
solve_merged comes from views {0, 2, 5}; a bad one among them costs the warp another pass."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402
g.build()
import torch  # noqa: E402
from bench import make_triangulation_workload  # noqa: E402
from mc3d_b200.triangulation import triangulate_multiview  # noqa: E402
from profile_run import timed  # noqa: E402

n = 17_000_000
kp, P = make_triangulation_workload(n, 8, torch.float32, 'cuda:0', seed=1)
out = torch.empty((n, 3), dtype=torch.float32, device='cuda:0')
for frac in (0.0, 0.01, 0.05, 0.2):
    k2 = kp.clone()
    if frac:
        gen = torch.Generator(device='cuda:0').manual_seed(3)
        bad = torch.rand((n, 8), device='cuda:0', generator=gen) < frac
        k2[..., 2][bad] = 0.0
        k2[..., 0][bad] = 5000.0 * torch.rand((int(bad.sum()),), device='cuda:0', generator=gen)
        k2[..., 1][bad] = 5000.0 * torch.rand((int(bad.sum()),), device='cuda:0', generator=gen)
    ms = timed(lambda: triangulate_multiview(k2, P, out=out))
    print(f'{frac:5.2f} of the views unusable: {n / ms * 1e3:.3e} joints/s, finite outputs {float(torch.isfinite(out).all(dim=1).float().mean()):.4f}')
