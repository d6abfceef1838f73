"""Per-pass times of the two-phase refinement step at one shard size: the graph-of-two-kernels variant (MC3D_REFINE_FUSED=0)
under `ncu --metrics gpu__time_duration.sum`, so that pass 1 (refine_costgrad_kernel) and pass 2 (refine_step2_kernel) are
timed separately.  Usage: MC3D_REFINE_FUSED=0 python profiles/refine_passes.py [frames] [steps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402
g.build()
import torch  # noqa: E402
from mc3d_b200 import refinement as rf  # noqa: E402
from mc3d_b200 import synthetic as syn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
gs, init, cams, _ = syn.refinement_inputs(n, n_cams=2, seed=0)
rows = rf.camera_rows(cams, list(cams))
eng = rf.RefineEngine(init, gs, rows, syn.EXAMPLE_BODY_LENGTHS, torch_dtype=torch.float32, device='cuda:0', lr=0.01,
                      betas=(0.9, 0.999), lambda_smooth=1e-6, lambda_body_length=1.0, patience=10 ** 9, tolerance=1e-5,
                      max_iter=10 ** 9, ignore_distortions=False, window=(0, n), n_window_frames=n, hist_capacity=64)
eng.run(steps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.run(steps)
e1.record()
torch.cuda.synchronize()
print(f'{n} frames: {e0.elapsed_time(e1) / steps * 1e3:.1f} us per step ({os.environ.get("MC3D_REFINE_FUSED", "default")})')
eng.close()
