"""Short fixed sequence of every hot-path kernel, for ncu (launch list and --set full captures).

    python profiles/profile_run.py            # plain run, prints CUDA-event times per kernel
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'triangulate_|decode_|refine_' ... python profiles/profile_run.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=3):
    import torch
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))


def main():
    import torch
    import __graft_entry__ as g
    g.build()
    from bench import make_triangulation_workload
    from mc3d_b200 import refinement as rf
    from mc3d_b200 import synthetic as syn
    from mc3d_b200.decode import decode_heatmaps
    from mc3d_b200.triangulation import triangulate_multiview
    dev = 'cuda:0'
    n = 17_000_000
    out = {}
    only = sys.argv[1] if len(sys.argv) > 1 else ''
    for io, dt, es in (('f32', torch.float32, 4), ('f64', torch.float64, 8)):
        for V in (8, 16):
            if only and only != f'tri_{io}_V{V}':
                continue
            nn = n if V == 8 else n // 2
            kp, P = make_triangulation_workload(nn, V, dt, dev, seed=1)
            res = torch.empty((nn, 3), dtype=dt, device=dev)
            ms = timed(lambda: triangulate_multiview(kp, P, out=res))
            gbs = nn * (3 * V + 3) * es / ms / 1e6
            out[f'triangulate_{io}_V{V}'] = (ms, nn / ms * 1e3, gbs)
            del kp, res
    if only == 'decode':
        hm = torch.rand((200_000, 64, 48), device=dev) * 0.05
        hm[:, 30:34, 20:24] += 0.8
        ms = timed(lambda: decode_heatmaps(hm))
        out['decode_64x48'] = (ms, hm.shape[0] / ms * 1e3, hm.numel() * 4 / ms / 1e6)
        ms = timed(lambda: decode_heatmaps(hm, want_moments=False))
        out['decode_64x48_kpts_only'] = (ms, hm.shape[0] / ms * 1e3, hm.numel() * 4 / ms / 1e6)
    if only == 'refine':
        gs, init, cams, _ = syn.refinement_inputs(100_000, n_cams=2, seed=0)
        rows = rf.camera_rows(cams, list(cams))
        eng = rf.RefineEngine(init, gs, rows, syn.EXAMPLE_BODY_LENGTHS, torch_dtype=torch.float32, device=dev, lr=0.01,
                              betas=(0.9, 0.999), lambda_smooth=1e-6, lambda_body_length=1.0, patience=10 ** 9, tolerance=1e-5,
                              max_iter=10 ** 9, ignore_distortions=False, window=(0, 100_000), n_window_frames=100_000,
                              hist_capacity=64)
        ms = timed(lambda: eng.run(16), reps=4) / 16
        out['refine_step_f32_T100k'] = (ms, 1e3 / ms, 100_000 * 17 * 120 / ms / 1e6)        # 120 B per joint-frame
        eng.close()
        eng = rf.RefineEngine(init[:12500], gs[:12500], rows, syn.EXAMPLE_BODY_LENGTHS, torch_dtype=torch.float32, device=dev, lr=0.01,
                              betas=(0.9, 0.999), lambda_smooth=1e-6, lambda_body_length=1.0, patience=10 ** 9, tolerance=1e-5,
                              max_iter=10 ** 9, ignore_distortions=False, window=(0, 12_500), n_window_frames=12_500, hist_capacity=64)
        ms = timed(lambda: eng.run(16), reps=4) / 16           # one persistent launch = 16 steps
        out['refine_step_f32_T12500'] = (ms, 1e3 / ms, 12_500 * 17 * 120 / ms / 1e6)
        eng.close()
    if only:
        for k, (ms, rate, gbs) in out.items():
            print(f'{k:28s} {ms:9.4f} ms  {rate:14.4g} units/s  {gbs:8.1f} GB/s algorithmic')
        return
    hm = torch.rand((200_000, 64, 48), device=dev) * 0.05
    hm[:, 30:34, 20:24] += 0.8
    ms = timed(lambda: decode_heatmaps(hm))
    out['decode_64x48'] = (ms, hm.shape[0] / ms * 1e3, hm.numel() * 4 / ms / 1e6)
    del hm
    gs, init, cams, _ = syn.refinement_inputs(100_000, n_cams=2, seed=0)
    rows = rf.camera_rows(cams, list(cams))
    for io, dt in (('f32', torch.float32), ('f64', torch.float64)):
        eng = rf.RefineEngine(init, gs, rows, syn.EXAMPLE_BODY_LENGTHS, torch_dtype=dt, device=dev, lr=0.01, betas=(0.9, 0.999),
                              lambda_smooth=1e-6, lambda_body_length=1.0, patience=10 ** 9, tolerance=1e-5, max_iter=10 ** 9,
                              ignore_distortions=False, window=(0, 100_000), n_window_frames=100_000, hist_capacity=64)
        ms = timed(lambda: eng.run(16), reps=4) / 16
        out[f'refine_step_{io}_T100k'] = (ms, 1e3 / ms, 0.0)
        eng.close()
        eng = rf.RefineEngine(init[:12500], gs[:12500], rows, syn.EXAMPLE_BODY_LENGTHS, torch_dtype=dt, device=dev, lr=0.01,
                              betas=(0.9, 0.999), lambda_smooth=1e-6, lambda_body_length=1.0, patience=10 ** 9, tolerance=1e-5,
                              max_iter=10 ** 9, ignore_distortions=False, window=(0, 12_500), n_window_frames=12_500, hist_capacity=64)
        ms = timed(lambda: eng.run(16), reps=4) / 16           # one persistent launch = 16 steps
        out[f'refine_step_{io}_T12500'] = (ms, 1e3 / ms, 0.0)
        eng.close()
    for k, (ms, rate, gbs) in out.items():
        print(f'{k:28s} {ms:9.4f} ms  {rate:14.4g} units/s  {gbs:8.1f} GB/s algorithmic')


if __name__ == '__main__':
    main()
