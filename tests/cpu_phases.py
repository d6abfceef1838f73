"""TEST INFRASTRUCTURE: a numpy statement of the three refinement phases that speaks the same control-block
protocol as csrc/refine.cu (include/mc3d.h), so the frame-sharded driver (refinement.RefineEngine with DistComm)
can be exercised over gloo on CPU tensors.  Built on the oracle's cost/gradient functions; never used by the product."""
import math

import numpy as np

from oracle import refine as R

CT_ACC, CT_STATE, CT_HIST = 0, 32, 64


class NumpyPhases:
    def __init__(self, cams, bones, ignore_distortions=False):
        self.cams, self.bones, self.ign = cams, bones, ignore_distortions
        self.engine = None          # set by the test after constructing the engine

    def _views(self):
        e = self.engine
        return (e.x_ext.numpy(), e.m.numpy(), e.v.numpy(), e.best.numpy(), e.g.numpy(), e.mu0[0].numpy(), e.S[0].numpy(),
                e.term_ok.numpy(), e.ctrl.numpy())

    def flags(self, pb, stream):
        """mc3d_refine_flags_*: static validity of the smoothness term ending at local frame s, s in [0, n + 2)."""
        x_ext, term_ok = self.engine_views()
        n, off, total = int(pb.n_frames), int(pb.frame_offset), int(pb.total_frames)
        term_ok[:] = 0
        for s_ in range(n + 2):
            g0 = s_ + off
            if g0 - 2 < 0 or g0 >= total:
                continue
            term_ok[s_ + 2] = int(np.isfinite(x_ext[s_:s_ + 3]).all())      # ext frames s, s+1, s+2 = local s-2 .. s

    def engine_views(self):
        e = self.engine
        return e.x_ext.numpy(), e.term_ok.numpy()

    def _derive(self, pb, ctrl, p):
        acc = ctrl[CT_ACC + 16 * p:CT_ACC + 16 * p + 8]
        with np.errstate(divide='ignore', invalid='ignore'):
            mu = acc[4] / acc[5]
            lik = acc[0] / acc[1]
            cs = pb.lambda_smooth * acc[2] / acc[3] if pb.lambda_smooth > 0 else 0.0
            cb = pb.lambda_body * (acc[6] - 2 * mu * acc[4] + mu * mu * acc[5]) / pb.aa if pb.lambda_body > 0 else 0.0
        return dict(n_lik=acc[1], n_s=acc[3], mu=mu, lik=lik, cs=cs, cb=cb, total=lik + cs + cb)

    def phase(self, pb, phase, step, end_of_iteration, stream):
        x_ext, m, v, best, g, mu0, S, term_ok, ctrl = self._views()
        p = step & 1
        if ctrl[CT_STATE + 16 * p + 5] != 0.0:
            if phase == 2:
                ctrl[CT_STATE + 16 * (p ^ 1):CT_STATE + 16 * (p ^ 1) + 16] = ctrl[CT_STATE + 16 * p:CT_STATE + 16 * p + 16]
                ctrl[CT_ACC + 16 * (p ^ 1):CT_ACC + 16 * (p ^ 1) + 8] = 0.0
            return
        n, J = int(pb.n_frames), int(pb.n_joints)
        off = int(pb.frame_offset)
        wb, we = int(pb.win_begin), int(pb.win_end)
        glob = np.arange(n) + off
        inw = (glob >= wb) & (glob < we)
        Sfull = np.zeros((n, J, 2, 2))
        Sfull[..., 0, 0], Sfull[..., 0, 1], Sfull[..., 1, 0], Sfull[..., 1, 1] = S[..., 0], S[..., 1], S[..., 1], S[..., 2]
        x = x_ext[2:n + 2].astype(np.float64)
        xe = x_ext.astype(np.float64)
        acc = ctrl[CT_ACC + 16 * p:CT_ACC + 16 * p + 8]
        if phase == 0:
            idx = np.where(inw)[0]
            for cam in self.cams:
                pix = R.project(x[idx], cam, self.ign)
                d = pix - mu0[idx]
                q = 0.5 * np.einsum('tja,tjab,tjb->tj', d, Sfull[idx], d)
                ok = np.isfinite(q)
                acc[0] += q[ok].sum()
                acc[1] += ok.sum()
            if pb.lambda_smooth > 0:
                for t in idx:
                    if glob[t] - 2 >= wb and term_ok[t + 2]:
                        D = xe[t + 2] - 2 * xe[t + 1] + xe[t]
                        acc[2] += (D * D).sum()
                        acc[3] += 1
            if pb.lambda_body > 0:
                for s, e, a in self.bones:
                    b = np.linalg.norm(x[idx, e] - x[idx, s], axis=1)
                    ok = np.isfinite(b)
                    acc[4] += (a * b[ok]).sum()
                    acc[5] += (b[ok] ** 2).sum()
                    acc[6] += a * a * ok.sum()
        elif phase == 1:
            dv = self._derive(pb, ctrl, p)
            g[:] = 0
            idx = np.where(inw)[0]
            gg = np.zeros((n, J, 3))
            for cam in self.cams:
                pix, Jm = R.project(x[idx], cam, self.ign, jac=True)
                d = pix - mu0[idx]
                Sd = np.einsum('tjab,tjb->tja', Sfull[idx], d)
                q = 0.5 * np.einsum('tja,tja->tj', d, Sd)
                gi = np.einsum('tjak,tja->tjk', Jm, Sd) / dv['n_lik']
                gi[~np.isfinite(q)] = 0
                gg[idx] += np.nan_to_num(gi)
            if pb.lambda_smooth > 0:
                sc = 2 * pb.lambda_smooth / dv['n_s']
                for t in idx:
                    for k, coef in ((0, 1.0), (1, -2.0), (2, 1.0)):
                        if term_ok[t + 2 + k] and glob[t] + k - 2 >= wb and glob[t] + k < we:
                            tt = t + k                       # term ending at local frame tt
                            D = xe[tt + 2] - 2 * xe[tt + 1] + xe[tt]
                            gg[t] += sc * coef * D
            if pb.lambda_body > 0:
                for s, e, a in self.bones:
                    vec = x[idx, e] - x[idx, s]
                    b = np.linalg.norm(vec, axis=1)
                    ok = np.isfinite(b) & (b > 0)
                    coef = np.where(ok, -2 * pb.lambda_body * dv['mu'] * (a - dv['mu'] * b) / pb.aa / np.where(ok, b, 1), 0.0)
                    gg[idx, e] += coef[:, None] * np.nan_to_num(vec)
                    gg[idx, s] -= coef[:, None] * np.nan_to_num(vec)
            gg[~np.isfinite(x).all(axis=2)] = 0
            g[:] = gg.astype(g.dtype)
            acc[7] += (g.astype(np.float64) ** 2).sum()
        else:
            dv = self._derive(pb, ctrl, p)
            st = ctrl[CT_STATE + 16 * p:CT_STATE + 16 * p + 8]
            clip = min(1.0, 1.0 / (math.sqrt(acc[7]) + 1e-6))
            step_n = st[0] + 1
            run_sum, run_cnt = st[1] + dv['total'], st[2] + 1
            bestc, no_imp, iters = st[3], st[4], st[6]
            improved = stop = False
            if end_of_iteration:
                mean = run_sum / run_cnt
                run_sum += mean
                run_cnt += 1
                improved = mean < bestc - pb.tolerance
                if improved:
                    bestc, no_imp = mean, 0
                else:
                    no_imp += 1
                iters += 1
                stop = no_imp >= pb.patience or iters > pb.max_iter
            gi = np.where(inw[:, None, None], g.astype(np.float64), 0.0) * clip
            dt = m.dtype
            mn = (m + (gi.astype(dt) - m) * dt.type(1 - pb.beta1)).astype(dt)
            vn = (v * dt.type(pb.beta2) + dt.type(1 - pb.beta2) * gi.astype(dt) ** 2).astype(dt)
            bc1, bc2 = 1 - pb.beta1 ** step_n, 1 - pb.beta2 ** step_n
            denom = np.sqrt(vn) * dt.type(1 / math.sqrt(bc2)) + dt.type(pb.eps)
            xn = (x_ext[2:n + 2] - dt.type(pb.lr / bc1) * (mn / denom)).astype(dt)
            m[:], v[:] = mn, vn
            x_ext[2:n + 2] = xn
            if improved:
                best[:] = xn
            nx = ctrl[CT_STATE + 16 * (p ^ 1):CT_STATE + 16 * (p ^ 1) + 8]
            nx[:] = [step_n, run_sum, run_cnt, bestc, no_imp, float(stop), iters, float(improved)]
            ctrl[CT_ACC + 16 * (p ^ 1):CT_ACC + 16 * (p ^ 1) + 8] = 0.0
            hs = int(step_n - 1)
            if hs < pb.hist_capacity:
                ctrl[CT_HIST + 4 * hs:CT_HIST + 4 * hs + 4] = [dv['total'], dv['lik'], dv['cs'], dv['cb']]
