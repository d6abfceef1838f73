// Trajectory refinement for sm_100a: the loss, its hand-derived gradient and the clipped Adam step that the
// reference's Optimized_3d_Pose_Estimation.sgd_optimize iterates with torch autograd
// (pose_refinement.py:836-889 costs, :1002-1091 loop).
//
// Two formulations of one optimiser step over the frame-sharded state (no host round trip inside either):
//
// three-phase step (mc3d_refine_phase_*: windows of a batch, host-driven multi-GPU exchange, very large shards)
//   A  costs   : per (frame, joint) reprojection Mahalanobis terms for every camera (camera-0 Gaussians, upstream
//                quirk Q1), second-difference smoothness terms, bone lengths -> 7 global sums (finite-masked,
//                nan_mean semantics)
//   B  gradient: closed-form gradient of the three terms (needs the global sums of A: counts, mu = a.b/b.b)
//                -> g, and the global sum of g^2
//   C  step    : clip_grad_norm_(1.0), Adam (torch.optim.Adam arithmetic), running-mean early-stopping
//                bookkeeping, conditional best-trajectory snapshot, cost history
// two-phase step (mc3d_refine_run_*, the default; see "two-phase step" below)
//   1  costs + the four gradient components the gradient is linear in + their 10 dot products -> 17 global sums
//   2  every thread derives the mix coefficients and |g|^2 from the sums, then the same clip / Adam / bookkeeping
//   as a persistent cooperative kernel (all iterations in one launch, grid barriers) or a graph of two kernels.
//
// Global sums live in a small double control block (ping-pong by step parity); every thread re-derives the scalars it
// needs, so there are no single-thread "finalise" launches.  The passes are grid-stride loops with one (frame, joint)
// item per thread iteration and no block barriers inside: temporal neighbours and bone end points come from global
// memory through L1, the cameras sit in shared memory (128-bit loads), sums are reduced thread -> warp -> block ->
// one double atomic per block.  The arithmetic type is the state dtype (float state -> float maths, as upstream's
// float32 run).  Several ranks exchange their partial sums and boundary frames INSIDE these kernels over NVLink peer
// memory ("in-kernel exchange" below); a host-driven driver may instead all-reduce the control block between phases.
#include "mc3d_common.cuh"
#include <math.h>
#include <stdlib.h>
#include <type_traits>

namespace mc3d {

constexpr int RF_THREADS = 256;
#ifndef MC3D_RF_GRID
#define MC3D_RF_GRID 8
#endif
#ifndef MC3D_RF_SMALL
#define MC3D_RF_SMALL 34         // shards up to MC3D_RF_SMALL x SMs x 512 items (~150 000 frames x 17 joints) run the two-phase persistent kernel
#endif
// control block layout (doubles)
constexpr int CT_ACC = 0;        // + 16 * parity : S_lik N_lik S_s N_s ab bb aa_ok gnorm2
constexpr int CT_STATE = 32;     // + 16 * parity : step run_sum run_cnt best no_improve stopped iters_done improved
constexpr int CT_HIST = 64;      // + 4 * step    : total lik smooth body

struct RefineDerived {
    double inv_nlik, smooth_scale, mu, body_c;
    double cost_lik, cost_s, cost_b, total;
    bool stopped;
};

__device__ __forceinline__ RefineDerived derive(const mc3d_refine_problem &pb, const double *acc, const double *st) {
    RefineDerived d;
    d.stopped = st[5] != 0.0;
    const double nl = acc[1], ns = acc[3];
    d.inv_nlik = 1.0 / nl;
    d.cost_lik = acc[0] / nl;
    d.smooth_scale = 2.0 * pb.lambda_smooth / ns;
    d.cost_s = pb.lambda_smooth > 0.0 ? pb.lambda_smooth * acc[2] / ns : 0.0;
    d.mu = acc[4] / acc[5];
    d.body_c = -2.0 * pb.lambda_body * d.mu / pb.aa;
    d.cost_b = pb.lambda_body > 0.0 ? pb.lambda_body * (acc[6] - 2.0 * d.mu * acc[4] + d.mu * d.mu * acc[5]) / pb.aa : 0.0;
    d.total = d.cost_lik + d.cost_s + d.cost_b;
    return d;
}

// ---- in-kernel exchange over peer memory (include/mc3d.h: mc3d_refine_xchg) ---------------------------------------
// Push model: a rank stores its contribution into every peer's block and then a sequence flag (release, system
// scope); consumers spin on flags in their OWN memory (acquire, system scope).  Flags carry the Adam step count, so
// nothing is ever reset and a replayed CUDA graph needs no per-step argument.
__device__ __forceinline__ bool xchg_on(const mc3d_refine_problem &pb) { return pb.xchg[pb.rank] != nullptr; }
__device__ __forceinline__ mc3d_refine_xchg *xchg_of(const mc3d_refine_problem &pb, int r) {
    return reinterpret_cast<mc3d_refine_xchg *>(pb.xchg[r]);
}
__device__ __forceinline__ long long ld_relaxed_sys(const int64_t *p) {
    long long v;
    asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(int64_t *p, long long v) {
    asm volatile("st.relaxed.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// one rank only: nothing leaves this GPU, device scope is enough
__device__ __forceinline__ void fence_xchg(const mc3d_refine_problem &pb) { if (pb.world > 1) fence_sys(); else fence_gpu(); }
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Wait until *flag >= target with relaxed polls; the caller fences ONCE after its last wait (acquire pattern).
// A peer that never arrives must not hang the GPU: give up after the timeout, mark the local block as failed (the
// host checks it after the run) and carry on with whatever is there.
__device__ __noinline__ void xchg_wait(const mc3d_refine_problem &pb, const int64_t *flag, long long target) {
    if (ld_relaxed_sys(flag) >= target) return;
    const unsigned long long t0 = globaltimer_ns();
    const unsigned long long limit = pb.spin_timeout_ns > 0 ? (unsigned long long)pb.spin_timeout_ns : 10000000000ULL;
    while (ld_relaxed_sys(flag) < target) {
        if (globaltimer_ns() - t0 > limit) {
            xchg_of(pb, pb.rank)->error = 1;
            __threadfence_system();
            return;
        }
    }
}

// Last block of a phase: store N partial sums (acc[0..N)) into slot sums[parity][rank][OFF..] of EVERY rank, then the
// flag.  `ticket` counts finished blocks and is reset by the last one for the next launch.
template <int N, int OFF>
__device__ __forceinline__ void xchg_publish(const mc3d_refine_problem &pb, int parity, const double *acc, int ticket_idx,
                                             bool grad_flag, long long seq) {
    __shared__ int is_last;
    mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(&mine->ticket[ticket_idx]), 1ULL);
        is_last = old == (unsigned long long)gridDim.x - 1ULL;
    }
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x == 0) mine->ticket[ticket_idx] = 0;
    __threadfence();
    if (threadIdx.x < N) {
        const double v = __ldcg(acc + threadIdx.x);
        for (int r = 0; r < pb.world; ++r) xchg_of(pb, r)->sums[parity][pb.rank][OFF + threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_xchg(pb);                                            // release: the sums before the flags
        for (int r = 0; r < pb.world; ++r) {
            mc3d_refine_xchg *xr = xchg_of(pb, r);
            st_relaxed_sys(grad_flag ? &xr->seq_grad[parity][pb.rank] : &xr->seq_costs[parity][pb.rank], seq);
        }
    }
}

// Totals over ranks of sums[parity][*][0..N) into tot[] (shared), added in rank order on every rank.
template <int N>
__device__ __forceinline__ void xchg_gather(const mc3d_refine_problem &pb, int parity, bool grad_flag, long long seq, double *tot) {
    mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
    if (threadIdx.x == 0) {
        for (int r = 0; r < pb.world; ++r)
            xchg_wait(pb, grad_flag ? &mine->seq_grad[parity][r] : &mine->seq_costs[parity][r], seq);
        fence_xchg(pb);                                            // acquire: the flags before the sums
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double s = 0.0;
        for (int r = 0; r < pb.world; ++r) s += __ldcg(&mine->sums[parity][r][threadIdx.x]);
        tot[threadIdx.x] = s;
    }
    __syncthreads();
}

// Arithmetic type of the per-element maths: the state dtype (float state -> float arithmetic, as upstream's torch
// float32 run; double state -> double).  Global sums are always accumulated in double.
template <typename C> struct Lim;
template <> struct Lim<float> { static __device__ __forceinline__ float big() { return 3.0e38f; } };
template <> struct Lim<double> { static __device__ __forceinline__ double big() { return 1.0e300; } };
template <typename C> __device__ __forceinline__ bool finite_c(C v) { return fabs(v) <= Lim<C>::big(); }
// Float state: the special-function unit's reciprocal / square root (one MUFU, <= 1-2 ulp) with one Newton step for the
// reciprocals of the projection -- no IEEE slow path (a subroutine call per division that also pins the schedule).  The
// float run is compared with upstream's float32 history at 1e-4; these differ from correctly rounded results in the last
// bit.  Double state: IEEE.
__device__ __forceinline__ float rcp_c(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return fmaf(r, fmaf(-v, r, 1.0f), r);
}
__device__ __forceinline__ double rcp_c(double v) { return 1.0 / v; }
__device__ __forceinline__ float sqrt_c(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ double sqrt_c(double v) { return sqrt(v); }
__device__ __forceinline__ float div_c(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double div_c(double a, double b) { return a / b; }

// Reprojection term of one camera: returns 0.5 d^T S d, and (when GRAD) adds J^T S d * scale to g[3].
constexpr int CAM_STRIDE = 28;      // K[9] R[9] T[3] dist[5] + 2 pad: a whole number of 16-byte shared-memory loads

// One camera from shared memory into registers with 128-bit loads (28 scalars = 7 / 14 loads for float / double).
template <typename C>
__device__ __forceinline__ void load_camera(const C *cam_smem, C (&cam)[CAM_STRIDE]) {
    constexpr int PER = 16 / sizeof(C);
    struct alignas(16) V { C v[PER]; };
#pragma unroll
    for (int i = 0; i < CAM_STRIDE / PER; ++i) {
        const V q = reinterpret_cast<const V *>(cam_smem)[i];
#pragma unroll
        for (int k = 0; k < PER; ++k) cam[i * PER + k] = q.v[k];
    }
}

// cam: K[9] R[9] T[3] dist[5] in the arithmetic type, in registers.
// affine_k: the last row of every camera's K is (0, 0, 1) -- the pixel is K applied to the distorted point with no second
// division.  Same values as the general form for finite input (0 x + 0 y + 1 = 1, 1 / 1 = 1, K0 - px 0 = K0), ~12 fewer
// instructions per camera.
template <bool GRAD, typename C>
__device__ __forceinline__ C reproject_term(const C (&cam)[CAM_STRIDE], bool ignore_dist, C X, C Y, C Z, C mx, C my, C s00,
                                            C s01, C s11, C scale, C *g, bool affine_k = false) {
    const C *K = cam, *R = cam + 9, *T = cam + 18, *D = cam + 21;
    const C one = (C)1, two = (C)2;
    const C xc = fma(R[0], X, fma(R[1], Y, fma(R[2], Z, T[0])));
    const C yc = fma(R[3], X, fma(R[4], Y, fma(R[5], Z, T[1])));
    const C zc = fma(R[6], X, fma(R[7], Y, fma(R[8], Z, T[2])));
    const C iz = rcp_c(zc);
    const C a = xc * iz, b = yc * iz;
    C xd = a, yd = b, j00 = one, j01 = (C)0, j11 = one;
    if (!ignore_dist) {
        const C k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4];
        const C r2 = fma(a, a, b * b);
        const C rad = fma(fma(fma(k3, r2, k2), r2, k1), r2, one);
        xd = fma(a, rad, fma(two * p1 * a, b, p2 * fma(two * a, a, r2)));
        yd = fma(b, rad, fma(p1, fma(two * b, b, r2), two * p2 * a * b));
        if (GRAD) {
            const C drad = fma(fma((C)3 * k3, r2, two * k2), r2, k1);
            j00 = rad + two * a * a * drad + two * p1 * b + (C)6 * p2 * a;
            j01 = two * a * b * drad + two * p1 * a + two * p2 * b;
            j11 = rad + two * b * b * drad + (C)6 * p1 * b + two * p2 * a;
        }
    }
    const C u = fma(K[0], xd, fma(K[1], yd, K[2]));
    const C v = fma(K[3], xd, fma(K[4], yd, K[5]));
    C is = one, px = u, py = v;
    if (!affine_k) {
        const C sden = fma(K[6], xd, fma(K[7], yd, K[8]));
        is = rcp_c(sden);
        px = u * is; py = v * is;
    }
    const C dx = px - mx, dy = py - my;
    const C sdx = fma(s00, dx, s01 * dy), sdy = fma(s01, dx, s11 * dy);
    const C q = (C)0.5 * fma(dx, sdx, dy * sdy);
    if (GRAD && finite_c(q)) {
        // pixel -> distorted normalised
        C gxd, gyd;
        if (affine_k) {
            gxd = sdx * K[0] + sdy * K[3];
            gyd = sdx * K[1] + sdy * K[4];
        } else {
            gxd = (sdx * (K[0] - px * K[6]) + sdy * (K[3] - py * K[6])) * is;
            gyd = (sdx * (K[1] - px * K[7]) + sdy * (K[4] - py * K[7])) * is;
        }
        // distorted -> normalised
        const C ga = fma(j00, gxd, j01 * gyd), gb = fma(j01, gxd, j11 * gyd);
        // normalised -> camera frame
        const C gxc = ga * iz, gyc = gb * iz, gzc = -(a * ga + b * gb) * iz;
        // camera -> world (R^T)
        g[0] = fma(scale, fma(R[0], gxc, fma(R[3], gyc, R[6] * gzc)), g[0]);
        g[1] = fma(scale, fma(R[1], gxc, fma(R[4], gyc, R[7] * gzc)), g[1]);
        g[2] = fma(scale, fma(R[2], gxc, fma(R[5], gyc, R[8] * gzc)), g[2]);
    }
    return q;
}

// Bone / adjacency tables are indexed per thread: keep them in shared memory (constant-bank reads with
// lane-divergent addresses serialise).
struct RefineTables {
    double bone_len[MC3D_MAX_BONES];
    int bone_start[MC3D_MAX_BONES], bone_end[MC3D_MAX_BONES];
    int adj_start[MC3D_MAX_JOINTS + 3];
    int adj_bone[2 * MC3D_MAX_BONES], adj_sign[2 * MC3D_MAX_BONES];
    // the same adjacency, resolved for the two-phase pass 1: per (joint, incident bone) the OTHER end point's offset inside a
    // frame (scalars), the target length, and whether this joint owns the bone's three cost sums (it is the bone's end)
    int2 adj_other_owner[2 * MC3D_MAX_BONES];
    double adj_len[2 * MC3D_MAX_BONES];
};

__device__ __forceinline__ void load_tables(RefineTables &tb, const mc3d_refine_problem &pb) {
    for (int i = threadIdx.x; i < MC3D_MAX_BONES; i += blockDim.x) {
        tb.bone_len[i] = pb.bone_len[i]; tb.bone_start[i] = pb.bone_start[i]; tb.bone_end[i] = pb.bone_end[i];
    }
    for (int i = threadIdx.x; i < 2 * MC3D_MAX_BONES; i += blockDim.x) {
        const int k = pb.adj_bone[i], sg = pb.adj_sign[i];
        tb.adj_bone[i] = k; tb.adj_sign[i] = sg;
        const bool valid = k >= 0 && k < MC3D_MAX_BONES;
        const int other = valid ? (sg > 0 ? pb.bone_start[k] : pb.bone_end[k]) : 0;
        tb.adj_other_owner[i] = make_int2(other * 3, sg > 0 ? 1 : 0);
        tb.adj_len[i] = valid ? pb.bone_len[k] : 0.0;
    }
    for (int i = threadIdx.x; i <= pb.n_joints; i += blockDim.x) tb.adj_start[i] = pb.adj_start[i];
}

template <int N>
__device__ __forceinline__ void block_reduce_add(double (&vals)[N], double *smem_red, double *global_acc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) vals[i] = warp_sum(vals[i]);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < N; ++i) smem_red[warp * N + i] = vals[i];
    __syncthreads();
    if (threadIdx.x < N) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += smem_red[w * N + threadIdx.x];
        if (s != 0.0) atomicAdd(global_acc + threadIdx.x, s);
    }
    __syncthreads();
}

// ---- smoothness-term validity flags --------------------------------------------------------------------------------
// term_ok[s + 2] = 1 when frames s, s-1, s-2 are entirely finite, for local s in [0, n + 2) (the two extra entries are
// the terms owned by the right neighbour, which the gradient of the last two local frames needs; their frames are in
// the halo).  upstream drops a whole frame's term when it is non-finite (nan_mean, pose_refinement.py:845).  A joint
// that starts non-finite stays frozen (its gradient is masked) and a finite one only becomes non-finite if the
// optimisation has already diverged, so the flags are computed once per run, not once per step.
template <typename T>
__global__ void refine_flags_kernel(const __grid_constant__ mc3d_refine_problem pb) {
    const long long n = pb.n_frames;
    const int per = pb.n_joints * 3;
    const T *x_ext = (const T *)pb.x;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n + 2; s += (long long)gridDim.x * blockDim.x) {
        bool ok = true;
        const long long g0 = s + pb.frame_offset;                  // global index of the term's last frame
        if (g0 - 2 < 0 || g0 >= pb.total_frames) ok = false;       // frames outside the trajectory
        for (int f = 0; f < 3 && ok; ++f) {
            const T *fr = x_ext + (s + 2 - f) * per;               // ext index = local + 2
            for (int i = 0; i < per; ++i) ok = ok && finite_c(fr[i]);
        }
        pb.term_ok[s + 2] = ok ? 1 : 0;
    }
}

// ---- kernel A: costs ------------------------------------------------------------------------------------------
// One thread per (frame, joint) item, grid-stride, no shared-memory staging and no block barriers in the loop:
// temporal neighbours and bone end points are read straight from global memory (each x element is read by the five
// items around it and by its bones, all within a few hundred bytes, so they hit L1).  Per-thread partial sums are
// float (a thread sees a few dozen items) and are widened to double for the warp / block / grid reduction.
struct ItemCursor {                 // (frame, joint) of a grid-stride loop without a division per item
    long long t, e;
    int j;
    long long de, dt;
    int dj;
    __device__ __forceinline__ ItemCursor(int J) {
        e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        de = (long long)gridDim.x * blockDim.x;
        t = e / J; j = (int)(e - t * J);
        dt = de / J; dj = (int)(de - dt * J);
    }
    __device__ __forceinline__ void next(int J) {
        e += de; t += dt; j += dj;
        if (j >= J) { j -= J; ++t; }
    }
};

// The grid-stride cost loop of one rank: acc[0..7) = this thread's partial sums (widened to double).
template <typename T>
__device__ __forceinline__ void costs_loop(const mc3d_refine_problem &pb, const RefineTables &tb, const T *camf, double (&acc)[7]) {
    const int J = pb.n_joints, C = pb.n_cams, NB = pb.n_bones, JS = J * 3;
    const T *x = (const T *)pb.x + 2LL * JS;                       // local frame 0 (halo frames sit before / after)
    const T *mu0 = (const T *)pb.mu0, *S = (const T *)pb.S;
    const long long gstride = pb.gauss_cam_stride;                 // 0: camera-0 Gaussians for every camera (upstream, Q1)
    const long long n_items = pb.n_frames * J;
    const bool ign = pb.ignore_distortions != 0;
    const bool do_smooth = pb.lambda_smooth > 0.0, do_body = pb.lambda_body > 0.0;
    const long long lo = pb.win_begin - pb.frame_offset, hi = pb.win_end - pb.frame_offset;   // window in local frames
    T accf[7] = {(T)0, (T)0, (T)0, (T)0, (T)0, (T)0, (T)0};
    for (ItemCursor it(J); it.e < n_items; it.next(J)) {
        const long long t = it.t, e = it.e;
        const int j = it.j;
        if (t < lo || t >= hi) continue;
        const T *xc = x + e * 3;
        const T X = xc[0], Y = xc[1], Z = xc[2];
        T mx = mu0[e * 2], my = mu0[e * 2 + 1];
        T s00 = S[e * 3], s01 = S[e * 3 + 1], s11 = S[e * 3 + 2];
        for (int c = 0; c < C; ++c) {
            T cam[CAM_STRIDE];
            load_camera(camf + c * CAM_STRIDE, cam);
            if (gstride && c > 0) {                                 // per-camera Gaussians (opt-in; upstream uses camera 0's, Q1)
                const long long ec = e + c * gstride;
                mx = mu0[ec * 2]; my = mu0[ec * 2 + 1];
                s00 = S[ec * 3]; s01 = S[ec * 3 + 1]; s11 = S[ec * 3 + 2];
            }
            const T q = reproject_term<false, T>(cam, ign, X, Y, Z, mx, my, s00, s01, s11, (T)0, nullptr);
            const bool ok = finite_c(q);
            accf[0] += ok ? q : (T)0;
            accf[1] += ok ? (T)1 : (T)0;
        }
        if (do_smooth && t - 2 >= lo && pb.term_ok[t + 2]) {
            const T *x1 = xc - JS, *x2 = xc - 2 * JS;
            const T a0 = X - (T)2 * x1[0] + x2[0], a1 = Y - (T)2 * x1[1] + x2[1], a2 = Z - (T)2 * x1[2] + x2[2];
            accf[2] += a0 * a0 + a1 * a1 + a2 * a2;
            accf[3] += j == 0 ? (T)1 : (T)0;
        }
        if (do_body) {
            const T *xf = x + t * JS;
            for (int k = j; k < NB; k += J) {
                const T *ps = xf + tb.bone_start[k] * 3, *pe = xf + tb.bone_end[k] * 3;
                const T v0 = pe[0] - ps[0], v1 = pe[1] - ps[1], v2 = pe[2] - ps[2];
                const T b = sqrt_c(v0 * v0 + v1 * v1 + v2 * v2);
                if (finite_c(b)) {
                    const T a = (T)tb.bone_len[k];
                    accf[4] += a * b; accf[5] += b * b; accf[6] += a * a;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) acc[i] = (double)accf[i];
}

template <typename T>
__device__ __forceinline__ void load_cameras_and_tables(const mc3d_refine_problem &pb, RefineTables &tb, T *camf) {
    load_tables(tb, pb);
    for (int i = threadIdx.x; i < pb.n_cams * CAM_STRIDE; i += blockDim.x)
        camf[i] = (i % CAM_STRIDE) >= 26 ? (T)0
                : pb.cams_dev ? (T)__ldcg(pb.cams_dev + (i / CAM_STRIDE) * 26 + i % CAM_STRIDE)      // learnt on the device
                              : (T)pb.cams[i / CAM_STRIDE][i % CAM_STRIDE];
}

template <typename T>
__global__ void __launch_bounds__(RF_THREADS)
refine_costs_kernel(const __grid_constant__ mc3d_refine_problem pb, int parity) {
    __shared__ double red[8 * 7];
    __shared__ RefineTables tb;
    __shared__ __align__(16) T camf[MC3D_MAX_VIEWS * CAM_STRIDE];
    double *ctrl = pb.ctrl;
    if (ctrl[CT_STATE + 16 * parity + 5] != 0.0) return;          // stopped
    const bool xchg = xchg_on(pb);
    const long long adam_step = (long long)ctrl[CT_STATE + 16 * parity];
    if (xchg && threadIdx.x == 0) {                                // the neighbours' boundary frames of this step are in my halo
        mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
        if (pb.rank > 0) xchg_wait(pb, &mine->halo_seq[0], adam_step);
        if (pb.rank < pb.world - 1) xchg_wait(pb, &mine->halo_seq[1], adam_step);
        fence_xchg(pb);
    }
    load_cameras_and_tables(pb, tb, camf);
    __syncthreads();
    double acc[7];
    costs_loop<T>(pb, tb, camf, acc);
    block_reduce_add<7>(acc, red, ctrl + CT_ACC + 16 * parity);
    if (xchg) xchg_publish<7, 0>(pb, parity, ctrl + CT_ACC + 16 * parity, 0, false, adam_step + 1);
}

// ---- kernel B: gradient ---------------------------------------------------------------------------------------
// The grid-stride gradient loop of one rank: writes g, returns this thread's partial sum of g^2.
template <typename T>
__device__ __forceinline__ double grad_loop(const mc3d_refine_problem &pb, const RefineTables &tb, const T *camf, const RefineDerived &dv) {
    const int J = pb.n_joints, C = pb.n_cams, JS = J * 3;
    const T *x = (const T *)pb.x + 2LL * JS;
    const T *mu0 = (const T *)pb.mu0, *S = (const T *)pb.S;
    const long long gstride = pb.gauss_cam_stride;                 // 0: camera-0 Gaussians for every camera (upstream, Q1)
    T *gout = (T *)pb.g;
    const long long n_items = pb.n_frames * J;
    const bool ign = pb.ignore_distortions != 0;
    const bool do_smooth = pb.lambda_smooth > 0.0, do_body = pb.lambda_body > 0.0;
    const long long lo = pb.win_begin - pb.frame_offset, hi = pb.win_end - pb.frame_offset;
    const T inv_nlik = (T)dv.inv_nlik, smooth_scale = (T)dv.smooth_scale, mu = (T)dv.mu, body_c = (T)dv.body_c;
    T gnf = (T)0;
    for (ItemCursor it(J); it.e < n_items; it.next(J)) {
        const long long t = it.t, e = it.e;
        const int j = it.j;
        T g[3] = {(T)0, (T)0, (T)0};
        const T *xc = x + e * 3;
        const T X = xc[0], Y = xc[1], Z = xc[2];
        const bool self_ok = finite_c(X) && finite_c(Y) && finite_c(Z);
        if (t >= lo && t < hi && self_ok) {
            T mx = mu0[e * 2], my = mu0[e * 2 + 1];
            T s00 = S[e * 3], s01 = S[e * 3 + 1], s11 = S[e * 3 + 2];
            for (int c = 0; c < C; ++c) {
                T cam[CAM_STRIDE];
                load_camera(camf + c * CAM_STRIDE, cam);
                if (gstride && c > 0) {
                    const long long ec = e + c * gstride;
                    mx = mu0[ec * 2]; my = mu0[ec * 2 + 1];
                    s00 = S[ec * 3]; s01 = S[ec * 3 + 1]; s11 = S[ec * 3 + 2];
                }
                reproject_term<true, T>(cam, ign, X, Y, Z, mx, my, s00, s01, s11, inv_nlik, g);
            }
            if (do_smooth) {
                // d/dx_t of sum_s ||D_s||^2 = 2 (D_t - 2 D_{t+1} + D_{t+2}) over the valid terms s (s - 2 >= lo, s < hi)
                const T two = (T)2;
                const bool k0 = t - 2 >= lo && pb.term_ok[t + 2];
                const bool k1 = t - 1 >= lo && t + 1 < hi && pb.term_ok[t + 3];
                const bool k2 = t + 2 < hi && pb.term_ok[t + 4];
                T s0 = (T)0, s1 = (T)0, s2 = (T)0;
                if (k0) {
                    s0 += xc[0] - two * xc[-JS] + xc[-2 * JS]; s1 += xc[1] - two * xc[1 - JS] + xc[1 - 2 * JS];
                    s2 += xc[2] - two * xc[2 - JS] + xc[2 - 2 * JS];
                }
                if (k1) {
                    s0 -= two * (xc[JS] - two * xc[0] + xc[-JS]); s1 -= two * (xc[1 + JS] - two * xc[1] + xc[1 - JS]);
                    s2 -= two * (xc[2 + JS] - two * xc[2] + xc[2 - JS]);
                }
                if (k2) {
                    s0 += xc[2 * JS] - two * xc[JS] + xc[0]; s1 += xc[1 + 2 * JS] - two * xc[1 + JS] + xc[1];
                    s2 += xc[2 + 2 * JS] - two * xc[2 + JS] + xc[2];
                }
                g[0] = fma(smooth_scale, s0, g[0]); g[1] = fma(smooth_scale, s1, g[1]); g[2] = fma(smooth_scale, s2, g[2]);
            }
            if (do_body) {
                const T *xf = x + t * JS;
                for (int q = tb.adj_start[j]; q < tb.adj_start[j + 1]; ++q) {
                    const int k = tb.adj_bone[q];
                    const T sign = (T)tb.adj_sign[q];
                    const T *ps = xf + tb.bone_start[k] * 3, *pe = xf + tb.bone_end[k] * 3;
                    const T v0 = pe[0] - ps[0], v1 = pe[1] - ps[1], v2 = pe[2] - ps[2];
                    const T b = sqrt_c(v0 * v0 + v1 * v1 + v2 * v2);
                    if (finite_c(b) && b > (T)0) {
                        const T coef = div_c(sign * body_c * ((T)tb.bone_len[k] - mu * b), b);
                        g[0] = fma(coef, v0, g[0]); g[1] = fma(coef, v1, g[1]); g[2] = fma(coef, v2, g[2]);
                    }
                }
            }
        }
        gout[e * 3 + 0] = g[0]; gout[e * 3 + 1] = g[1]; gout[e * 3 + 2] = g[2];
        gnf += g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
    }
    return (double)gnf;
}

template <typename T>
__global__ void __launch_bounds__(RF_THREADS)
refine_grad_kernel(const __grid_constant__ mc3d_refine_problem pb, int parity) {
    __shared__ double red[8];
    __shared__ RefineTables tb;
    __shared__ __align__(16) T camf[MC3D_MAX_VIEWS * CAM_STRIDE];
    __shared__ double tot[8];
    double *ctrl = pb.ctrl;
    const double *state = ctrl + CT_STATE + 16 * parity;
    if (state[5] != 0.0) return;                                   // stopped
    const bool xchg = xchg_on(pb);
    const long long adam_step = (long long)state[0];
    if (xchg) xchg_gather<7>(pb, parity, false, adam_step + 1, tot);
    const RefineDerived dv = derive(pb, xchg ? tot : ctrl + CT_ACC + 16 * parity, state);
    load_cameras_and_tables(pb, tb, camf);
    __syncthreads();
    double gn[1] = {grad_loop<T>(pb, tb, camf, dv)};
    block_reduce_add<1>(gn, red, ctrl + CT_ACC + 16 * parity + 7);
    if (xchg) xchg_publish<1, 7>(pb, parity, ctrl + CT_ACC + 16 * parity + 7, 1, true, adam_step + 1);
}

// ---- kernel C: clip + Adam + bookkeeping ------------------------------------------------------------------------
// Clip + Adam over this rank's elements, boundary frames stored into the neighbours' halos (in-kernel exchange), and
// the bookkeeping for the next step (block 0).  acc: the 8 global sums of this step; st: state entering it.
// Returns true when this thread stored into a peer's memory.
// Elements between the four gradient components in `gc`: n_frames x J x 3 rounded up to a multiple of 4, so that each
// component starts 16-byte aligned (pairs of scalars are read with one access).
__host__ __device__ __forceinline__ long long gc_stride(const mc3d_refine_problem &pb) {
    return ((long long)pb.n_frames * pb.n_joints * 3 + 3) / 4 * 4;
}

// Two-phase step: the gradient is combined here from its four stored components with the scalars of this step.
template <typename T>
struct GradMix {
    bool on;
    T alpha, sigma, beta, gamma;           // g = alpha g1 + sigma gs + beta G2' + gamma G3
    double mu;                             // this step's a.b / b.b (next step's mu_prev)
    bool three;                            // three stored components: g = gA + beta G2' + gamma G3, gA = alpha g1 + sigma gs
};

// Scalars of one optimiser step that every thread derives from the step's sums and the state entering it: the clip factor,
// Adam's bias corrections, the running-mean early-stopping bookkeeping (pose_refinement.py:1069-1089, quirk Q5) and what
// happens to the best-trajectory snapshot.
template <typename T>
struct StepScalars {
    double step, run_sum, run_cnt, best, no_imp, iters;
    bool improved, stop, clip_nan, wr_new, wr_old;
    T step_size, inv_bc2_sqrt, w1, b2, w2, eps, clipT;
    __device__ __forceinline__ void adam(T gi, T &mi, T &vi, T &xi) const {
        gi = clip_nan ? (T)NAN : gi * clipT;
        mi = mi + (gi - mi) * w1;                                  // exp_avg.lerp_(grad, 1 - beta1)
        vi = vi * b2 + w2 * gi * gi;                               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const T denom = sqrt_c(vi) * inv_bc2_sqrt + eps;
        xi = xi - step_size * div_c(mi, denom);                    // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
};

// Contains one block barrier (the two pow() are evaluated by one thread per block).
template <typename T>
__device__ __forceinline__ StepScalars<T> step_scalars(const mc3d_refine_problem &pb, int end_of_iteration, double gnorm2,
                                                       const double *st, const RefineDerived &dv, double *bias, bool *best_pending,
                                                       bool last_of_launch, bool bias_ready = false) {
    StepScalars<T> ss;
    const double gnorm = sqrt(gnorm2);
    const double clip = fmin(1.0, 1.0 / (gnorm + 1e-6));           // torch clip_grad_norm_(max_norm=1.0)
    ss.step = st[0] + 1.0;
    ss.run_sum = st[1] + dv.total; ss.run_cnt = st[2] + 1.0;
    ss.best = st[3]; ss.no_imp = st[4]; ss.iters = st[6];
    ss.improved = false; ss.stop = false;
    if (end_of_iteration) {
        const double mean = ss.run_sum / ss.run_cnt;               // running mean over costs AND earlier means (Q5)
        ss.run_sum += mean; ss.run_cnt += 1.0;
        ss.improved = mean < ss.best - pb.tolerance;
        if (ss.improved) { ss.best = mean; ss.no_imp = 0.0; } else { ss.no_imp += 1.0; }
        ss.iters += 1.0;
        ss.stop = (ss.no_imp >= (double)pb.patience) || (ss.iters > (double)pb.max_iter);
    }
    // The best-trajectory snapshot (x after the update of a step whose running mean improved).  The persistent kernel defers it
    // (best_pending): while consecutive steps improve, the snapshot would be overwritten again at once, so nothing is written
    // until a step does NOT improve -- then the value of x from before this step's update is the snapshot -- or the launch
    // ends.  Saves 3 of the 12 scalars a step writes per item while the cost is falling.
    ss.wr_new = ss.improved; ss.wr_old = false;
    if (best_pending) {
        ss.wr_new = ss.improved && (last_of_launch || ss.stop);
        ss.wr_old = !ss.improved && *best_pending;
        *best_pending = ss.improved && !ss.wr_new;
    }
    if (!bias_ready) {                                             // (the fused sweep evaluates them ahead, off the step's critical path)
        __syncthreads();                                           // bias[] may still be read from the previous step
        if (threadIdx.x == 0) {                                    // two double pow() per block, not per thread
            bias[0] = 1.0 - pow(pb.beta1, ss.step);
            bias[1] = 1.0 - pow(pb.beta2, ss.step);
        }
        __syncthreads();
    }
    ss.step_size = (T)(pb.lr / bias[0]);
    ss.inv_bc2_sqrt = (T)(1.0 / sqrt(bias[1]));
    ss.w1 = (T)(1.0 - pb.beta1); ss.b2 = (T)pb.beta2; ss.w2 = (T)(1.0 - pb.beta2); ss.eps = (T)pb.eps; ss.clipT = (T)clip;
    ss.clip_nan = !(clip == clip);
    return ss;
}

// Block 0, thread 0: the state entering the next step, the cost history of this one.
template <typename T>
__device__ __forceinline__ void write_next_state(const mc3d_refine_problem &pb, int parity, const StepScalars<T> &ss,
                                                 const RefineDerived &dv, double mu, bool both_parities, const double *counts) {
    double *ctrl = pb.ctrl;
    double *nx = ctrl + CT_STATE + 16 * (parity ^ 1);
    nx[0] = ss.step; nx[1] = ss.run_sum; nx[2] = ss.run_cnt; nx[3] = ss.best; nx[4] = ss.no_imp;
    nx[5] = ss.stop ? 1.0 : 0.0; nx[6] = ss.iters; nx[7] = ss.improved ? 1.0 : 0.0;
    nx[8] = mu;
    nx[9] = counts ? counts[0] : 0.0; nx[10] = counts ? counts[1] : 0.0;   // N_lik, N_s of this step (the next one assumes them)
    if (both_parities && ss.stop)                                  // a persistent kernel leaves its loop here: the stopped
        for (int i = 0; i < 11; ++i) ctrl[CT_STATE + 16 * parity + i] = nx[i];     // state must be found at either parity
    for (int i = 0; i < 8; ++i) ctrl[CT_ACC + 16 * (parity ^ 1) + i] = 0.0;
    const long long hs = (long long)(ss.step - 1.0);
    if (hs < pb.hist_capacity) {
        double *h = ctrl + CT_HIST + 4 * hs;
        h[0] = dv.total; h[1] = dv.cost_lik; h[2] = dv.cost_s; h[3] = dv.cost_b;
    }
}

// Flag the neighbours' halos (after the caller's system fence).
__device__ __forceinline__ void halo_flags(const mc3d_refine_problem &pb, long long seq) {
    if (pb.rank > 0) st_relaxed_sys(&xchg_of(pb, pb.rank - 1)->halo_seq[1], seq);      // after the caller's system fence
    if (pb.rank < pb.world - 1) st_relaxed_sys(&xchg_of(pb, pb.rank + 1)->halo_seq[0], seq);
}

// boundary_first (persistent kernel, several ranks): block 0 updates the elements of the first / last two frames before
// anything else, stores them into the neighbours' halos and flags them at once, so that the halo travels while the
// rest of the grid is still in its Adam pass; every block skips those elements in its regular share.
template <typename T>
__device__ __forceinline__ bool step_loop(const mc3d_refine_problem &pb, int parity, int end_of_iteration, double gnorm2,
                                          const double *st, const RefineDerived &dv, bool xchg, double *bias, bool both_parities,
                                          const GradMix<T> mix, bool boundary_first = false, long long halo_flag_seq = 0,
                                          bool *best_pending = nullptr, bool last_of_launch = true, const double *counts = nullptr,
                                          const StepScalars<T> *pre = nullptr) {
    const StepScalars<T> ss = pre ? *pre : step_scalars<T>(pb, end_of_iteration, gnorm2, st, dv, bias, best_pending, last_of_launch);
    const bool improved = ss.improved, wr_new = ss.wr_new, wr_old = ss.wr_old;
    T *x = (T *)pb.x + 2LL * pb.n_joints * 3;                       // skip the two halo frames
    T *m = (T *)pb.m, *v = (T *)pb.v, *bestx = (T *)pb.best;
    const T *g = (const T *)pb.g;
    const long long n3 = gc_stride(pb);                             // elements per gradient component
    const T *c1 = (const T *)pb.gc, *cs = c1 + n3, *c2 = cs + n3, *c3 = c2 + n3;
    const long long per_frame = (long long)pb.n_joints * 3;
    const long long n = pb.n_frames * per_frame;
    const long long lo = (pb.win_begin - pb.frame_offset) * per_frame, hi = (pb.win_end - pb.frame_offset) * per_frame;
    auto adam = [&](T gi, T &mi, T &vi, T &xi) { ss.adam(gi, mi, vi, xi); };
    // two scalars per access: the halo offset of x (2 J 3 scalars) is always 8-byte (float) / 16-byte (double) aligned
    struct alignas(2 * sizeof(T)) Vec2 { T a, b; };
    const long long n2 = n >> 1;
    const long long tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    // in-kernel exchange: my first / last two frames are the neighbours' halo frames -- store them there directly
    const long long halo_n = 2 * per_frame;
    T *left_halo = nullptr, *right_halo = nullptr;
    if (xchg && pb.rank > 0)                   // right halo of rank - 1: its frames n_left + 2, n_left + 3
        left_halo = reinterpret_cast<T *>(reinterpret_cast<char *>(pb.xchg[pb.rank - 1]) + MC3D_XCHG_X_OFFSET) + (pb.n_frames_left + 2) * per_frame;
    if (xchg && pb.rank < pb.world - 1)        // left halo of rank + 1: its frames 0, 1
        right_halo = reinterpret_cast<T *>(reinterpret_cast<char *>(pb.xchg[pb.rank + 1]) + MC3D_XCHG_X_OFFSET);
    bool pushed = false;
    auto push = [&](long long i, T xi) {
        if (left_halo && i < halo_n) { left_halo[i] = xi; pushed = true; }
        if (right_halo && i >= n - halo_n) { right_halo[i - (n - halo_n)] = xi; pushed = true; }
    };
    auto grad_at = [&](long long i) -> T {
        T gi = !mix.on ? g[i]
             : mix.three ? fma(mix.beta, cs[i], fma(mix.gamma, c2[i], c1[i]))      // components stored as gA, G2', G3
                         : fma(mix.alpha, c1[i], fma(mix.sigma, cs[i], fma(mix.beta, c2[i], mix.gamma * c3[i])));
        if (!(i >= lo && i < hi)) gi = (T)0;
        return gi;
    };
    auto one_element = [&](long long i) {
        T gi = grad_at(i), mi = m[i], vi = v[i], xi = x[i];
        if (wr_old) bestx[i] = xi;
        adam(gi, mi, vi, xi);
        m[i] = mi; v[i] = vi; x[i] = xi;
        if (wr_new) bestx[i] = xi;
        if (xchg) push(i, xi);
    };
    const bool bf = boundary_first && (left_halo || right_halo);
    auto is_boundary = [&](long long i) { return (left_halo && i < halo_n) || (right_halo && i >= n - halo_n); };
    if (bf && blockIdx.x == 0) {
        // at most 2 * halo_n elements: [0, halo_n) and [n - halo_n, n) (they may overlap on a tiny shard)
        const long long second = (n - halo_n > halo_n) ? n - halo_n : halo_n;
        const long long n_first = halo_n < n ? halo_n : n, n_b = n_first + (n > second ? n - second : 0);
        for (long long q = threadIdx.x; q < n_b; q += blockDim.x) {
            const long long i = q < n_first ? q : second + (q - n_first);
            if (is_boundary(i)) one_element(i);
        }
        if (pushed) fence_sys();
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_sys();
            halo_flags(pb, halo_flag_seq);
        }
        pushed = false;
    }
    // Interior fast path (the common case: whole-window batch, two-phase step): every element lies inside the window and the
    // boundary elements -- if there are neighbours at all -- are block 0's, so the loop carries no per-element tests; pairs of
    // elements [e0, e1) with e0 even.  Everything else (windows of a batch, the graph variant's per-element halo pushes, the
    // three-phase step) takes the general loop below.
    const bool fast = mix.on && lo <= 0 && hi >= n && (bf || !(left_halo || right_halo));
    if (fast) {
        const long long e0 = (bf && left_halo) ? (halo_n < n ? halo_n : n) : 0;
        long long e1 = (bf && right_halo) ? n - halo_n : n;
        if (e1 < e0) e1 = e0;
        const Vec2 *q1 = reinterpret_cast<const Vec2 *>(c1), *qs = reinterpret_cast<const Vec2 *>(cs);
        const Vec2 *q2 = reinterpret_cast<const Vec2 *>(c2), *q3 = reinterpret_cast<const Vec2 *>(c3);
        Vec2 *pm = reinterpret_cast<Vec2 *>(m), *pv = reinterpret_cast<Vec2 *>(v), *px = reinterpret_cast<Vec2 *>(x);
        Vec2 *pbest = reinterpret_cast<Vec2 *>(bestx);
        const T al = mix.alpha, si = mix.sigma, be = mix.beta, ga = mix.gamma;
        if (mix.three) {
            for (long long i2 = (e0 >> 1) + tid0; i2 < (e1 >> 1); i2 += nthr) {
                const Vec2 a1 = q1[i2], a2 = qs[i2], a3 = q2[i2];              // gA, G2', G3
                Vec2 mv = pm[i2], vv = pv[i2], xv = px[i2];
                const T ga_ = fma(be, a2.a, fma(ga, a3.a, a1.a));
                const T gb_ = fma(be, a2.b, fma(ga, a3.b, a1.b));
                if (wr_old) pbest[i2] = xv;
                adam(ga_, mv.a, vv.a, xv.a);
                adam(gb_, mv.b, vv.b, xv.b);
                pm[i2] = mv; pv[i2] = vv; px[i2] = xv;
                if (wr_new) pbest[i2] = xv;
            }
        } else {
            for (long long i2 = (e0 >> 1) + tid0; i2 < (e1 >> 1); i2 += nthr) {
                const Vec2 a1 = q1[i2], as = qs[i2], a2 = q2[i2], a3 = q3[i2];
                Vec2 mv = pm[i2], vv = pv[i2], xv = px[i2];
                const T ga_ = fma(al, a1.a, fma(si, as.a, fma(be, a2.a, ga * a3.a)));
                const T gb_ = fma(al, a1.b, fma(si, as.b, fma(be, a2.b, ga * a3.b)));
                if (wr_old) pbest[i2] = xv;
                adam(ga_, mv.a, vv.a, xv.a);
                adam(gb_, mv.b, vv.b, xv.b);
                pm[i2] = mv; pv[i2] = vv; px[i2] = xv;
                if (wr_new) pbest[i2] = xv;
            }
        }
        if ((e1 & 1) && e1 > e0 && tid0 == 0) one_element(e1 - 1);             // unpaired last interior element
    } else {
    for (long long i2 = tid0; i2 < n2; i2 += nthr) {
        const long long i = i2 << 1;
        if (bf) {
            const bool ba = is_boundary(i), bb = is_boundary(i + 1);
            if (ba || bb) {                                        // a pair that touches the boundary: scalar, boundary part skipped
                if (!ba) one_element(i);
                if (!bb) one_element(i + 1);
                continue;
            }
        }
        Vec2 gv;
        if (mix.on && mix.three) {
            gv.a = grad_at(i); gv.b = grad_at(i + 1);
        } else if (mix.on) {
            const Vec2 a1 = reinterpret_cast<const Vec2 *>(c1)[i2], as = reinterpret_cast<const Vec2 *>(cs)[i2];
            const Vec2 a2 = reinterpret_cast<const Vec2 *>(c2)[i2], a3 = reinterpret_cast<const Vec2 *>(c3)[i2];
            gv.a = fma(mix.alpha, a1.a, fma(mix.sigma, as.a, fma(mix.beta, a2.a, mix.gamma * a3.a)));
            gv.b = fma(mix.alpha, a1.b, fma(mix.sigma, as.b, fma(mix.beta, a2.b, mix.gamma * a3.b)));
        } else {
            gv = reinterpret_cast<const Vec2 *>(g)[i2];
        }
        Vec2 mv = reinterpret_cast<Vec2 *>(m)[i2], vv = reinterpret_cast<Vec2 *>(v)[i2], xv = reinterpret_cast<Vec2 *>(x)[i2];
        if (!(i >= lo && i < hi)) gv.a = (T)0;
        if (!(i + 1 >= lo && i + 1 < hi)) gv.b = (T)0;
        if (wr_old) reinterpret_cast<Vec2 *>(bestx)[i2] = xv;
        adam(gv.a, mv.a, vv.a, xv.a);
        adam(gv.b, mv.b, vv.b, xv.b);
        reinterpret_cast<Vec2 *>(m)[i2] = mv;
        reinterpret_cast<Vec2 *>(v)[i2] = vv;
        reinterpret_cast<Vec2 *>(x)[i2] = xv;
        if (wr_new) reinterpret_cast<Vec2 *>(bestx)[i2] = xv;
        if (xchg) { push(i, xv.a); push(i + 1, xv.b); }
    }
    if ((n & 1) && tid0 == 0 && !(bf && is_boundary(n - 1))) one_element(n - 1);      // odd tail
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) write_next_state<T>(pb, parity, ss, dv, mix.on ? mix.mu : 0.0, both_parities, counts);
    return pushed;
}

template <typename T>
__global__ void __launch_bounds__(RF_THREADS)
refine_step_kernel(const __grid_constant__ mc3d_refine_problem pb, int parity, int end_of_iteration) {
    __shared__ double tot[8];
    __shared__ double bias[2];
    double *ctrl = pb.ctrl;
    const double *st = ctrl + CT_STATE + 16 * parity;
    if (st[5] != 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {                 // carry the stopped state forward
            for (int i = 0; i < 16; ++i) ctrl[CT_STATE + 16 * (parity ^ 1) + i] = ctrl[CT_STATE + 16 * parity + i];
            for (int i = 0; i < 8; ++i) ctrl[CT_ACC + 16 * (parity ^ 1) + i] = 0.0;
        }
        return;
    }
    const bool xchg = xchg_on(pb);
    const long long adam_step = (long long)st[0];
    if (xchg) xchg_gather<8>(pb, parity, true, adam_step + 1, tot);     // every rank has finished its gradient pass
    const double *acc = xchg ? tot : ctrl + CT_ACC + 16 * parity;
    const RefineDerived dv = derive(pb, acc, st);
    const bool pushed = step_loop<T>(pb, parity, end_of_iteration, acc[7], st, dv, xchg, bias, false, GradMix<T>{false, 0, 0, 0, 0, 0.0, false});
    if (xchg && pb.world > 1) {                                    // flag the halos once every block's stores are out
        __shared__ int is_last;
        if (pushed) __threadfence_system();
        __syncthreads();
        mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(&mine->ticket[2]), 1ULL);
            is_last = old == (unsigned long long)gridDim.x - 1ULL;
            if (is_last) {
                mine->ticket[2] = 0;
                fence_sys();
                halo_flags(pb, adam_step + 1);
            }
        }
    }
}

// ---- grid barrier of the persistent kernel --------------------------------------------------------------------------
// Every block takes a ticket; the last one resets the counter and releases the generation flag the others poll
// (relaxed polls, one fence on each side).  A wait that never ends (the blocks were not co-resident) gives up after the
// time-out like the cross-rank waits.
__device__ __forceinline__ void grid_barrier(const mc3d_refine_problem &pb, int idx, long long seq) {
    __shared__ int is_last;
    mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_gpu();
        const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(&mine->ticket[idx]), 1ULL);
        is_last = old == (unsigned long long)gridDim.x - 1ULL;
        if (is_last) {
            mine->ticket[idx] = 0;
            fence_gpu();
            st_relaxed_sys(&mine->gen[idx], seq);
        } else {
            xchg_wait(pb, &mine->gen[idx], seq);
        }
        fence_gpu();
    }
    __syncthreads();
}

// ==== two-phase step ================================================================================================
// Pass 1 (costs + gradient components).  The gradient is linear in three scalars that need the global sums,
//     g = alpha g1 + sigma gs + beta (G2 - mu G3),
// so this pass stores g1, gs, G2' = G2 - mu_prev G3 and G3 (mu_prev = last step's mu: G2' is the small, already
// cancelled combination, and the coefficient of G3 becomes the tiny beta (mu_prev - mu)) together with the 10 dot
// products that give |g|^2 as a quadratic form.  sums: [0..7) as in pass A, [7..17) = 11 1s 12 13 ss s2 s3 22 23 33.
constexpr int NS2 = MC3D_REFINE_SUMS2;

// Everything pass 1 needs to know that does not depend on the item.
template <typename T>
struct P1Const {
    int J, C, JS, lo, hi;
    bool ign, do_smooth, do_body, affine_k;
    long long gstride;
    T mu_prev;
    T alpha_a, sigma_a;                                            // three-component form: the 1/N_lik and 2 lambda_s/N_s assumed
    const unsigned char *tok;                                      // tok[t] = smoothness term ending at local frame t
};

template <typename T>
__device__ __forceinline__ P1Const<T> p1_const(const mc3d_refine_problem &pb, T mu_prev) {
    P1Const<T> k;
    k.J = pb.n_joints; k.C = pb.n_cams; k.JS = k.J * 3;
    const long long lo_ = pb.win_begin - pb.frame_offset, hi_ = pb.win_end - pb.frame_offset;
    k.lo = (int)(lo_ < -4 ? -4 : (lo_ > pb.n_frames + 4 ? pb.n_frames + 4 : lo_));     // clamped: only comparisons with
    k.hi = (int)(hi_ < -4 ? -4 : (hi_ > pb.n_frames + 4 ? pb.n_frames + 4 : hi_));     // t - 2 .. t + 2 matter
    k.ign = pb.ignore_distortions != 0;
    k.affine_k = true;
    for (int c = 0; c < pb.n_cams; ++c)                            // read from where the kernels take their cameras
        for (int i = 6; i < 9; ++i) {
            const double kv = pb.cams_dev ? __ldcg(pb.cams_dev + c * 26 + i) : pb.cams[c][i];
            k.affine_k = k.affine_k && kv == (i == 8 ? 1.0 : 0.0);
        }
    k.do_smooth = pb.lambda_smooth > 0.0; k.do_body = pb.lambda_body > 0.0;
    k.gstride = pb.gauss_cam_stride;                               // 0: camera-0 Gaussians for every camera (upstream, Q1)
    k.mu_prev = mu_prev;
    k.alpha_a = k.sigma_a = (T)0;
    k.tok = pb.term_ok + 2;
    return k;
}

// One (frame t, joint j) item of pass 1.  Lean form (round 2): every bone is evaluated once per END POINT with the sign folded
// away -- with u = x_joint - x_other both sign v = u and sign c2 v = c2' u hold exactly -- and its three cost sums are taken by
// the end point that owns it; the smoothness terms are selected, not branched over.  xc points at the item's own position with
// its frame neighbours JS scalars apart (global memory, or a warp's shared-memory window), mup / Sp at its Gaussian (camera
// c's at + c * gstride items, global memory only), tk = the three smoothness flags tok[t], tok[t+1], tok[t+2].  The four
// gradient components go to out[q * ostride + k]; a[] collects the NS2 partial sums.
template <typename T>
struct P1Own {                      // what an item loads for itself: its position and its (camera-0) Gaussian
    T X, Y, Z, mx, my, s00, s01, s11;
};

template <typename T, bool THREE = false>
__device__ __forceinline__ void costgrad_item(const P1Const<T> &pc, const RefineTables &tb, const T *camf, int t, int j,
                                              const T *xc, const P1Own<T> &own, const T *mup, const T *Sp, unsigned tk0,
                                              unsigned tk1, unsigned tk2, T *out, long long ostride, T (&a)[NS2]) {
    const int J = pc.J, JS = pc.JS, lo = pc.lo, hi = pc.hi;
    T g1[3] = {(T)0, (T)0, (T)0}, gs[3] = {(T)0, (T)0, (T)0}, g2[3] = {(T)0, (T)0, (T)0}, g3[3] = {(T)0, (T)0, (T)0};
    if (t >= lo && t < hi) {
        const T X = own.X, Y = own.Y, Z = own.Z;
        const bool self_ok = finite_c(X) && finite_c(Y) && finite_c(Z);
        T mx = own.mx, my = own.my;
        T s00 = own.s00, s01 = own.s01, s11 = own.s11;
        for (int c = 0; c < pc.C; ++c) {
            T cam[CAM_STRIDE];
            load_camera(camf + c * CAM_STRIDE, cam);
            if (pc.gstride && c > 0) {
                const long long ec = c * pc.gstride;
                mx = mup[ec * 2]; my = mup[ec * 2 + 1];
                s00 = Sp[ec * 3]; s01 = Sp[ec * 3 + 1]; s11 = Sp[ec * 3 + 2];
            }
            const T q = reproject_term<true, T>(cam, pc.ign, X, Y, Z, mx, my, s00, s01, s11, (T)1, g1, pc.affine_k);   // adds only when finite
            const bool ok = finite_c(q);
            a[0] += ok ? q : (T)0;
            a[1] += ok ? (T)1 : (T)0;
        }
        if (pc.do_smooth) {
            const T two = (T)2;
            const bool k0 = t - 2 >= lo && tk0;
            const bool k1 = self_ok && t - 1 >= lo && t + 1 < hi && tk1;
            const bool k2 = self_ok && t + 2 < hi && tk2;
            // x has two halo frames on either side: all five frames are addressable whatever the flags say
            const T *xm2 = xc - 2 * JS, *xm1 = xc - JS, *xp1 = xc + JS, *xp2 = xc + 2 * JS;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const T c0 = k == 0 ? X : (k == 1 ? Y : Z);
                const T d0 = c0 - two * xm1[k] + xm2[k];              // the cost term ending at this frame
                const T d1 = xp1[k] - two * c0 + xm1[k];
                const T d2 = xp2[k] - two * xp1[k] + c0;
                a[2] += k0 ? d0 * d0 : (T)0;
                T acc_s = k0 ? d0 : (T)0;
                acc_s = k1 ? acc_s - two * d1 : acc_s;
                acc_s = k2 ? acc_s + d2 : acc_s;
                gs[k] = self_ok ? acc_s : (T)0;                       // a non-finite joint is frozen
            }
            a[3] += (k0 && j == 0) ? (T)1 : (T)0;
        }
        if (pc.do_body && self_ok) {
            const T *xf = xc - j * 3;
            for (int q = tb.adj_start[j]; q < tb.adj_start[j + 1]; ++q) {   // the bones at this joint
                const int2 oo = tb.adj_other_owner[q];
                const T *po = xf + oo.x;
                const T u0 = X - po[0], u1 = Y - po[1], u2 = Z - po[2];      // = sign (end - start)
                const T len = sqrt_c(u0 * u0 + u1 * u1 + u2 * u2);
                if (finite_c(len)) {
                    const T al = (T)tb.adj_len[q];
                    if (oo.y) { a[4] += al * len; a[5] += len * len; a[6] += al * al; }    // the bone's cost sums, once
                    if (len > (T)0) {
                        const T c2 = div_c(al - pc.mu_prev * len, len);                      // G2 - mu_prev G3
                        g2[0] = fma(c2, u0, g2[0]); g2[1] = fma(c2, u1, g2[1]); g2[2] = fma(c2, u2, g2[2]);
                        g3[0] += u0; g3[1] += u1; g3[2] += u2;
                    }
                }
            }
        }
        if (!self_ok) g1[0] = g1[1] = g1[2] = (T)0;
    }
    if (THREE) {
        // gA = alpha g1 + sigma gs with the counts of the previous step (they change only when a value turns non-finite; the
        // caller checks them against this step's totals and repeats the pass in that case); sums [7..13) = AA A2 A3 22 23 33
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const T gA = fma(pc.alpha_a, g1[k], pc.sigma_a * gs[k]);
            out[k] = gA; out[ostride + k] = g2[k]; out[2 * ostride + k] = g3[k];
            a[7] = fma(gA, gA, a[7]);        a[8] = fma(gA, g2[k], a[8]);     a[9] = fma(gA, g3[k], a[9]);
            a[10] = fma(g2[k], g2[k], a[10]); a[11] = fma(g2[k], g3[k], a[11]); a[12] = fma(g3[k], g3[k], a[12]);
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        out[k] = g1[k]; out[ostride + k] = gs[k]; out[2 * ostride + k] = g2[k]; out[3 * ostride + k] = g3[k];
        a[7] = fma(g1[k], g1[k], a[7]);   a[8] = fma(g1[k], gs[k], a[8]);   a[9] = fma(g1[k], g2[k], a[9]);
        a[10] = fma(g1[k], g3[k], a[10]); a[11] = fma(gs[k], gs[k], a[11]); a[12] = fma(gs[k], g2[k], a[12]);
        a[13] = fma(gs[k], g3[k], a[13]); a[14] = fma(g2[k], g2[k], a[14]); a[15] = fma(g2[k], g3[k], a[15]);
        a[16] = fma(g3[k], g3[k], a[16]);
    }
}

// Grid-stride form: one item per thread and trip, neighbours from global memory through L1 (32-bit item indices: a shard
// holds < 2^31 joint-frames).
template <typename T, bool THREE = false>
__device__ __forceinline__ void costgrad_loop(const mc3d_refine_problem &pb, const RefineTables &tb, const T *camf, T mu_prev,
                                              double (&acc)[NS2], T alpha_a = (T)0, T sigma_a = (T)0) {
    P1Const<T> pc = p1_const<T>(pb, mu_prev);
    pc.alpha_a = alpha_a; pc.sigma_a = sigma_a;
    const int J = pc.J;
    const T *x = (const T *)pb.x + 2LL * pc.JS;
    const T *mu0 = (const T *)pb.mu0, *S = (const T *)pb.S;
    const int n_items = (int)(pb.n_frames * J);
    const long long n3 = gc_stride(pb);
    T *o1 = (T *)pb.gc;
    T a[NS2];
#pragma unroll
    for (int i = 0; i < NS2; ++i) a[i] = (T)0;
    const int nthr = (int)(gridDim.x * blockDim.x);
    int e = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    int t = e / J, j = e - t * J;
    const int dt = nthr / J, dj = nthr - dt * J;
    for (; e < n_items; e += nthr, t += dt, j += dj) {
        if (j >= J) { j -= J; ++t; }
        const long long e2 = (long long)e * 2, e3 = (long long)e * 3;
        P1Own<T> own;
        if (t >= pc.lo && t < pc.hi) {
            own.X = x[e3]; own.Y = x[e3 + 1]; own.Z = x[e3 + 2];
            own.mx = mu0[e2]; own.my = mu0[e2 + 1]; own.s00 = S[e3]; own.s01 = S[e3 + 1]; own.s11 = S[e3 + 2];
        }
        costgrad_item<T, THREE>(pc, tb, camf, t, j, x + e3, own, mu0 + e2, S + e3, pc.tok[t], pc.tok[t + 1], pc.tok[t + 2], o1 + e3, n3, a);
    }
#pragma unroll
    for (int i = 0; i < NS2; ++i) acc[i] = (double)a[i];
}

// Scalars of the step from the 17 totals: the usual derived costs, the mix coefficients and |g|^2.
template <typename T>
__device__ __forceinline__ GradMix<T> mix_of(const mc3d_refine_problem &pb, const double *tot, const RefineDerived &dv, double mu_prev,
                                             double &gnorm2) {
    const bool do_smooth = pb.lambda_smooth > 0.0, do_body = pb.lambda_body > 0.0;
    const double al = dv.inv_nlik, si = do_smooth ? dv.smooth_scale : 0.0;
    const double be = do_body ? dv.body_c : 0.0, ga = do_body ? dv.body_c * (mu_prev - dv.mu) : 0.0;
    // rounded to the state type first: |g|^2 is the norm of the gradient that is actually applied
    GradMix<T> m{true, (T)al, (T)si, (T)be, (T)ga, do_body ? dv.mu : 0.0, false};
    const double A = (double)m.alpha, S = (double)m.sigma, B = (double)m.beta, G = (double)m.gamma;
    gnorm2 = A * A * tot[7] + S * S * tot[11] + B * B * tot[14] + G * G * tot[16] +
             2.0 * (A * S * tot[8] + A * B * tot[9] + A * G * tot[10] + S * B * tot[12] + S * G * tot[13] + B * G * tot[15]);
    if (gnorm2 < 0.0) gnorm2 = 0.0;
    return m;
}

// The same for the three-component form (persistent kernel): gA already carries alpha and sigma.
template <typename T>
__device__ __forceinline__ GradMix<T> mix3_of(const mc3d_refine_problem &pb, const double *tot, const RefineDerived &dv, double mu_prev,
                                              double &gnorm2) {
    const bool do_smooth = pb.lambda_smooth > 0.0, do_body = pb.lambda_body > 0.0;
    const double be = do_body ? dv.body_c : 0.0, ga = do_body ? dv.body_c * (mu_prev - dv.mu) : 0.0;
    GradMix<T> m{true, (T)dv.inv_nlik, (T)(do_smooth ? dv.smooth_scale : 0.0), (T)be, (T)ga, do_body ? dv.mu : 0.0, true};
    const double B = (double)m.beta, G = (double)m.gamma;
    gnorm2 = tot[7] + B * B * tot[10] + G * G * tot[12] + 2.0 * (B * tot[8] + G * tot[9] + B * G * tot[11]);
    if (gnorm2 < 0.0) gnorm2 = 0.0;
    return m;
}

// Exchange of the NS2 sums between the two kernels of the graph variant: the last block of pass 1 publishes (plain stores,
// one system fence, a sequence flag per peer), pass 2 gathers.
__device__ __forceinline__ void publish2(const mc3d_refine_problem &pb, int parity, long long seq) {   // last block only
    mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
    if (threadIdx.x < NS2) {
        const double v = __ldcg(&mine->acc2[parity][threadIdx.x]);
        for (int r = 0; r < pb.world; ++r) xchg_of(pb, r)->sums2[parity][pb.rank][threadIdx.x] = v;
        mine->acc2[parity][threadIdx.x] = 0.0;                     // ready for the step after next
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_xchg(pb);
        for (int r = 0; r < pb.world; ++r) st_relaxed_sys(&xchg_of(pb, r)->seq2[parity][pb.rank], seq);
    }
}

__device__ __forceinline__ void gather2(const mc3d_refine_problem &pb, int parity, long long seq, double *tot) {
    mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
    if (threadIdx.x == 0) {
        for (int r = 0; r < pb.world; ++r) xchg_wait(pb, &mine->seq2[parity][r], seq);
        fence_xchg(pb);
    }
    __syncthreads();
    if (threadIdx.x < NS2) {
        double s = 0.0;
        for (int r = 0; r < pb.world; ++r) s += __ldcg(&mine->sums2[parity][r][threadIdx.x]);
        tot[threadIdx.x] = s;
    }
    __syncthreads();
}

// LL exchange of the NS2 sums (persistent kernel): every 8-byte word carries 32 data bits and the 32-bit sequence
// number of the step, stored with one relaxed system-scope store per peer -- no fence between data and flag, no
// separate flag -- and the receivers poll the words of their own block until the sequence matches (the scheme of NCCL's
// LL protocol).  The local rank's words double as the grid barrier.
// `retry`: the repeated pass 1 of a step (its counts of finite terms differed from the assumed ones) goes through its own
// slot, so that the first attempt's words stay in place for blocks and ranks that have not read them yet.
__device__ __forceinline__ void publish2_ll(const mc3d_refine_problem &pb, int parity, long long seq, bool retry) {   // last block only
    mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
    unsigned long long bits = 0;
    if (threadIdx.x < 2 * NS2) bits = (unsigned long long)__double_as_longlong(__ldcg(&mine->acc2[parity][threadIdx.x >> 1]));
    __syncthreads();
    if (threadIdx.x < NS2) mine->acc2[parity][threadIdx.x] = 0.0;  // ready for a repeated pass / the step after next ...
    __syncthreads();
    if (threadIdx.x == 0) fence_gpu();                             // ... before anybody can see the words
    __syncthreads();
    if (threadIdx.x < 2 * NS2) {
        const unsigned long long half = (threadIdx.x & 1) ? (bits >> 32) : (bits & 0xffffffffULL);
        const long long word = (long long)(((unsigned long long)(unsigned int)seq << 32) | half);
        for (int r = 0; r < pb.world; ++r) {
            mc3d_refine_xchg *xr = xchg_of(pb, r);
            st_relaxed_sys(retry ? &xr->ll_retry[pb.rank][threadIdx.x] : &xr->ll[parity][pb.rank][threadIdx.x], word);
        }
    }
}

__device__ __forceinline__ void gather2_ll(const mc3d_refine_problem &pb, int parity, long long seq, double *tot, unsigned int *halves,
                                           bool retry, bool acquire = true) {
    mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
    const unsigned int want = (unsigned int)seq;
    const unsigned long long limit = pb.spin_timeout_ns > 0 ? (unsigned long long)pb.spin_timeout_ns : 10000000000ULL;
    for (int q = threadIdx.x; q < 2 * NS2 * pb.world; q += blockDim.x) {
        const int r = q / (2 * NS2), k = q - r * (2 * NS2);
        const int64_t *w = retry ? &mine->ll_retry[r][k] : &mine->ll[parity][r][k];
        unsigned long long word = (unsigned long long)ld_relaxed_sys(w);
        if ((unsigned int)(word >> 32) != want) {
            const unsigned long long t0 = globaltimer_ns();
            while ((unsigned int)((word = (unsigned long long)ld_relaxed_sys(w)) >> 32) != want)
                if (globaltimer_ns() - t0 > limit) { mine->error = 3; break; }
        }
        halves[q] = (unsigned int)word;
    }
    __syncthreads();
    // acquire what the other blocks of this GPU wrote before their tickets (device scope: the peers' halo stores are acquired
    // where their own flags are waited for).  The fused sweep passes acquire = false: a block reads only what it wrote itself
    // (its range) and its neighbours' edges, which have their own flags and fences.
    if (acquire && threadIdx.x == 0) fence_gpu();
    if (threadIdx.x < NS2) {
        double s = 0.0;
        for (int r = 0; r < pb.world; ++r) {
            const unsigned long long lo = halves[r * 2 * NS2 + 2 * threadIdx.x], hi = halves[r * 2 * NS2 + 2 * threadIdx.x + 1];
            s += __longlong_as_double((long long)((hi << 32) | lo));
        }
        tot[threadIdx.x] = s;
    }
    __syncthreads();
}

__device__ __forceinline__ bool take_ticket(mc3d_refine_xchg *mine, int idx) {      // true in the last block
    __shared__ int is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_gpu();
        const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(&mine->ticket[idx]), 1ULL);
        is_last = old == (unsigned long long)gridDim.x - 1ULL;
        if (is_last) { mine->ticket[idx] = 0; fence_gpu(); }
    }
    __syncthreads();
    return is_last != 0;
}

template <typename T>
__global__ void __launch_bounds__(RF_THREADS)
refine_costgrad_kernel(const __grid_constant__ mc3d_refine_problem pb, int parity) {
    __shared__ double red[8 * NS2];
    __shared__ RefineTables tb;
    __shared__ __align__(16) T camf[MC3D_MAX_VIEWS * CAM_STRIDE];
    const double *st = pb.ctrl + CT_STATE + 16 * parity;
    if (st[5] != 0.0) return;                                      // stopped
    mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
    const long long adam_step = (long long)st[0];
    if (pb.world > 1 && threadIdx.x == 0) {                        // the neighbours' boundary frames of this step are in my halo
        if (pb.rank > 0) xchg_wait(pb, &mine->halo_seq[0], adam_step);
        if (pb.rank < pb.world - 1) xchg_wait(pb, &mine->halo_seq[1], adam_step);
        fence_sys();
    }
    load_cameras_and_tables(pb, tb, camf);
    __syncthreads();
    double acc[NS2];
    costgrad_loop<T>(pb, tb, camf, (T)st[8], acc);
    block_reduce_add<NS2>(acc, red, mine->acc2[parity]);
    if (take_ticket(mine, 0)) publish2(pb, parity, adam_step + 1);
}

template <typename T>
__global__ void __launch_bounds__(RF_THREADS)
refine_step2_kernel(const __grid_constant__ mc3d_refine_problem pb, int parity) {
    __shared__ double tot[24];
    __shared__ double bias[2];
    double *ctrl = pb.ctrl;
    const double *st = ctrl + CT_STATE + 16 * parity;
    if (st[5] != 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0)                   // carry the stopped state forward
            for (int i = 0; i < 16; ++i) ctrl[CT_STATE + 16 * (parity ^ 1) + i] = ctrl[CT_STATE + 16 * parity + i];
        return;
    }
    mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
    const long long adam_step = (long long)st[0];
    gather2(pb, parity, adam_step + 1, tot);                       // every rank has finished pass 1
    const RefineDerived dv = derive(pb, tot, st);
    double gnorm2;
    const GradMix<T> mix = mix_of<T>(pb, tot, dv, (double)(T)st[8], gnorm2);   // mu_prev as pass 1 used it
    const bool pushed = step_loop<T>(pb, parity, 1, gnorm2, st, dv, true, bias, false, mix);
    if (pb.world > 1) {                                            // flag the halos once every block's stores are out
        if (pushed) fence_sys();
        if (take_ticket(mine, 2) && threadIdx.x == 0) {
            fence_sys();
            halo_flags(pb, adam_step + 1);
        }
    }
}

// Pass 1 as a grid-stride loop with its totals exchanged (grid barrier + cross-rank sums in one), repeated once when the counts
// of finite terms it assumed (`counts`, updated) are not the ones it found.  first_attempt = 1: only the repetition (the caller
// has made the first attempt in another form).
template <typename T>
__device__ __forceinline__ void pass1_checked(const mc3d_refine_problem &pb, const RefineTables &tb, const T *camf, int parity,
                                              long long seq, T mu_prev, double (&counts)[2], int first_attempt, double *red,
                                              double *tot, unsigned int *halves) {
    mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
    const bool do_smooth = pb.lambda_smooth > 0.0;
    for (int attempt = first_attempt; attempt < 2; ++attempt) {
        const T alpha_a = counts[0] > 0.0 ? (T)(1.0 / counts[0]) : (T)0;
        const T sigma_a = (do_smooth && counts[1] > 0.0) ? (T)(2.0 * pb.lambda_smooth / counts[1]) : (T)0;
        {
            double acc[NS2];
            costgrad_loop<T, true>(pb, tb, camf, mu_prev, acc, alpha_a, sigma_a);
            block_reduce_add<NS2>(acc, red, mine->acc2[parity]);
        }
        if (take_ticket(mine, 0)) publish2_ll(pb, parity, seq, attempt != 0);
        gather2_ll(pb, parity, seq, tot, halves, attempt != 0);
        const bool same = tot[1] == counts[0] && (!do_smooth || tot[3] == counts[1]) &&
                          !(attempt == 0 && (pb.test_flags & 1) && (seq & 3) == 0);             // (test hook: see mc3d.h)
        counts[0] = tot[1]; counts[1] = tot[3];
        if (same) break;                                           // identical decision in every block and rank
    }
}

// ==== fused sweep =======================================================================================================
// Adam of step s and pass 1 of step s + 1 in ONE sweep over the shard, and one grid-wide meeting per step (the exchange of the
// sums) instead of two.  Pass 2 is bound by memory (it streams 18 scalars per item and does ~60 operations on them), pass 1 by
// instruction issue (~700 per item on 8 loaded scalars); run one after the other, each leaves the other resource idle.  Here
// every block owns a contiguous range of items and walks it in chunks of one item per thread; in one trip a thread
//   1. starts the asynchronous copies (cp.async, 4 / 8 bytes each, coalesced) of the 18 scalars of three elements of Adam chunk
//      i + 1 -- gA, G2', G3, m, v, x -- into its own shared-memory slots,
//   2. evaluates pass 1 of step s + 1 for one item of chunk i (the copies are in flight meanwhile),
//   3. waits for its copies, applies Adam of step s to the three elements and stores m, v, x,
// then the block meets (the next chunk's pass 1 reads the x just written).  Adam chunks start one 2-frame edge (2 J items) into
// the range, so that chunk i of pass 1 finds x_{s+1} on all items it touches -- itself, two frames either side, its frame's
// bones -- once Adam chunk i is done; a trip's Adam writes begin exactly where its pass 1 reads end.  The two edges of the range
// are updated first and announced with a per-block flag (blk_seq, the Adam step count): the neighbouring blocks read them,
// the rank's first / last block also stores them into the neighbour ranks' halo frames.  In-place: every element of x, m, v
// and the components is read and written once per step, as in the two-pass form.
// Everything is staged in 8-byte units -- two floats or one double: three units of (gA, G2', G3, m, v, x) of the next Adam
// chunk per thread, and per trip as many pass-1 items per thread as a unit holds elements (2 in float, 1 in double).
template <typename T> constexpr int sweep_items() { return (int)(8 / sizeof(T)); }
// staged scalars per thread: the 18 units + two buffers of (mu0, Sigma^-1) for the items of the next / the current trip + four
// tile slots of updated x (pass 1 reads x there): 320 bytes per thread in either type
template <typename T> constexpr int sweep_stage() { return 18 * sweep_items<T>() + 2 * 5 * sweep_items<T>() + 4 * 3 * sweep_items<T>(); }

template <int BYTES>
__device__ __forceinline__ void cp_async_bytes(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Item range [lo, hi) of block `q` of `G` in the fused sweep (include/mc3d.h: mc3d_refine_sweep_range).  Boundaries at multiples
// of 4 items, so that pairs of elements are 8-byte aligned in every array.  The block next to a neighbour RANK starts its sweep
// later than the others -- it waits for the halo that crosses NVLink -- and everybody meets again at the end of the step, so
// it gets one trip less than the rest.
__host__ __device__ inline void sweep_range(int n_items, int G, int J, int rank, int world, int trip, int q, int &lo, int &hi) {
    const int per = n_items / G;
    const int cap = per < 480 ? per : 480;
    int short_len = (per - trip > cap ? per - trip : cap) & ~3;
    const bool ok = G >= 4 && short_len >= 4 * J + 32;
    const bool short_l = rank > 0 && ok, short_r = rank < world - 1 && ok;
    const int len_l = short_l ? short_len : 0, len_r = short_r ? short_len : 0;
    const int mid_blocks = G - (short_l ? 1 : 0) - (short_r ? 1 : 0), mid_items = n_items - len_l - len_r;
    auto start = [&](int b) -> int {
        if (b <= 0) return 0;
        if (b >= G) return n_items;
        return len_l + (int)(((long long)mid_items * (b - (short_l ? 1 : 0)) / mid_blocks) & ~3LL);
    };
    lo = start(q);
    hi = start(q + 1);
}

template <typename T>
__device__ __forceinline__ void fused_sweep_run(const mc3d_refine_problem &pb, const RefineTables &tb, const T *camf, int parity,
                                                long long n_iters, double *red, double *tot, double *bias, unsigned int *halves,
                                                T *stage) {
    constexpr int SWEEP_ITEMS = sweep_items<T>();
    struct alignas(8) Unit { T e[SWEEP_ITEMS]; };
    __shared__ T adamc[8];                                         // Adam's constants of the step in the state type
    double *ctrl = pb.ctrl;
    mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
    const int J = pb.n_joints, JS = J * 3, tid = threadIdx.x, NT = RF_THREADS;
    const int n_items = (int)(pb.n_frames * J);
    const int b = blockIdx.x, G = gridDim.x;
    int r_lo, r_hi;                                                // my items
    sweep_range(n_items, G, J, pb.rank, pb.world > 1 ? pb.world : 1, sweep_items<T>() * RF_THREADS, b, r_lo, r_hi);
    const int E = 2 * J;                                           // items of a 2-frame edge
    const bool do_smooth = pb.lambda_smooth > 0.0;
    double st[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) st[i] = __ldcg(ctrl + CT_STATE + 16 * parity + i);
    if (st[5] != 0.0) return;                                      // stopped
    long long seq = (long long)st[0] + 1;
    if (pb.world > 1) {                                            // the neighbours' boundary frames of this step are in my halo
        if (tid == 0) {
            if (pb.rank > 0) xchg_wait(pb, &mine->halo_seq[0], seq - 1);
            if (pb.rank < pb.world - 1) xchg_wait(pb, &mine->halo_seq[1], seq - 1);
            fence_sys();
        }
        __syncthreads();
    }
    double counts[2] = {st[9], st[10]};
    pass1_checked<T>(pb, tb, camf, parity, seq, (T)st[8], counts, 0, red, tot, halves);      // pass 1 of the launch's first step
    bool best_pending = false;
    T *x = (T *)pb.x + 2LL * JS;                                    // local frame 0
    T *m = (T *)pb.m, *v = (T *)pb.v, *bestx = (T *)pb.best;
    const T *mu0 = (const T *)pb.mu0, *S = (const T *)pb.S;
    const long long n3 = gc_stride(pb);
    T *gc = (T *)pb.gc;
    const T *c1 = gc, *c2 = gc + n3, *c3 = gc + 2 * n3;             // gA, G2', G3
    const int n = n_items * 3, halo_n = 2 * JS;
    const long long lo_ = (pb.win_begin - pb.frame_offset) * (long long)JS, hi_ = (pb.win_end - pb.frame_offset) * (long long)JS;
    const int lo_el = (int)(lo_ < 0 ? 0 : (lo_ > n ? n : lo_)), hi_el = (int)(hi_ < 0 ? 0 : (hi_ > n ? n : hi_));
    const bool whole = lo_el <= 0 && hi_el >= n;
    T *left_halo = nullptr, *right_halo = nullptr;                  // my first / last two frames are the neighbours' halo frames
    if (pb.rank > 0 && b == 0)
        left_halo = reinterpret_cast<T *>(reinterpret_cast<char *>(pb.xchg[pb.rank - 1]) + MC3D_XCHG_X_OFFSET) + (pb.n_frames_left + 2) * (long long)JS;
    if (pb.rank < pb.world - 1 && b == G - 1)
        right_halo = reinterpret_cast<T *>(reinterpret_cast<char *>(pb.xchg[pb.rank + 1]) + MC3D_XCHG_X_OFFSET);
    // elements: left edge [eL0, A0), interior units [A0, A1) (A0 and A1 - A0 whole units), right edge and an odd leftover [A1, eR1)
    const int eL0 = r_lo * 3, A0 = eL0 + 3 * E, eR1 = r_hi * 3;
    const int A1 = A0 + ((eR1 - 3 * E - A0) / SWEEP_ITEMS) * SWEEP_ITEMS;
    const int n_edge = (A0 - eL0) + (eR1 - A1);
    constexpr int CI = SWEEP_ITEMS * RF_THREADS, CE = 3 * CI;       // items / elements per trip
    const int n_c = (r_hi - r_lo + CI - 1) / CI, n_a = (A1 - A0 + CE - 1) / CE;
    Unit *stage2 = reinterpret_cast<Unit *>(stage);                 // [18][NT] units
    T *stage_ms = stage + 18 * SWEEP_ITEMS * NT;                    // [2][SWEEP_ITEMS][5][NT]
    // x as Adam left it, chunk k in tile slot k % 3 (and, when k % 3 == 0, also in slot 3, so that chunks k - 1 and k always lie
    // side by side somewhere): everything pass 1 reads of x -- the item itself, two frames either side, its frame's bones -- lies
    // in Adam chunks i - 1 and i, so trip i reads shared memory through ONE base pointer instead of L2, where the item's own
    // position would be the first thing every warp of the block waits for after the trip's barrier.  Items whose window
    // leaves the block's interior (its first / last ~4 J items) read global memory.
    T *xtile = stage_ms + 2 * SWEEP_ITEMS * 5 * NT;                 // [4][CE]
    auto issue = [&](int k) {                                   // Adam chunk k + the Gaussians of pass-1 chunk k
        const int p0 = A0 / SWEEP_ITEMS + k * (CE / SWEEP_ITEMS) + tid;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int pi = p0 + r * NT;
            if (SWEEP_ITEMS * pi < A1) {
                cp_async_bytes<8>(stage2 + (0 * 3 + r) * NT + tid, reinterpret_cast<const Unit *>(c1) + pi);
                cp_async_bytes<8>(stage2 + (1 * 3 + r) * NT + tid, reinterpret_cast<const Unit *>(c2) + pi);
                cp_async_bytes<8>(stage2 + (2 * 3 + r) * NT + tid, reinterpret_cast<const Unit *>(c3) + pi);
                cp_async_bytes<8>(stage2 + (3 * 3 + r) * NT + tid, reinterpret_cast<const Unit *>(m) + pi);
                cp_async_bytes<8>(stage2 + (4 * 3 + r) * NT + tid, reinterpret_cast<const Unit *>(v) + pi);
                cp_async_bytes<8>(stage2 + (5 * 3 + r) * NT + tid, reinterpret_cast<const Unit *>(x) + pi);
            }
        }
        if (k < n_c) {
            T *ms = stage_ms + (k & 1) * (SWEEP_ITEMS * 5 * NT);
#pragma unroll
            for (int u = 0; u < SWEEP_ITEMS; ++u) {
                const int e = r_lo + k * CI + u * NT + tid;
                if (e < r_hi) {
                    cp_async_bytes<sizeof(T)>(ms + (u * 5 + 0) * NT + tid, mu0 + 2LL * e);
                    cp_async_bytes<sizeof(T)>(ms + (u * 5 + 1) * NT + tid, mu0 + 2LL * e + 1);
                    cp_async_bytes<sizeof(T)>(ms + (u * 5 + 2) * NT + tid, S + 3LL * e);
                    cp_async_bytes<sizeof(T)>(ms + (u * 5 + 3) * NT + tid, S + 3LL * e + 1);
                    cp_async_bytes<sizeof(T)>(ms + (u * 5 + 4) * NT + tid, S + 3LL * e + 2);
                }
            }
        }
        cp_async_commit();
    };
    // the operands of a thread's first edge element, fetched before the step's meeting so that the edge update -- the first
    // thing on the next step's critical path -- does not start with an L2 round trip
    T epre[6] = {(T)0, (T)0, (T)0, (T)0, (T)0, (T)0};
    bool edge_pre = false;
    auto fetch_edge = [&]() {
        if (tid < n_edge) {
            const int i = tid < A0 - eL0 ? eL0 + tid : A1 + (tid - (A0 - eL0));
            epre[0] = c1[i]; epre[1] = c2[i]; epre[2] = c3[i]; epre[3] = m[i]; epre[4] = v[i]; epre[5] = x[i];
        }
        edge_pre = true;
    };
    issue(0);                                                       // Adam chunk 0 and the first trip's Gaussians: always a step ahead
    for (long long it = 0; it < n_iters; ++it) {
        const RefineDerived dv = derive(pb, tot, st);
        double gnorm2;
        const GradMix<T> mix = mix3_of<T>(pb, tot, dv, (double)(T)st[8], gnorm2);   // mu_prev as pass 1 used it
        const bool last_iter = it + 1 == n_iters;
        const StepScalars<T> ss = step_scalars<T>(pb, 1, gnorm2, st, dv, bias, &best_pending, last_iter, it > 0);
        if (last_iter || ss.stop) {                                // Adam alone (the launch ends here); halos leave first (block 0)
            cp_async_wait_all();                                   // (the staged chunk is not used)
            __syncthreads();
            if (tid == 0) fence_gpu();                             // this pass reads what other blocks wrote (grid-stride order)
            __syncthreads();
            step_loop<T>(pb, parity, 1, gnorm2, st, dv, true, bias, true, mix, true, seq, nullptr, true, counts, &ss);
            grid_barrier(pb, 2, seq);
            return;
        }
        if (tid == 0) {
            if (b == 0) write_next_state<T>(pb, parity, ss, dv, mix.mu, true, counts);
            adamc[0] = ss.clip_nan ? (T)NAN : ss.clipT; adamc[1] = ss.w1; adamc[2] = ss.b2; adamc[3] = ss.w2;
            adamc[4] = ss.inv_bc2_sqrt; adamc[5] = ss.eps; adamc[6] = ss.step_size;
        }
        __syncthreads();
        const bool wr_old = ss.wr_old, wr_new = ss.wr_new;
        const T beta = mix.beta, gamma = mix.gamma;
        // Adam of one element with the constants from shared memory (registers are pass 1's)
        auto adam_el = [&](int i, T ga, T g2, T g3, T &mi, T &vi, T &xi) {
            T gi = fma(beta, g2, fma(gamma, g3, ga));
            if (!whole && !(i >= lo_el && i < hi_el)) gi = (T)0;
            gi *= adamc[0];                                        // clip (NaN when the norm is)
            mi = mi + (gi - mi) * adamc[1];                        // exp_avg.lerp_(grad, 1 - beta1)
            vi = vi * adamc[2] + adamc[3] * gi * gi;               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
            const T denom = sqrt_c(vi) * adamc[4] + adamc[5];
            xi = xi - adamc[6] * div_c(mi, denom);                 // param.addcdiv_(exp_avg, denom, value=-step_size)
        };
        auto finish = [&](int k) {
            cp_async_wait_all();
            const int p0 = A0 / SWEEP_ITEMS + k * (CE / SWEEP_ITEMS) + tid;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int pi = p0 + r * NT;
                if (SWEEP_ITEMS * pi < A1) {
                    const Unit ga = stage2[(0 * 3 + r) * NT + tid], g2 = stage2[(1 * 3 + r) * NT + tid], g3 = stage2[(2 * 3 + r) * NT + tid];
                    Unit mv = stage2[(3 * 3 + r) * NT + tid], vv = stage2[(4 * 3 + r) * NT + tid], xv = stage2[(5 * 3 + r) * NT + tid];
                    if (wr_old) reinterpret_cast<Unit *>(bestx)[pi] = xv;
#pragma unroll
                    for (int q = 0; q < SWEEP_ITEMS; ++q) adam_el(SWEEP_ITEMS * pi + q, ga.e[q], g2.e[q], g3.e[q], mv.e[q], vv.e[q], xv.e[q]);
                    reinterpret_cast<Unit *>(m)[pi] = mv; reinterpret_cast<Unit *>(v)[pi] = vv; reinterpret_cast<Unit *>(x)[pi] = xv;
                    reinterpret_cast<Unit *>(xtile + (k % 3) * CE)[r * NT + tid] = xv;
                    if (k % 3 == 0) reinterpret_cast<Unit *>(xtile + 3 * CE)[r * NT + tid] = xv;
                    if (wr_new) reinterpret_cast<Unit *>(bestx)[pi] = xv;
                }
            }
        };
        // 1. the two edges of my range, announced at once (a thread's first edge element was fetched before the meeting)
        bool pushed = false;
        for (int q = tid; q < n_edge; q += NT) {
            const int i = q < A0 - eL0 ? eL0 + q : A1 + (q - (A0 - eL0));
            const bool pre = edge_pre && q == tid;
            T mi = pre ? epre[3] : m[i], vi = pre ? epre[4] : v[i], xi = pre ? epre[5] : x[i];
            if (wr_old) bestx[i] = xi;
            adam_el(i, pre ? epre[0] : c1[i], pre ? epre[1] : c2[i], pre ? epre[2] : c3[i], mi, vi, xi);
            m[i] = mi; v[i] = vi; x[i] = xi;
            if (wr_new) bestx[i] = xi;
            if (left_halo && i < halo_n) { left_halo[i] = xi; pushed = true; }
            if (right_halo && i >= n - halo_n) { right_halo[i - (n - halo_n)] = xi; pushed = true; }
        }
        if (pushed) fence_sys();
        __syncthreads();
        // 2. Adam chunk 0 while the neighbours announce theirs
        if (tid == 0) {
            if (left_halo) { fence_sys(); st_relaxed_sys(&xchg_of(pb, pb.rank - 1)->halo_seq[1], seq); }
            if (right_halo) { fence_sys(); st_relaxed_sys(&xchg_of(pb, pb.rank + 1)->halo_seq[0], seq); }
            fence_gpu();
            st_relaxed_sys(&mine->blk_seq[b], seq);                // my edges hold x of step `seq`
            // the left neighbour's edge (thread 32 waits for the right one meanwhile); a system-scope acquire only where a peer stored
            if (b > 0) { xchg_wait(pb, &mine->blk_seq[b - 1], seq); fence_gpu(); }
            else if (pb.rank > 0) { xchg_wait(pb, &mine->halo_seq[0], seq); fence_sys(); }
        } else if (tid == 32) {
            if (b < G - 1) { xchg_wait(pb, &mine->blk_seq[b + 1], seq); fence_gpu(); }
            else if (pb.rank < pb.world - 1) { xchg_wait(pb, &mine->halo_seq[1], seq); fence_sys(); }
        }
        finish(0);
        __syncthreads();
        // 3. the sweep
        P1Const<T> pc = p1_const<T>(pb, (T)mix.mu);
        pc.alpha_a = counts[0] > 0.0 ? (T)(1.0 / counts[0]) : (T)0;
        pc.sigma_a = (do_smooth && counts[1] > 0.0) ? (T)(2.0 * pb.lambda_smooth / counts[1]) : (T)0;
        T a[NS2];
#pragma unroll
        for (int i = 0; i < NS2; ++i) a[i] = (T)0;
        const int dt = NT / J, dj = NT - dt * J;
        int e = r_lo + tid;
        int t = e / J, j = e - t * J;
        for (int i = 0; i < n_c; ++i) {
            issue(i + 1);                                           // Adam chunk i + 1 (if any), the Gaussians of the next trip
            const T *ms = stage_ms + (i & 1) * (SWEEP_ITEMS * 5 * NT);
            // chunks i - 1 and i side by side in the tile slots: element el of x at xwin[el] for lower <= el < A1
            const int lower = A0 + (i > 0 ? i - 1 : 0) * CE;
            const T *xwin = xtile + (i > 0 ? (i - 1) % 3 : 0) * CE - lower;
#pragma unroll
            for (int u = 0; u < SWEEP_ITEMS; ++u, e += NT, t += dt, j += dj) {
                if (j >= J) { j -= J; ++t; }
                if (e < r_hi) {
                    const int el = 3 * e;
                    const bool in_tiles = el - 2 * JS >= lower && el + 2 * JS + 2 < A1;
                    const T *xc = in_tiles ? xwin + el : x + el;
                    const P1Own<T> own{xc[0], xc[1], xc[2], ms[(u * 5 + 0) * NT + tid], ms[(u * 5 + 1) * NT + tid],
                                       ms[(u * 5 + 2) * NT + tid], ms[(u * 5 + 3) * NT + tid], ms[(u * 5 + 4) * NT + tid]};
                    costgrad_item<T, true>(pc, tb, camf, t, j, xc, own, mu0 + 2LL * e, S + 3LL * e, pc.tok[t], pc.tok[t + 1], pc.tok[t + 2],
                                           gc + 3LL * e, n3, a);          // (the pointers serve cameras > 0 of per-camera Gaussians)
                }
            }
            finish(i + 1);
            __syncthreads();
        }
        {
            double acc[NS2];
#pragma unroll
            for (int i = 0; i < NS2; ++i) acc[i] = (double)a[i];
            block_reduce_add<NS2>(acc, red, mine->acc2[parity ^ 1]);
        }
        if (take_ticket(mine, 0)) publish2_ll(pb, parity ^ 1, seq + 1, false);
        fetch_edge();                                               // (my own edges: written by this block, final since its last trip)
        issue(0);                                                   // the next step's Adam chunk 0 travels during the meeting
        if (tid == NT - 1) {                                        // Adam's bias corrections of the next step, while the words travel
            bias[0] = 1.0 - pow(pb.beta1, ss.step + 1.0);
            bias[1] = 1.0 - pow(pb.beta2, ss.step + 1.0);
        }
        gather2_ll(pb, parity ^ 1, seq + 1, tot, halves, false, false);   // the step's one grid-wide meeting (+ cross-rank sums)
        const bool same = tot[1] == counts[0] && (!do_smooth || tot[3] == counts[1]) && !((pb.test_flags & 1) && ((seq + 1) & 3) == 0);
        counts[0] = tot[1]; counts[1] = tot[3];
        if (!same) {
            if (tid == 0) fence_gpu();                             // the repeated pass reads x in grid-stride order
            __syncthreads();
            pass1_checked<T>(pb, tb, camf, parity ^ 1, seq + 1, (T)mix.mu, counts, 1, red, tot, halves);
            edge_pre = false;                                      // the components were written again (by other blocks):
            cp_async_wait_all();                                   // stage chunk 0 anew
            __syncthreads();
            issue(0);
        }
        st[0] = ss.step; st[1] = ss.run_sum; st[2] = ss.run_cnt; st[3] = ss.best; st[4] = ss.no_imp; st[5] = 0.0;
        st[6] = ss.iters; st[7] = ss.improved ? 1.0 : 0.0; st[8] = mix.mu; st[9] = counts[0]; st[10] = counts[1];
        parity ^= 1;
        seq += 1;
    }
}

// All iterations inside one persistent cooperative kernel: two grid barriers per step.
// BLOCKS = CTAs per SM the register allocation is held to: 2 = up to 128 registers (fastest while
// a step is latency-bound), 3 = 80 registers (fastest from ~25 000 frames x 17 joints per GPU up: 122 vs 144 us per
// step at 100 000 frames).
template <typename T, int BLOCKS>
__global__ void __launch_bounds__(RF_THREADS, BLOCKS)
refine_fused2_kernel(const __grid_constant__ mc3d_refine_problem pb, int first_parity, long long n_iters, int sweep) {
    extern __shared__ __align__(16) unsigned char sweep_smem[];   // fused sweep: 18 staged scalars per thread
    __shared__ double red[8 * NS2];
    __shared__ RefineTables tb;
    __shared__ __align__(16) T camf[MC3D_MAX_VIEWS * CAM_STRIDE];
    __shared__ double tot[24];
    __shared__ double bias[2];
    __shared__ unsigned int halves[2 * NS2 * MC3D_MAX_PEERS];
    double *ctrl = pb.ctrl;
    mc3d_refine_xchg *mine = xchg_of(pb, pb.rank);
    load_cameras_and_tables(pb, tb, camf);
    bool best_pending = false;                                     // the best snapshot is owed (see step_flags)
    __syncthreads();
    if (sweep) {
        fused_sweep_run<T>(pb, tb, camf, first_parity & 1, n_iters, red, tot, bias, halves, reinterpret_cast<T *>(sweep_smem));
        return;
    }
    int parity = first_parity & 1;
    for (long long it = 0; it < n_iters; ++it, parity ^= 1) {
        double st[11];                                             // state entering this step (written before the last barrier)
#pragma unroll
        for (int i = 0; i < 11; ++i) st[i] = __ldcg(ctrl + CT_STATE + 16 * parity + i);
        if (st[5] != 0.0) break;                                   // stopped: identical decision in every block and rank
        const long long seq = (long long)st[0] + 1;
        if (pb.world > 1) {
            if (threadIdx.x == 0) {
                if (pb.rank > 0) xchg_wait(pb, &mine->halo_seq[0], seq - 1);
                if (pb.rank < pb.world - 1) xchg_wait(pb, &mine->halo_seq[1], seq - 1);
                fence_sys();
            }
            __syncthreads();
        }
        // pass 1, three stored components: gA = alpha g1 + sigma gs with the counts N_lik, N_s the previous step found (they
        // change only when a value turns non-finite).  If this step's totals say otherwise -- or nothing is known yet: the
        // first step after the state was initialised -- the pass is repeated once with the right counts.
        double counts[2] = {st[9], st[10]};
        pass1_checked<T>(pb, tb, camf, parity, seq, (T)st[8], counts, 0, red, tot, halves);
        const RefineDerived dv = derive(pb, tot, st);
        double gnorm2;
        const GradMix<T> mix = mix3_of<T>(pb, tot, dv, (double)(T)st[8], gnorm2);  // mu_prev as pass 1 used it
        step_loop<T>(pb, parity, 1, gnorm2, st, dv, true, bias, true, mix, true, seq,      // halos leave first (block 0)
                     &best_pending, it + 1 == n_iters, counts);
        grid_barrier(pb, 2, seq);                                   // closes the step; local only
    }
}

// ---- preparation: camera-0 means and inverse covariances (pose_refinement.py:663-668, :885) ----------------------
template <typename T>
__global__ void refine_prepare_kernel(const T *__restrict__ gauss, long long n_frames, int n_cams, int n_joints, int cam,
                                      double eps, T *__restrict__ mu0, T *__restrict__ S) {
    const long long n = n_frames * n_joints;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long t = e / n_joints;
        const int j = (int)(e - t * n_joints);
        const T *gp = gauss + ((t * n_cams + cam) * n_joints + j) * 6;
        const T c00 = gp[2] + (T)eps, c01 = gp[3], c10 = gp[4], c11 = gp[5] + (T)eps;     // cov + eps I in the state dtype
        const double det = (double)c00 * (double)c11 - (double)c01 * (double)c10;
        const double i00 = (double)c11 / det, i11 = (double)c00 / det;
        const double i01 = -0.5 * ((double)c01 + (double)c10) / det;                        // symmetric part: same d^T S d
        mu0[e * 2] = gp[0]; mu0[e * 2 + 1] = gp[1];
        S[e * 3] = (T)i00; S[e * 3 + 1] = (T)i01; S[e * 3 + 2] = (T)i11;
    }
}

// ---- standalone projection (project_points_torch, pose_refinement.py:94-179) -----------------------------------
struct CamParam { double c[26]; };

template <typename T>
__global__ void project_points_kernel(const T *__restrict__ pts, long long n, const __grid_constant__ CamParam cam,
                                      int ignore_dist, T *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double X = (double)pts[i * 3], Y = (double)pts[i * 3 + 1], Z = (double)pts[i * 3 + 2];
        const double *K = cam.c, *R = cam.c + 9, *Tt = cam.c + 18, *D = cam.c + 21;
        const double xc = fma(R[0], X, fma(R[1], Y, fma(R[2], Z, Tt[0])));
        const double yc = fma(R[3], X, fma(R[4], Y, fma(R[5], Z, Tt[1])));
        const double zc = fma(R[6], X, fma(R[7], Y, fma(R[8], Z, Tt[2])));
        const double a = xc / zc, b = yc / zc;
        double xd = a, yd = b;
        if (!ignore_dist) {
            const double r2 = fma(a, a, b * b);
            const double rad = fma(fma(fma(D[4], r2, D[1]), r2, D[0]), r2, 1.0);
            xd = fma(a, rad, fma(2.0 * D[2] * a, b, D[3] * fma(2.0 * a, a, r2)));
            yd = fma(b, rad, fma(D[2], fma(2.0 * b, b, r2), 2.0 * D[3] * a * b));
        }
        const double s = fma(K[6], xd, fma(K[7], yd, K[8]));
        out[i * 2] = (T)(fma(K[0], xd, fma(K[1], yd, K[2])) / s);
        out[i * 2 + 1] = (T)(fma(K[3], xd, fma(K[4], yd, K[5])) / s);
    }
}

template <typename T>
int project_points(const T *d_pts, long long n, const double *cam26, int ignore_dist, T *d_out, cudaStream_t stream) {
    if (n < 0 || !cam26) { set_error("bad arguments"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (n == 0) return MC3D_OK;
    if (!d_pts || !d_out) { set_error("NULL device pointer"); return MC3D_ERR_INVALID_ARGUMENT; }
    CamParam cp;
    for (int i = 0; i < 26; ++i) cp.c[i] = cam26[i];
    long long grid = (n + 255) / 256;
    if (grid > (long long)sm_count() * 8) grid = (long long)sm_count() * 8;
    project_points_kernel<T><<<(unsigned)grid, 256, 0, stream>>>(d_pts, n, cp, ignore_dist, d_out);
    count_launch();
    MC3D_CUDA_TRY(cudaGetLastError());
    return MC3D_OK;
}

static int validate(const mc3d_refine_problem *pb) {
    if (!pb) { set_error("NULL problem"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (pb->n_joints < 1 || pb->n_joints > MC3D_MAX_JOINTS) { set_error("n_joints=%d outside [1, %d]", pb->n_joints, MC3D_MAX_JOINTS); return MC3D_ERR_INVALID_ARGUMENT; }
    if (pb->n_cams < 1 || pb->n_cams > MC3D_MAX_VIEWS) { set_error("n_cams=%d outside [1, %d]", pb->n_cams, MC3D_MAX_VIEWS); return MC3D_ERR_INVALID_ARGUMENT; }
    if (pb->n_bones < 0 || pb->n_bones > MC3D_MAX_BONES) { set_error("n_bones=%d outside [0, %d]", pb->n_bones, MC3D_MAX_BONES); return MC3D_ERR_INVALID_ARGUMENT; }
    if (pb->n_frames < 0) { set_error("n_frames < 0"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (!pb->x || !pb->m || !pb->v || !pb->best || !pb->g || !pb->mu0 || !pb->S || !pb->term_ok || !pb->ctrl) {
        set_error("NULL device pointer in refine problem");
        return MC3D_ERR_INVALID_ARGUMENT;
    }
    if (pb->xchg[0] || pb->world != 0 || pb->rank != 0) {
        if (pb->world < 1 || pb->world > MC3D_MAX_PEERS || pb->rank < 0 || pb->rank >= pb->world) {
            set_error("in-kernel exchange: rank %d / world %d outside [0, %d]", pb->rank, pb->world, MC3D_MAX_PEERS);
            return MC3D_ERR_INVALID_ARGUMENT;
        }
        for (int r = 0; r < pb->world; ++r)
            if (!pb->xchg[r]) { set_error("in-kernel exchange: xchg[%d] is NULL", r); return MC3D_ERR_INVALID_ARGUMENT; }
        if ((char *)pb->x != (char *)pb->xchg[pb->rank] + MC3D_XCHG_X_OFFSET) {
            set_error("in-kernel exchange: x must sit at byte %d of this rank's exchange allocation", MC3D_XCHG_X_OFFSET);
            return MC3D_ERR_INVALID_ARGUMENT;
        }
        if (pb->world > 1 && (pb->n_frames < 2 || (pb->rank > 0 && pb->n_frames_left < 2))) {
            set_error("in-kernel exchange: every rank needs at least two frames");
            return MC3D_ERR_INVALID_ARGUMENT;
        }
    }
    if (pb->gauss_cam_stride != 0 && pb->gauss_cam_stride != pb->n_frames * pb->n_joints) {
        set_error("gauss_cam_stride must be 0 (camera-0 Gaussians) or n_frames * n_joints (per-camera Gaussians)");
        return MC3D_ERR_INVALID_ARGUMENT;
    }
    for (int k = 0; k < pb->n_bones; ++k)
        if (pb->bone_start[k] < 0 || pb->bone_start[k] >= pb->n_joints || pb->bone_end[k] < 0 || pb->bone_end[k] >= pb->n_joints) {
            set_error("bone %d references a joint outside [0, %d)", k, pb->n_joints);
            return MC3D_ERR_INVALID_ARGUMENT;
        }
    return MC3D_OK;
}

template <typename T>
int refine_phase(const mc3d_refine_problem *pb, int phase, long long step_index, int end_of_iteration, cudaStream_t stream) {
    int st = validate(pb);
    if (st != MC3D_OK) return st;
    if (pb->n_frames == 0) return MC3D_OK;
    const int parity = (int)(step_index & 1);
    const int J = pb->n_joints;
    const long long n_items = (long long)pb->n_frames * J;
    long long grid = (n_items + RF_THREADS - 1) / RF_THREADS;
    if (grid > (long long)sm_count() * MC3D_RF_GRID) grid = (long long)sm_count() * MC3D_RF_GRID;
    if (phase == 0) {
        refine_costs_kernel<T><<<(unsigned)grid, RF_THREADS, 0, stream>>>(*pb, parity);
    } else if (phase == 1) {
        refine_grad_kernel<T><<<(unsigned)grid, RF_THREADS, 0, stream>>>(*pb, parity);
    } else if (phase == 2) {
        const long long n = (long long)pb->n_frames * J * 3;
        long long g2 = (n + RF_THREADS * 4 - 1) / (RF_THREADS * 4);
        if (g2 > (long long)sm_count() * 8) g2 = (long long)sm_count() * 8;
        if (g2 < 1) g2 = 1;
        refine_step_kernel<T><<<(unsigned)g2, RF_THREADS, 0, stream>>>(*pb, parity, end_of_iteration);
    } else {
        set_error("bad phase %d", phase);
        return MC3D_ERR_INVALID_ARGUMENT;
    }
    count_launch();
    MC3D_CUDA_TRY(cudaGetLastError());
    return MC3D_OK;
}

// Every rank must pick the same step variant (they meet inside the kernels), so the size that decides is the LARGEST
// shard of the run (frame_shard sizes differ by at most one frame), not this rank's own.
static bool sweep_wanted(const mc3d_refine_problem *pb, size_t elem_size) {
    // float state only: in double the step is bound by the FP64 pipe and its 128 registers already spill, so the sweep's
    // bookkeeping costs more than the overlap returns (measured with one item per thread and trip, 2 CTAs per SM: 223 vs 165 us
    // per step at 100 000 frames, 39 vs 33 at 12 500; parity suite green in both forms)
    const char *envs = getenv("MC3D_REFINE_SWEEP");                 // 0 forbids the fused sweep, 2 allows it for double state (measurement)
    return !(envs && atoi(envs) == 0) && (elem_size == 4 || (envs && atoi(envs) == 2));
}
static bool shard_is_small(const mc3d_refine_problem *pb, size_t elem_size) {
    const long long world = pb->world > 1 ? pb->world : 1;
    const long long frames = pb->world > 1 ? (pb->total_frames + world - 1) / world : pb->n_frames;
    // the fused sweep walks block-owned ranges of any length (measured to 200 000 frames); the double-state two-pass persistent
    // kernel beats the graphs of kernels at every size measured (400 000 frames: 597 vs 803 us per step); the float two-pass
    // form (MC3D_REFINE_SWEEP=0) was measured up to ~150 000 frames x 17 joints
    if (sweep_wanted(pb, elem_size) || elem_size == 8) return frames * pb->n_joints < (1LL << 30);
    return frames * pb->n_joints <= (long long)MC3D_RF_SMALL * sm_count() * 2 * RF_THREADS;
}

// Two-phase step: persistent kernel for small shards, otherwise a CUDA graph of (costgrad, step2) pairs.
template <typename T>
int refine_run_two_phase(const mc3d_refine_problem *pb, long long first_step, long long n_iters, cudaStream_t stream) {
    const long long n_items = (long long)pb->n_frames * pb->n_joints;
    const char *env = getenv("MC3D_REFINE_FUSED");                  // 1 forces the persistent kernel, 0 forbids it
    const int fused_env = env ? atoi(env) : -1;
    const bool small = shard_is_small(pb, sizeof(T));
    mc3d_refine_problem prob = *pb;
    if (fused_env == 1 || (fused_env != 0 && small)) {
        const char *envb = getenv("MC3D_REFINE_BLOCKS");            // 2 / 3 force a register build (measurement)
        int sweep = sweep_wanted(pb, sizeof(T)) ? 1 : 0;
        // the fused sweep is fastest spill-free (2 CTAs x 128 registers: 94 vs 109 us per step at 100 000 frames), and so is the
        // double-state kernel (165 vs 229 us: at 80 registers it spills 1.1 KB); only the float two-pass form of large shards
        // gains from 3 CTAs x 80 registers
        const bool big = envb ? atoi(envb) >= 3 : (!sweep && sizeof(T) == 4 && n_items > 425000);
        auto kern = big ? refine_fused2_kernel<T, 3> : refine_fused2_kernel<T, 2>;
        // Fused sweep (Adam of step s beside pass 1 of step s + 1, one grid-wide meeting per step): every block needs a range
        // that holds its two 2-frame edges and some interior; MC3D_REFINE_SWEEP=0 forbids it (measurement).
        size_t dyn = (size_t)RF_THREADS * sweep_stage<T>() * sizeof(T);
        if (!sweep) dyn = 0;
        if (dyn > 0) { const int as = func_max_smem_once((const void *)kern, 200 * 1024); if (as != MC3D_OK) return as; }
        int per_sm = 0;
        MC3D_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RF_THREADS, dyn));
        if (per_sm < 1) { set_error("persistent refinement kernel does not fit on an SM"); return MC3D_ERR_UNSUPPORTED; }
        if (per_sm > MC3D_RF_GRID) per_sm = MC3D_RF_GRID;
        long long grid = (n_items + RF_THREADS - 1) / RF_THREADS;
        if (grid > (long long)sm_count() * per_sm) grid = (long long)sm_count() * per_sm;      // all blocks co-resident
        if (grid < 1) grid = 1;
        // (the sweep indexes elements with 32-bit integers: 3 n_items must stay below 2^31)
        if (grid > MC3D_XCHG_BLOCK_FLAGS || n_items / grid < 4LL * pb->n_joints + 32 || n_items >= 700000000LL) sweep = 0;
        int parity = (int)(first_step & 1);
        long long iters = n_iters;
        void *args[] = {(void *)&prob, (void *)&parity, (void *)&iters, (void *)&sweep};
        MC3D_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)grid), dim3(RF_THREADS), args, dyn, stream));
        count_launch();
        return MC3D_OK;
    }
    long long grid1 = (n_items + RF_THREADS - 1) / RF_THREADS;
    if (grid1 > (long long)sm_count() * MC3D_RF_GRID) grid1 = (long long)sm_count() * MC3D_RF_GRID;
    long long grid2 = (n_items * 3 + RF_THREADS * 4 - 1) / (RF_THREADS * 4);
    if (grid2 > (long long)sm_count() * 8) grid2 = (long long)sm_count() * 8;
    if (grid2 < 1) grid2 = 1;
    auto one = [&](long long step, cudaStream_t s) -> int {
        const int parity = (int)(step & 1);
        refine_costgrad_kernel<T><<<(unsigned)grid1, RF_THREADS, 0, s>>>(prob, parity);
        refine_step2_kernel<T><<<(unsigned)grid2, RF_THREADS, 0, s>>>(prob, parity);
        MC3D_CUDA_TRY(cudaGetLastError());
        return MC3D_OK;
    };
    long long done = 0;
    int st = MC3D_OK;
    if ((first_step & 1) && n_iters > 0) { st = one(first_step, stream); if (st != MC3D_OK) return st; count_launch(2); done = 1; }
    const long long pairs = (n_iters - done) / 2;
    if (pairs >= 4 && !stream_is_capturing(stream)) {          // a capturing caller gets plain launches (they join its capture)
        static thread_local cudaStream_t cap_stream = nullptr;
        if (!cap_stream) MC3D_CUDA_TRY(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const int unroll = pairs >= 32 ? 8 : 1;
        MC3D_CUDA_TRY(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
        for (int u = 0; u < 2 * unroll && st == MC3D_OK; ++u) st = one(first_step + done + u, cap_stream);
        cudaError_t ce = cudaStreamEndCapture(cap_stream, &graph);
        if (st != MC3D_OK) { if (graph) cudaGraphDestroy(graph); return st; }
        if (ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return cuda_fail(ce, "cudaStreamEndCapture"); }
        GraphGuard guard{graph, nullptr};                          // destroys the graph and its exec on every exit path
        MC3D_CUDA_TRY(cudaGraphInstantiate(&exec, graph, 0));
        guard.exec = exec;
        const long long launches = pairs / unroll;
        for (long long i = 0; i < launches; ++i) MC3D_CUDA_TRY(cudaGraphLaunch(exec, stream));
        count_launch((int)(launches * 2 * unroll * 2));
        done += launches * 2 * unroll;
        MC3D_CUDA_TRY(cudaStreamSynchronize(stream));
    }
    for (; done < n_iters; ++done) { st = one(first_step + done, stream); if (st != MC3D_OK) return st; count_launch(2); }
    return MC3D_OK;
}

// n whole-window iterations on one GPU.  Pairs of iterations (parity 0,1) are captured once into a CUDA graph and
// replayed, so the per-iteration cost is three graph kernel nodes and no host work.
template <typename T>
int refine_run(const mc3d_refine_problem *pb, long long first_step, long long n_iters, cudaStream_t stream) {
    int st = validate(pb);
    if (st != MC3D_OK) return st;
    long long done = 0;
    if (pb->gc && pb->xchg[0] && n_iters > 0 && pb->n_frames > 0) {
        // Measured on B200 (float state, us per step at 400 / 12 500 / 100 000 frames x 17 joints on one GPU):
        //   two-phase persistent 13 / 26 / 122, (a three-phase persistent kernel: 18 / 32 / 144, removed), graph of three kernels
        //   17 / 40 / 134, graph of two kernels 17 / 43 / 164.  The two-phase step moves 25 % more bytes (four gradient
        //   components instead of one) but evaluates the projections once and has one reduction fewer; beyond the
        //   sizes measured (MC3D_RF_SMALL) the byte count is assumed to win and the three-kernel graph is used.
        const char *env2 = getenv("MC3D_REFINE_TWO_PHASE");          // 1 forces the two-phase step, 0 forbids it
        const int two_env = env2 ? atoi(env2) : -1;
        if (two_env == 1 || (two_env != 0 && shard_is_small(pb, sizeof(T)))) return refine_run_two_phase<T>(pb, first_step, n_iters, stream);
    }
    // One rank, big shard: the exchange protocol (tickets, fences, flags) would only cost time -- run the plain kernels.
    mc3d_refine_problem plain = *pb;
    if (pb->world <= 1) {
        plain.world = plain.rank = 0;
        for (int r = 0; r < MC3D_MAX_PEERS; ++r) plain.xchg[r] = nullptr;
        pb = &plain;
    }
    auto one = [&](long long step, cudaStream_t s) -> int {
        for (int ph = 0; ph < 3; ++ph) {
            int s2 = refine_phase<T>(pb, ph, step, 1, s);
            if (s2 != MC3D_OK) return s2;
        }
        return MC3D_OK;
    };
    if ((first_step & 1) && n_iters > 0) { st = one(first_step, stream); if (st != MC3D_OK) return st; done = 1; }
    const long long pairs = (n_iters - done) / 2;
    if (pairs >= 4 && !stream_is_capturing(stream)) {          // a capturing caller gets plain launches (they join its capture)
        // The legacy default stream cannot be captured: record on a private stream, replay on the caller's.
        static thread_local cudaStream_t cap_stream = nullptr;
        if (!cap_stream) MC3D_CUDA_TRY(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const int unroll = pairs >= 32 ? 8 : 1;                     // 2*unroll iterations per graph launch
        MC3D_CUDA_TRY(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
        const long long before = mc3d_launch_count();
        for (int u = 0; u < 2 * unroll && st == MC3D_OK; ++u) st = one(first_step + done + u, cap_stream);
        cudaError_t ce = cudaStreamEndCapture(cap_stream, &graph);
        count_launch((int)(before - mc3d_launch_count()));          // recording is not launching
        if (st != MC3D_OK) { if (graph) cudaGraphDestroy(graph); return st; }
        if (ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return cuda_fail(ce, "cudaStreamEndCapture"); }
        GraphGuard guard{graph, nullptr};                          // destroys the graph and its exec on every exit path
        MC3D_CUDA_TRY(cudaGraphInstantiate(&exec, graph, 0));
        guard.exec = exec;
        const long long launches = pairs / unroll;
        for (long long i = 0; i < launches; ++i) MC3D_CUDA_TRY(cudaGraphLaunch(exec, stream));
        count_launch((int)(launches * 2 * unroll * 3));
        done += launches * 2 * unroll;
        MC3D_CUDA_TRY(cudaStreamSynchronize(stream));               // the exec must outlive its launches
    }
    for (; done < n_iters; ++done) { st = one(first_step + done, stream); if (st != MC3D_OK) return st; }
    return MC3D_OK;
}

// What refine_run would launch for this problem (same decisions, no launch).
static const char *refine_plan(const mc3d_refine_problem *pb) {
    if (!pb) return "invalid";
    const bool small = shard_is_small(pb, 4) && shard_is_small(pb, 8);      // the dtype is not known here: the stricter answer
    const char *e2 = getenv("MC3D_REFINE_TWO_PHASE"), *ef = getenv("MC3D_REFINE_FUSED");
    const int two_env = e2 ? atoi(e2) : -1, fused_env = ef ? atoi(ef) : -1;
    const bool fused = fused_env == 1 || (fused_env != 0 && small);
    if (pb->gc && pb->xchg[0] && (two_env == 1 || (two_env != 0 && small)))
        return fused ? "two-phase step, persistent cooperative kernel (float state: fused sweep -- Adam of step s beside pass 1 of step "
                       "s + 1, one grid-wide meeting per step; double state: two passes, 2 grid barriers; 1 exchange of 13 sums + halo "
                       "stores per step)"
                     : "two-phase step, CUDA graph of 2 kernels per step (1 exchange of 17 sums + halo stores per step)";
    if (pb->xchg[0] && pb->world > 1)
        return "three-phase step, CUDA graph of 3 kernels per step with the in-kernel exchange (2 exchanges + halo stores per step)";
    return "three-phase step, CUDA graph of 3 kernels per step (one rank, or host-driven exchange)";
}

template <typename T>
int refine_flags(const mc3d_refine_problem *pb, cudaStream_t stream) {
    int st = validate(pb);
    if (st != MC3D_OK) return st;
    long long grid = (pb->n_frames + 2 + 127) / 128;
    if (grid > (long long)sm_count() * 8) grid = (long long)sm_count() * 8;
    refine_flags_kernel<T><<<(unsigned)grid, 128, 0, stream>>>(*pb);
    count_launch();
    MC3D_CUDA_TRY(cudaGetLastError());
    return MC3D_OK;
}

template <typename T>
int refine_prepare(const T *d_gauss, long long n_frames, int n_cams, int n_joints, int cam, double eps, T *d_mu0, T *d_S,
                   cudaStream_t stream) {
    if (n_frames < 0 || n_cams < 1 || n_joints < 1 || cam < 0 || cam >= n_cams) { set_error("bad gaussians shape"); return MC3D_ERR_INVALID_ARGUMENT; }
    if (n_frames == 0) return MC3D_OK;
    if (!d_gauss || !d_mu0 || !d_S) { set_error("NULL device pointer"); return MC3D_ERR_INVALID_ARGUMENT; }
    const long long n = n_frames * n_joints;
    long long grid = (n + 255) / 256;
    if (grid > (long long)sm_count() * 8) grid = (long long)sm_count() * 8;
    refine_prepare_kernel<T><<<(unsigned)grid, 256, 0, stream>>>(d_gauss, n_frames, n_cams, n_joints, cam, eps, d_mu0, d_S);
    count_launch();
    MC3D_CUDA_TRY(cudaGetLastError());
    return MC3D_OK;
}

}  // namespace mc3d

extern "C" {
int mc3d_refine_problem_size(void) { return (int)sizeof(mc3d_refine_problem); }
int mc3d_refine_sweep_range(int64_t n_items, int grid, int n_joints, int rank, int world, int elem_size, int block, int64_t *lo, int64_t *hi) {
    if (n_items < 0 || n_items >= 700000000LL || grid < 1 || n_joints < 1 || rank < 0 || world < 1 || rank >= world ||
        (elem_size != 4 && elem_size != 8) || block < 0 || block >= grid || !lo || !hi) {
        mc3d::set_error("mc3d_refine_sweep_range: bad arguments");
        return MC3D_ERR_INVALID_ARGUMENT;
    }
    int a = 0, b = 0;
    mc3d::sweep_range((int)n_items, grid, n_joints, rank, world, (8 / elem_size) * mc3d::RF_THREADS, block, a, b);
    *lo = a; *hi = b;
    return MC3D_OK;
}
const char *mc3d_refine_plan(const mc3d_refine_problem *pb) { return mc3d::refine_plan(pb); }
int mc3d_project_points_f32(const float *d_points, int64_t n, const double *cam26, int ignore_distortions, float *d_out, void *stream) {
    return mc3d::project_points<float>(d_points, n, cam26, ignore_distortions, d_out, (cudaStream_t)stream);
}
int mc3d_project_points_f64(const double *d_points, int64_t n, const double *cam26, int ignore_distortions, double *d_out, void *stream) {
    return mc3d::project_points<double>(d_points, n, cam26, ignore_distortions, d_out, (cudaStream_t)stream);
}
int mc3d_refine_prepare_f32(const float *d_gauss, int64_t n_frames, int n_cams, int n_joints, int cam, double eps,
                            float *d_mu0, float *d_S, void *stream) {
    return mc3d::refine_prepare<float>(d_gauss, n_frames, n_cams, n_joints, cam, eps, d_mu0, d_S, (cudaStream_t)stream);
}
int mc3d_refine_prepare_f64(const double *d_gauss, int64_t n_frames, int n_cams, int n_joints, int cam, double eps,
                            double *d_mu0, double *d_S, void *stream) {
    return mc3d::refine_prepare<double>(d_gauss, n_frames, n_cams, n_joints, cam, eps, d_mu0, d_S, (cudaStream_t)stream);
}
int mc3d_refine_flags_f32(const mc3d_refine_problem *pb, void *stream) { return mc3d::refine_flags<float>(pb, (cudaStream_t)stream); }
int mc3d_refine_flags_f64(const mc3d_refine_problem *pb, void *stream) { return mc3d::refine_flags<double>(pb, (cudaStream_t)stream); }
int mc3d_refine_phase_f32(const mc3d_refine_problem *pb, int phase, int64_t step_index, int end_of_iteration, void *stream) {
    return mc3d::refine_phase<float>(pb, phase, step_index, end_of_iteration, (cudaStream_t)stream);
}
int mc3d_refine_phase_f64(const mc3d_refine_problem *pb, int phase, int64_t step_index, int end_of_iteration, void *stream) {
    return mc3d::refine_phase<double>(pb, phase, step_index, end_of_iteration, (cudaStream_t)stream);
}
int mc3d_refine_run_f32(const mc3d_refine_problem *pb, int64_t first_step, int64_t n_iters, void *stream) {
    return mc3d::refine_run<float>(pb, first_step, n_iters, (cudaStream_t)stream);
}
int mc3d_refine_run_f64(const mc3d_refine_problem *pb, int64_t first_step, int64_t n_iters, void *stream) {
    return mc3d::refine_run<double>(pb, first_step, n_iters, (cudaStream_t)stream);
}
}
